"""Randomised campaign on the GPU (gpurun -- python tools/gpu_fuzz.py SEED SECONDS): mixed batches decoded under the default\nand a random planner setting; coefficients against the encoder, pixels identical between the settings."""
import sys, os, time, random
sys.path.insert(0, os.getcwd())
import numpy as np
from jpeg_rust_b200 import Batch, LAYOUT_SPEC, EXT_DRI, synth
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 7)
t_end = time.time() + float(sys.argv[2]) if len(sys.argv) > 2 else time.time() + 60
nimg = bad = nb = 0
KEYS = ["JPGPU_SUBSEQ_BITS", "JPGPU_LOOKBACK_BITS", "JPGPU_WRITE_PARTS", "JPGPU_INTERVAL_MODE", "JPGPU_GROUPS", "JPGPU_SEG_BITS",
        "JPGPU_SYNC_MULTI", "JPGPU_VERIFY_MULTI"]
while time.time() < t_end:
    files, gts, meta = [], [], []
    big = rng.random() < 0.15
    for it in range(6 if big else rng.choice([1, 3, 40, 200])):
        sub = rng.choice(["420", "420", "422", "444", "440", "gray"])
        if big:
            w, h = rng.choice([(4096, 4096), (8192, 2048), (6000, 4000), (1920, 1080), (3840, 2160)])
        else:
            w = rng.choice([rng.randint(8, 300), rng.randint(300, 1400), rng.choice([512, 1024, 2048])])
            h = rng.choice([rng.randint(8, 300), rng.randint(300, 1000), rng.choice([512, 1024])])
        ri = rng.choice([0, 0, 1, 2, 3, 5, 8, 16, 33, 64, 100, 256, 1000])
        seed, q = rng.randint(0, 10 ** 6), rng.choice([30, 60, 85, 95])
        f, g = synth.synth_jpeg(seed, w, h, sub, quality=q, restart_interval=ri, want_coefs=True, optimize=rng.random() < 0.3,
                                noise_sigma=rng.choice([6.0, 6.0, 20.0]))
        files.append(f); gts.append(g); meta.append((seed, w, h, sub, ri, q))
    outs = []
    for variant in range(2):
        for k in KEYS: os.environ.pop(k, None)
        if variant == 1:
            os.environ["JPGPU_SUBSEQ_BITS"] = str(rng.choice([1024, 2048, 4096, 8192]))
            os.environ["JPGPU_LOOKBACK_BITS"] = str(rng.choice([64, 256, 1024, 4096]))
            os.environ["JPGPU_WRITE_PARTS"] = str(rng.choice([1, 2, 4]))
            os.environ["JPGPU_GROUPS"] = str(rng.choice([1, 2, 3]))
            if rng.random() < 0.5: os.environ["JPGPU_INTERVAL_MODE"] = str(rng.choice([0, 1]))
            os.environ["JPGPU_SYNC_MULTI"] = str(rng.choice([0, 1, 1]))
            os.environ["JPGPU_VERIFY_MULTI"] = str(rng.choice([0, 1]))
        env = {k: os.environ.get(k) for k in KEYS if os.environ.get(k)}
        b = Batch(files, layout=LAYOUT_SPEC, ext=EXT_DRI)
        b.upload().decode()
        if rng.random() < 0.3:
            b.decode(); b.decode()      # captured into a CUDA graph, replayed
        o = b.download()
        st, _ = b.results()
        for i in range(len(files)):
            ok = st[i] == 0 and all(np.array_equal(a[:len(g)], g[:len(a)]) for a, g in zip(b.coefficients(i), gts[i]))
            if not ok:
                bad += 1
                print("MISMATCH", meta[i], "status", st[i], "env", env, flush=True)
        outs.append(o)
        b.close()
    for i in range(len(files)):
        if not np.array_equal(outs[0][i], outs[1][i]):
            bad += 1
            print("PIXELS DIFFER between planner settings", meta[i], flush=True)
    nimg += len(files); nb += 1
print("done", nb, "batches", nimg, "images,", bad, "problems", flush=True)
