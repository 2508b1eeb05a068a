"""Randomised campaign for the compose path (gpurun -- python tools/gpu_fuzz_scans.py SEED SECONDS): files of non-interleaved
scans, decoded through jpgpu_parse_scans in mixed batches with ordinary files, in both up-sampling layouts.  Every scan's
coefficients against the encoder's; pixels against the same coefficients coded as one interleaved scan (+-1); damaged files
must fail alone."""
import sys, os, time, random
sys.path.insert(0, os.getcwd())
import numpy as np
from jpeg_rust_b200 import EXT_DRI, LAYOUT_SPEC, LAYOUT_SPEC_FANCY, Batch, decode_scans, synth
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 11)
t_end = time.time() + (float(sys.argv[2]) if len(sys.argv) > 2 else 60)
nb = nimg = bad = 0
while time.time() < t_end:
    layout = rng.choice([LAYOUT_SPEC, LAYOUT_SPEC_FANCY])
    files, inter, gts, damaged = [], [], [], []
    for it in range(rng.choice([1, 3, 12, 40])):
        sub = rng.choice(["420", "420", "422", "444", "440"])
        w, h = rng.choice([(rng.randint(8, 300), rng.randint(8, 300)), (rng.randint(300, 1400), rng.randint(300, 900)), (1920, 1080)])
        ri = rng.choice([0, 0, 1, 3, 8, 40, 300])
        seed, q = rng.randint(0, 10 ** 6), rng.choice([30, 60, 85, 95])
        planar = rng.random() < 0.7
        f, g = synth.synth_jpeg(seed, w, h, sub, quality=q, restart_interval=ri, want_coefs=True, planar_scans=planar)
        ref = synth.synth_jpeg(seed, w, h, sub, quality=q, restart_interval=ri)
        dmg = planar and rng.random() < 0.15
        if dmg:
            f = bytearray(f)
            k = bytes(f).rindex(b"\xff\xda") + 10
            if rng.random() < 0.5:
                del f[rng.randint(k + 4, len(f) - 1):]
            else:
                for _ in range(4): f[rng.randint(k, len(f) - 3)] ^= 1 << rng.randint(0, 7)
            f = bytes(f)
        files.append(f); inter.append(ref); gts.append(g if planar else None); damaged.append(dmg)
    outs, st, coefs = decode_scans(files, ext=EXT_DRI, layout=layout, want_coefs=True)
    b = Batch(inter, ext=EXT_DRI, layout=layout)
    b.upload().decode()
    want = b.download()
    b.results()
    b.close()
    for i in range(len(files)):
        if damaged[i]:
            continue
        ok = st[i] == 0
        if ok and gts[i] is not None:
            ok = all(c is not None and np.array_equal(c, g) for c, g in zip(coefs[i], gts[i]))
        if ok:
            d = np.abs(outs[i].astype(np.int16) - want[i].astype(np.int16))
            ok = d.max() <= 1
        if not ok:
            bad += 1
            print("PROBLEM image", i, "status", st[i], "layout", layout, flush=True)
    nb += 1; nimg += len(files)
print("done", nb, "batches", nimg, "images,", bad, "problems", flush=True)
