"""Randomised campaign through the CPU simulation of the parallel algorithm (python tools/sim_fuzz.py SEED SECONDS):\ncoefficients against the encoder under random planner settings.  No GPU needed."""
import sys, os, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sim_ffi as S
from jpeg_rust_b200 import synth
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
t_end = time.time() + float(sys.argv[2] if len(sys.argv) > 2 else 600)
n = bad = 0
while time.time() < t_end:
    sub = rng.choice(["420", "420", "422", "444", "440", "gray"])
    w = rng.choice([rng.randint(8, 400), rng.randint(400, 2500), rng.choice([512, 1024, 2048, 4096])])
    h = rng.choice([rng.randint(8, 400), rng.randint(400, 2500), rng.choice([512, 1024, 2048])])
    ri = rng.choice([0, 0, 1, 2, 3, 5, 8, 16, 33, 64, 100, 256, 1000])
    q = rng.choice([30, 60, 85, 95])
    sb = rng.choice([0, 1024, 2048, 4096, 8192])
    os.environ["JPGPU_LOOKBACK_BITS"] = str(rng.choice([64, 256, 1024, 1024, 4096]))
    os.environ["JPGPU_WRITE_PARTS"] = str(rng.choice([1, 1, 2, 4]))
    if rng.random() < 0.3: os.environ["JPGPU_INTERVAL_MODE"] = str(rng.choice([0, 1]))
    else: os.environ.pop("JPGPU_INTERVAL_MODE", None)
    seed = rng.randint(0, 10**6)
    f, gt = synth.synth_jpeg(seed, w, h, sub, quality=q, restart_interval=ri, want_coefs=True)
    rs, diag = S.decode_batch([f], layout=1, ext=2, sub_bits=sb)
    ok = rs[0].status == 0 and all(np.array_equal(a, g) for a, g in zip(rs[0].coefs, gt))
    n += 1
    if not ok:
        bad += 1
        print("MISMATCH", dict(seed=seed, w=w, h=h, sub=sub, ri=ri, q=q, sb=sb, lb=os.environ["JPGPU_LOOKBACK_BITS"], wp=os.environ["JPGPU_WRITE_PARTS"], im=os.environ.get("JPGPU_INTERVAL_MODE")), "status", rs[0].status, flush=True)
print("done", n, "images,", bad, "mismatches", flush=True)
