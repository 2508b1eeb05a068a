"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck are too slow for the whole GPU suite):
  compute-sanitizer --tool memcheck python tests/sanitizer_smoke.py
Covers every kernel: fused kernels of all sampling modes, restart intervals, the gather path (REF layout), repairs
(short look-back, multi-symbol tables), damaged streams (zero-filled tails), replan, the pipelined group decode, graph
capture and replay, float output and the host-to-host pipeline."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("JPGPU_LOOKBACK_BITS", "256")   # force plenty of repairs
os.environ.setdefault("JPGPU_GROUPS", "3")

import oracle_ffi as O  # noqa: E402
from jpeg_rust_b200 import EXT_DRI, LAYOUT_REF, LAYOUT_SPEC, Batch, synth  # noqa: E402


def main():
    rng = np.random.default_rng(3)
    files = []
    for i, (sub, ri) in enumerate([("420", 0), ("444", 3), ("gray", 0), ("422", 7), ("440", 0), ("420", 1)]):
        files.append(synth.synth_jpeg(40 + i, 200 + 24 * i, 120 + 8 * i, sub, restart_interval=ri))
    damaged = bytearray(files[0])
    sos = bytes(damaged).index(b"\xff\xda") + 14
    for _ in range(6):
        damaged[int(rng.integers(sos, len(damaged) - 2))] ^= 0x10
    files.append(bytes(damaged))
    files.append(files[1][:len(files[1]) // 2])
    b = Batch(files, ext=EXT_DRI, layout=LAYOUT_SPEC)
    for _ in range(3):           # kernel by kernel, captured into a CUDA graph, replayed
        b.upload().decode()
    outs = b.download()
    statuses, _ = b.results()
    for i in range(6):
        assert statuses[i] == 0, (i, statuses[i])
        ref = O.decode(files[i], layout=O.LAYOUT_SPEC, ext=EXT_DRI).rgb
        assert np.abs(outs[i].astype(int) - ref.astype(int)).max() <= 1
    ref_files = [synth.synth_jpeg(60, 250, 131, "420"), synth.synth_jpeg(61, 61, 45, "444")]
    b.replan(ref_files, layout=LAYOUT_REF)
    b.upload().decode()
    outs = b.download()
    statuses, _ = b.results()
    for i, f in enumerate(ref_files):
        o = O.decode(f, layout=O.LAYOUT_REF)
        assert statuses[i] == o.status == 0
        assert np.abs(outs[i].astype(int) - o.rgb.astype(int)).max() <= 1
    # float planes (converted inside the IDCT/colour kernel), then the host-to-host pipeline with single-copy transfers
    from jpeg_rust_b200 import Pipeline, _ffi
    b.replan(files[:6], ext=EXT_DRI, layout=LAYOUT_SPEC)
    b.set_output_format(_ffi.OUT_F32_PLANAR)
    b.upload().decode()
    f32 = b.download()
    statuses, _ = b.results()
    assert all(s == 0 for s in statuses) and f32[0].dtype == np.float32 and 0.0 <= float(f32[0].min()) and float(f32[0].max()) <= 1.0
    b.close()
    p = Pipeline(files, ext=EXT_DRI, layout=LAYOUT_SPEC, chunk=3)
    p.run().sync()
    st, _ = p.results()
    assert st[:6] == [0] * 6 and st[6] != 0 or st[7] != 0
    for i in range(6):
        ref = O.decode(files[i], layout=O.LAYOUT_SPEC, ext=EXT_DRI).rgb
        assert np.abs(p.image(i).astype(int) - ref.astype(int)).max() <= 1
    p.close()
    # the compose path: a file of non-interleaved scans, chroma interpolated (block IDCT + compose_colour_kernel)
    from jpeg_rust_b200 import LAYOUT_SPEC_FANCY, decode_scans
    planar = [synth.synth_jpeg(70, 250, 131, "420", restart_interval=3, planar_scans=True), synth.synth_jpeg(71, 64, 40, "422", planar_scans=True)]
    outs, st = decode_scans(planar + [files[0]], ext=EXT_DRI, layout=LAYOUT_SPEC_FANCY)
    assert st == [0, 0, 0] and outs[0].shape == (131, 250, 3)
    print("sanitizer smoke ok")


if __name__ == "__main__":
    main()
