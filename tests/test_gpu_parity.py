"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle.

Gate (BASELINE.json north_star): Huffman-decoded coefficients bit-exact; output samples
within +-1 per 8-bit channel (max and mean |delta| asserted/reported below)."""
import hashlib

import numpy as np
import pytest

import oracle_ffi as O
from conftest import fixture_bytes
from jpeg_rust_b200 import (EXT_DRI, EXT_NONE, EXT_SKIP_APPN, LAYOUT_REF, LAYOUT_SPEC, Batch, JPEGImage, JPEGPanic,
                            _ffi, synth)

pytestmark = pytest.mark.gpu

SAMPLE_TOL = 1          # LSB per 8-bit channel
MEAN_TOL = 0.01         # mean |delta| (observed ~1e-5)


def assert_samples(got, want, what=""):
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert d.max() <= SAMPLE_TOL, f"{what}: max |delta| = {d.max()}"
    assert d.mean() <= MEAN_TOL, f"{what}: mean |delta| = {d.mean()}"
    return int(d.max()), float(d.mean())


def run_batch(files, layout=LAYOUT_SPEC, ext=EXT_NONE):
    b = Batch(files, ext=ext, layout=layout)
    b.upload().decode()
    outs = b.download()
    statuses, br = b.results()
    coefs = [b.coefficients(i) if statuses[i] == 0 else None for i in range(len(files))]
    launches = b.launch_count()
    b.close()
    return outs, statuses, br, coefs, launches


def compare_with_oracle(files, layout, ext=EXT_NONE, gts=None):
    outs, statuses, br, coefs, launches = run_batch(files, layout, ext)
    assert launches > 0
    worst = (0, 0.0)
    for i, f in enumerate(files):
        o = O.decode(f, layout=layout, ext=ext)
        assert o.status == 0, o.msg
        assert statuses[i] == 0, _ffi.status_string(statuses[i])
        assert len(coefs[i]) == len(o.coefs)
        for c, (a, w) in enumerate(zip(coefs[i], o.coefs)):
            assert a.shape == w.shape, f"image {i} comp {c}: {a.shape} vs {w.shape}"
            assert np.array_equal(a, w), f"image {i} comp {c}: coefficients differ"
        if gts is not None:
            for a, w in zip(coefs[i], gts[i]):
                assert np.array_equal(a[:len(w)], w[:len(a)])
        if not (ext & EXT_DRI):
            assert br[i] == o.bytes_read
        m = assert_samples(outs[i], o.rgb, f"image {i}")
        worst = (max(worst[0], m[0]), max(worst[1], m[1]))
    return worst


KNOWN = {  # SURVEY.md §4 known answers
    "lena.jpeg": ("ba5ce1b7b3b108bb2bf354a742b7c1217138d58cefd78abe347e05255c060c06", 90694),
    "lena-bw.jpeg": ("0fa4cc6820aa8a2d1d2448b5ad5f6b036ff007c17ef8b0369c5207655775dea2", 21494),
}


@pytest.mark.parametrize("name", ["lena.jpeg", "lena-bw.jpeg"])
def test_fixture_reference_semantics(name):
    """configs[0]/[1]: single-image decode with the reference's semantics (REF layout)."""
    data = fixture_bytes(name)
    img = JPEGImage.parse(data, layout=LAYOUT_REF)
    o = O.decode(data, layout=O.LAYOUT_REF)
    assert (img.width(), img.height()) == (o.width, o.height)
    assert img.bytes_read == KNOWN[name][1] == o.bytes_read
    assert_samples(img.rgb(), o.rgb, name)
    outs, statuses, br, coefs, _ = run_batch([data], LAYOUT_REF)
    stream = b"".join(c.astype("<i2").tobytes() for c in coefs[0])
    assert hashlib.sha256(stream).hexdigest() == KNOWN[name][0]


def test_huff_simple0_panics_like_the_reference_and_decodes_with_extension():
    data = fixture_bytes("huff_simple0.jpg")
    with pytest.raises(JPEGPanic) as e:
        JPEGImage.parse(data, layout=LAYOUT_REF)
    assert e.value.status == _ffi.PANIC_APP12_14          # mod.rs:446
    img = JPEGImage.parse(data, ext=EXT_SKIP_APPN, layout=LAYOUT_REF)
    rgb = img.rgb()
    assert (rgb[:, :8] == 0).all() and (rgb[:, 8:] == 255).all()
    assert img.bytes_read == 8


def test_2x2_chroma_spec_layout():
    """The only 4:2:0 fixture; SPEC layout (REF layout for H2V2 is covered by test_ref_layout_*)."""
    data = fixture_bytes("2x2-chroma.jpeg")
    outs, statuses, br, coefs, _ = run_batch([data], LAYOUT_SPEC)
    assert statuses[0] == 0 and br[0] == 145019
    stream = b"".join(c.astype("<i2").tobytes() for c in coefs[0])
    assert hashlib.sha256(stream).hexdigest() == "d29cc5cc8e09c5def6a31f652905c99f00e9db477b5c91d1666a01c02fd839a7"
    assert_samples(outs[0], O.decode(data, layout=O.LAYOUT_SPEC).rgb)


@pytest.mark.parametrize("sub", ["420", "422", "444", "440", "gray"])
def test_synthetic_shapes(sub):
    """Ragged, tiny and aligned sizes in one mixed batch; encoder ground truth + oracle."""
    shapes = [(64, 64), (17, 9), (250, 131), (1, 1), (333, 200), (8, 8), (128, 16), (640, 360)]
    files, gts = [], []
    for i, (w, h) in enumerate(shapes):
        f, g = synth.synth_jpeg(100 + i, w, h, sub, want_coefs=True)
        files.append(f)
        gts.append(g)
    compare_with_oracle(files, LAYOUT_SPEC, gts=gts)


@pytest.mark.parametrize("sub,ri", [("444", 1), ("444", 7), ("420", 3), ("gray", 1), ("422", 16), ("444", 80)])
def test_restart_intervals(sub, ri):
    """configs[3] semantics at small size: DRI corpus + DRI-invariance (SURVEY.md §8c pin 4)."""
    files, plain = [], []
    for i, (w, h) in enumerate([(640, 480), (250, 131), (64, 64)]):
        files.append(synth.synth_jpeg(200 + i, w, h, sub, restart_interval=ri))
        plain.append(synth.synth_jpeg(200 + i, w, h, sub, restart_interval=0))
    compare_with_oracle(files, LAYOUT_SPEC, ext=EXT_DRI)
    a = run_batch(files, LAYOUT_SPEC, EXT_DRI)
    b = run_batch(plain, LAYOUT_SPEC)
    for i in range(len(files)):
        assert all(np.array_equal(x, y) for x, y in zip(a[3][i], b[3][i]))
        assert np.array_equal(a[0][i], b[0][i])


def test_dri_rejected_without_extension():
    f = synth.synth_jpeg(1, 64, 64, "444", restart_interval=4)
    with pytest.raises(JPEGPanic) as e:
        JPEGImage.parse(f)
    assert e.value.status == _ffi.PANIC_DRI               # mod.rs:427


def test_ref_layout_equals_spec_where_the_reference_is_correct():
    files = [synth.synth_jpeg(300, 640, 480, "422"), synth.synth_jpeg(301, 320, 240, "444"),
             synth.synth_jpeg(302, 200, 96, "gray")]
    compare_with_oracle(files, LAYOUT_REF)


def test_2x2_chroma_reference_layout():
    """configs[2]: 2x2-chroma.jpeg decoded with the reference's own 4:2:0 placement (chroma tiling, spill into the
    next row, 1763 of 1786 MCUs: decoder.rs:191-192, 259-312, 347-379) through the single-image entry point."""
    data = fixture_bytes("2x2-chroma.jpeg")
    img = JPEGImage.parse(data, layout=LAYOUT_REF)
    o = O.decode(data, layout=O.LAYOUT_REF)
    assert img.bytes_read == o.bytes_read == 144537
    assert_samples(img.rgb(), o.rgb, "2x2-chroma REF")
    outs, statuses, br, coefs, _ = run_batch([data], LAYOUT_REF)
    stream = b"".join(c.astype("<i2").tobytes() for c in coefs[0])
    assert hashlib.sha256(stream).hexdigest() == "04cf33d3a2401666bcf9972782886d8e4418ac2eff240308ed2706b83f6379b0"
    spec = O.decode(data, layout=O.LAYOUT_SPEC).rgb
    assert (outs[0] != spec).any(axis=2).mean() > 0.5     # the two layouts really differ on this file


REF_SHAPES = [("420", 64, 64), ("420", 48, 32), ("420", 752, 592), ("420", 250, 131), ("420", 512, 64), ("444", 61, 45),
              ("gray", 33, 17), ("422", 100, 40), ("440", 64, 48), ("420", 16, 16), ("420", 8, 8), ("444", 7, 5),
              ("420", 1920, 1080)]


def test_ref_layout_where_it_differs_from_spec():
    """REF placement through the gather path, one mixed batch; shapes whose placement panics in the reference
    (index out of bounds) must report that panic."""
    files = [synth.synth_jpeg(700 + w + h, w, h, sub) for sub, w, h in REF_SHAPES]
    outs, statuses, br, coefs, _ = run_batch(files, LAYOUT_REF)
    for i, f in enumerate(files):
        o = O.decode(f, layout=O.LAYOUT_REF)
        if o.status != 0:
            assert statuses[i] == o.status, (REF_SHAPES[i], statuses[i], o.msg)
            continue
        assert statuses[i] == 0, (REF_SHAPES[i], _ffi.status_string(statuses[i]))
        for a, w in zip(coefs[i], o.coefs):
            assert np.array_equal(a, w)
        assert br[i] == o.bytes_read
        assert_samples(outs[i], o.rgb, str(REF_SHAPES[i]))


@pytest.mark.parametrize("q", [5, 50, 95, 100])
def test_quality_extremes(q):
    files = [synth.synth_jpeg(400 + i, 320, 240, "420", quality=q) for i in range(2)]
    compare_with_oracle(files, LAYOUT_SPEC)


def test_flat_images_degenerate_sync():
    flat = np.full((256, 256, 3), 128, np.uint8)
    files = [synth.encode(flat, "420"), synth.encode(flat, "gray"), synth.encode(np.zeros((512, 512, 3), np.uint8), "444")]
    compare_with_oracle(files, LAYOUT_SPEC)


def test_1080p_420_full_size_parity():
    """configs[2] shape at full size: one 1920x1080 4:2:0 image against the oracle and encoder ground truth."""
    f, g = synth.synth_jpeg(0, 1920, 1080, "420", want_coefs=True)
    worst = compare_with_oracle([f], LAYOUT_SPEC, gts=[g])
    print("1080p 4:2:0 max/mean |delta|:", worst)


def test_batch_is_deterministic_and_order_independent():
    files = [synth.synth_jpeg(500 + i, 320 + 16 * i, 200, "420") for i in range(6)]
    a = run_batch(files)
    b = run_batch(files[::-1])
    c = run_batch(files)
    for i in range(len(files)):
        assert np.array_equal(a[0][i], c[0][i])
        assert np.array_equal(a[0][i], b[0][len(files) - 1 - i])


def test_bad_inputs_are_reported_per_image():
    good = synth.synth_jpeg(600, 64, 64, "420")
    cut = good[:len(good) // 2]                       # truncated entropy data
    outs, statuses, br, coefs, _ = run_batch([good, cut, good])
    assert statuses[0] == 0 and statuses[2] == 0
    assert statuses[1] != 0
    assert np.array_equal(outs[0], outs[2])


def test_full_size_batch_properties():
    """Size-independent properties at batch scale (64 x 1080p): every copy of the same file decodes to
    identical bytes, bytes_read equals the scan length minus EOI, ground-truth coefficients are exact."""
    base = [synth.synth_jpeg(i, 1920, 1080, "420", want_coefs=True) for i in range(4)]
    files = [base[i % 4][0] for i in range(64)]
    b = Batch(files, layout=LAYOUT_SPEC)
    b.upload().decode()
    outs = b.download()
    statuses, br = b.results()
    assert all(s == 0 for s in statuses)
    for i in range(64):
        assert np.array_equal(outs[i], outs[i % 4])
    for i in (0, 1, 2, 3, 63):
        got = b.coefficients(i)
        for a, w in zip(got, base[i % 4][1]):
            assert np.array_equal(a, w)
    b.close()


def test_waves_through_one_batch_object():
    """a job larger than one GPU's memory runs wave after wave through one planned batch whose
    arenas are reused (jpgpu_batch_replan), outputs kept in one caller-owned device arena (set_device_output)."""
    from jpeg_rust_b200 import decode_waves
    files = [synth.synth_jpeg(800 + i, 160 + 16 * (i % 3), 120 + 8 * (i % 2), ["420", "444", "gray", "422"][i % 4]) for i in range(11)]
    files[4] = files[4][:len(files[4]) // 2]                     # one broken image in the middle of a wave
    outs, statuses, br = decode_waves(files, 4)
    ref, ref_st, ref_br = run_batch(files)[:3]
    assert statuses == ref_st and statuses[4] != 0
    for i in range(len(files)):
        if statuses[i] == 0:
            assert br[i] == ref_br[i]
            assert np.array_equal(outs[i].cpu().numpy(), ref[i]), i


def test_replan_grows_and_shrinks():
    small = [synth.synth_jpeg(820, 64, 64, "420")]
    big = [synth.synth_jpeg(821 + i, 640, 480, "420") for i in range(3)]
    b = Batch(small, layout=LAYOUT_SPEC)
    for files in (small, big, small, big):
        b.replan(files)
        b.upload().decode()
        outs = b.download()
        statuses, _ = b.results()
        assert all(s == 0 for s in statuses)
        for f, o in zip(files, outs):
            assert_samples(o, O.decode(f, layout=O.LAYOUT_SPEC).rgb)
    b.close()


@pytest.mark.parametrize("ri", [240, 16, 1])
def test_4k_444_dense_restart_intervals_full_size(ri):
    """configs[3]: synthetic 3840x2160 4:4:4 with dense restart intervals (intra-image parallelism from RSTn):
    coefficients against the encoder's ground truth and the oracle, samples within 1 LSB, and identical to the
    same image without restart markers (DRI-invariance, SURVEY.md §8c pin 4)."""
    f, g = synth.synth_jpeg(3, 3840, 2160, "444", restart_interval=ri, want_coefs=True)
    plain = synth.synth_jpeg(3, 3840, 2160, "444")
    worst = compare_with_oracle([f], LAYOUT_SPEC, ext=EXT_DRI, gts=[g])
    a = run_batch([f], LAYOUT_SPEC, EXT_DRI)
    b = run_batch([plain], LAYOUT_SPEC)
    assert np.array_equal(a[0][0], b[0][0])
    print("4K 4:4:4 Ri =", ri, "max/mean |delta|:", worst)


def test_corrupted_scans_never_derail_the_batch():
    """Random damage inside the entropy-coded data (bit flips, with and without restart markers): every image gets a
    status, nothing crashes or hangs, undamaged images of the same batch decode exactly as they do alone."""
    rng = np.random.default_rng(11)
    good = [synth.synth_jpeg(950 + i, 320, 240, s, restart_interval=ri) for i, (s, ri) in enumerate([("420", 0), ("444", 5), ("gray", 0), ("422", 2)])]
    files = []
    for k in range(96):
        f = bytearray(good[k % 4])
        sos = bytes(f).index(b"\xff\xda") + 14
        for _ in range(1 + k % 5):
            pos = int(rng.integers(sos, len(f) - 2))
            f[pos] ^= 1 << int(rng.integers(0, 8))
        if k % 7 == 0:
            f = f[:int(rng.integers(sos + 8, len(f)))]
        files.append(bytes(f))
    files += good
    outs, statuses, br, coefs, _ = run_batch(files, LAYOUT_SPEC, EXT_DRI)
    alone = run_batch(good, LAYOUT_SPEC, EXT_DRI)
    for i in range(4):
        assert statuses[96 + i] == 0
        assert np.array_equal(outs[96 + i], alone[0][i])
    agree = 0
    for i in range(96):
        o = O.decode(files[i], layout=O.LAYOUT_SPEC, ext=EXT_DRI)
        if statuses[i] == 0 and o.status == 0:
            assert all(np.array_equal(a, b) for a, b in zip(coefs[i], o.coefs)), i
            agree += 1
    print("damaged images still decodable by both:", agree, "of 96")


def test_image_specific_huffman_tables_and_16bit_dqt():
    """Files the reference accepts but the Annex-K corpus does not cover: optimised (T.81 K.2) Huffman tables, a
    different set per image (many LUT sets in one batch), and 16-bit quantisation tables (mod.rs:245-256)."""
    files, gts = [], []
    for i, (sub, q, opt, wide) in enumerate([("420", 85, True, False), ("444", 30, True, False), ("gray", 95, True, False),
                                             ("422", 8, False, True), ("420", 3, True, True), ("440", 60, True, False)]):
        f, g = synth.synth_jpeg(1000 + i, 200 + 8 * i, 136, sub, quality=q, want_coefs=True, optimize=opt, dqt16=wide)
        files.append(f)
        gts.append(g)
    outs, statuses, br, coefs, _ = run_batch(files)
    n_oracle = 0
    for i, (f, g) in enumerate(zip(files, gts)):
        assert statuses[i] == 0, _ffi.status_string(statuses[i])
        assert all(np.array_equal(a, b) for a, b in zip(coefs[i], g))
        o = O.decode(f, layout=O.LAYOUT_SPEC)
        if o.status == 0:               # within the reference's subset (no 1-bit code in any table)
            n_oracle += 1
            assert all(np.array_equal(a, b) for a, b in zip(coefs[i], o.coefs)) and br[i] == o.bytes_read
            assert_samples(outs[i], o.rgb, f"image {i}")
        else:
            assert o.status in (9, 10)  # huffman.rs:156/162: a table holds a 1-bit code
    assert n_oracle >= 3


def test_one_bit_huffman_codes_decode_although_the_reference_cannot():
    """A table with a single symbol gets a 1-bit code; the reference cannot decode those (huffman.rs:61, 212 -> the
    oracle reports the panic).  The GPU path decodes them: coefficients equal the encoder's, pixels equal those of the
    same image coded with the Annex-K tables."""
    flat = np.full((72, 104, 3), 77, np.uint8)
    grad = np.tile(np.arange(104, dtype=np.uint8)[None, :, None] * 2, (72, 1, 3))
    for img, sub in ((flat, "420"), (flat, "gray"), (grad, "444")):
        f, g = synth.encode(img, sub, want_coefs=True, optimize=True)
        plain = synth.encode(img, sub)
        assert O.decode(f, layout=1).status != 0 or sub == "444"
        outs, statuses, br, coefs, _ = run_batch([f, plain])
        assert statuses == [0, 0]
        assert all(np.array_equal(a, b) for a, b in zip(coefs[0], g))
        assert np.array_equal(outs[0], outs[1])


def test_planar_output_holds_the_same_samples():
    """OUT_RGB_PLANAR is the interleaved output (decoder.rs:317-331) transposed to three planes —
    every fused kernel variant, ragged widths (byte path of the copy-out) and the gather path (REF placement)."""
    from jpeg_rust_b200 import OUT_RGB_INTERLEAVED, OUT_RGB_PLANAR
    files = [synth.synth_jpeg(500, 640, 480, "420"), synth.synth_jpeg(501, 320, 240, "422"), synth.synth_jpeg(502, 256, 64, "444"),
             synth.synth_jpeg(503, 128, 200, "440"), synth.synth_jpeg(504, 512, 96, "gray"), synth.synth_jpeg(505, 251, 131, "420"),
             synth.synth_jpeg(506, 77, 50, "444"), fixture_bytes("lena.jpeg")]
    for layout in (LAYOUT_SPEC, LAYOUT_REF):
        b = Batch(files, layout=layout)
        b.upload().decode()
        inter = b.download()
        b.set_output_format(OUT_RGB_PLANAR).idct()
        planar = b.download()
        st, _ = b.results()
        assert all(s == 0 for s in st), st
        for i in range(len(files)):
            h, w = inter[i].shape[:2]
            assert planar[i].shape == (3, h, w)
            assert np.array_equal(planar[i], inter[i].transpose(2, 0, 1)), f"layout {layout} image {i}"
        t = b.device_tensor(0)
        assert tuple(t.shape) == planar[0].shape and np.array_equal(t.cpu().numpy(), planar[0])
        b.set_output_format(OUT_RGB_INTERLEAVED).idct()
        again = b.download()
        b.ctx.sync()
        assert all(np.array_equal(x, y) for x, y in zip(again, inter))
        b.close()


@pytest.mark.parametrize("mode", ["0", "1"])
def test_restart_interval_modes_agree(mode, monkeypatch):
    """Images with restart intervals decode either like any other image (look-back synchronisation, mode 0) or, when
    the intervals are short against a subsequence, with every decode thread starting at an interval boundary and no
    synchronisation pass at all (mode 1; the planner picks per image).  Both must give the oracle's coefficients."""
    monkeypatch.setenv("JPGPU_INTERVAL_MODE", mode)
    files = [synth.synth_jpeg(700, 640, 480, "420", restart_interval=2), synth.synth_jpeg(701, 333, 200, "444", restart_interval=40),
             synth.synth_jpeg(702, 512, 512, "gray", restart_interval=64), synth.synth_jpeg(703, 320, 240, "422", restart_interval=0)]
    compare_with_oracle(files, LAYOUT_SPEC, ext=EXT_DRI)


def test_scans_handed_over_in_device_memory():
    """jpgpu_batch_set_device_scans: the entropy-coded bytes already lie in device memory, image i at base + offsets[i]
    at any alignment (here: whole files packed back to back with odd gaps, the last one ending with the buffer).
    Same output as the host upload."""
    import torch
    files = [synth.synth_jpeg(900 + i, 200 + 24 * i, 96 + 8 * i, ["420", "444", "gray", "422", "440"][i % 5], restart_interval=(3 if i == 2 else 0))
             for i in range(7)]
    ref = run_batch(files, LAYOUT_SPEC, EXT_DRI)
    b = Batch(files, layout=LAYOUT_SPEC, ext=EXT_DRI)
    blob, offsets, pos = bytearray(), [], 0
    for i, f in enumerate(files):
        gap = (i * 5 + 1) % 7          # odd alignments
        blob += b"\xa5" * gap
        pos += gap
        scan_at = len(f) - b.descs[i].scan_len          # the scan runs to the end of the file (mod.rs:371-385)
        offsets.append(pos + scan_at)
        blob += f
        pos += len(f)
    dev = torch.from_numpy(np.frombuffer(bytes(blob), dtype=np.uint8).copy()).cuda()
    assert dev.numel() == offsets[-1] + b.descs[len(files) - 1].scan_len
    b.set_device_scans(dev.data_ptr(), offsets).decode()
    outs = b.download()
    st, br = b.results()
    assert st == ref[1] and br == ref[2]
    for i in range(len(files)):
        assert np.array_equal(outs[i], ref[0][i]), i
        assert all(np.array_equal(x, y) for x, y in zip(b.coefficients(i), ref[3][i]))
    b.close()


def test_large_image_32bit_offsets():
    """A 16384 x 8192 4:2:0 image (134 Mpixel, 201 M coefficients, 403 MB of RGB): every per-image offset of the
    kernels is exercised far beyond 2^24.  Coefficients against the encoder's own (the oracle's O(N^4) IDCT would
    take minutes here), samples against the same image carrying restart markers (DRI-invariance: two different
    decode paths — look-back synchronisation vs interval starts — must agree byte for byte), planar against interleaved."""
    from jpeg_rust_b200 import OUT_RGB_PLANAR
    w, h = 16384, 8192
    f, gt = synth.synth_jpeg(77, w, h, "420", want_coefs=True)
    fr = synth.synth_jpeg(77, w, h, "420", restart_interval=64)
    b = Batch([f, fr], layout=LAYOUT_SPEC, ext=EXT_DRI)
    b.upload().decode()
    st, _ = b.results()
    assert st == [0, 0], st
    for i in range(2):
        got = b.coefficients(i)
        for a, c in zip(got, gt):
            assert np.array_equal(a[:len(c)], c[:len(a)])
    t0, t1 = b.device_tensor(0), b.device_tensor(1)
    import torch
    assert torch.equal(t0, t1)
    inter = t0.clone()
    b.set_output_format(OUT_RGB_PLANAR).idct()
    b.ctx.sync()
    assert torch.equal(b.device_tensor(0), inter.permute(2, 0, 1))
    b.close()


@pytest.mark.parametrize("env", [{}, {"JPGPU_SUBSEQ_BITS": "4096", "JPGPU_LOOKBACK_BITS": "256", "JPGPU_WRITE_PARTS": "2"},
                                 {"JPGPU_SUBSEQ_BITS": "8192", "JPGPU_LOOKBACK_BITS": "1024", "JPGPU_INTERVAL_MODE": "0"}])
def test_randomised_batches(env, monkeypatch):
    """Fixed-seed mixed batches (sampling, size, quality, restart interval all over the place) under different planner
    settings: coefficients against the encoder's own for every image, statuses clean."""
    import random
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = random.Random(4242 + len(env))
    files, gts = [], []
    for it in range(120):
        sub = rng.choice(["420", "420", "422", "444", "440", "gray"])
        w = rng.choice([rng.randint(8, 300), rng.randint(300, 1400), rng.choice([512, 1024, 2048])])
        h = rng.choice([rng.randint(8, 300), rng.randint(300, 1000), rng.choice([512, 1024])])
        ri = rng.choice([0, 0, 1, 2, 3, 5, 8, 16, 33, 64, 100, 256, 1000])
        f, g = synth.synth_jpeg(rng.randint(0, 10 ** 6), w, h, sub, quality=rng.choice([30, 60, 85, 95]), restart_interval=ri, want_coefs=True)
        files.append(f)
        gts.append(g)
    b = Batch(files, layout=LAYOUT_SPEC, ext=EXT_DRI)
    b.upload().decode()
    st, _ = b.results()
    assert all(s == 0 for s in st), [(i, s) for i, s in enumerate(st) if s]
    for i in range(len(files)):
        for a, g in zip(b.coefficients(i), gts[i]):
            assert np.array_equal(a[:len(g)], g[:len(a)]), i
    b.close()


# ---------------------------------------------------------------------------------------------------------------------
# The path bench.py times: jpgpu_batch_decode of >= 192 images = three image groups on three auxiliary streams with the
# planner's own subsequence / look-back choice (the largest batch under -m gpu used to be 120
# images = one group on one stream).

def _grouped_batch_check(files, gts, expect_groups, oracle_every):
    from jpeg_rust_b200 import plan_info
    info = plan_info(files, layout=LAYOUT_SPEC)
    assert info["groups"] == expect_groups, info
    b = Batch(files, layout=LAYOUT_SPEC)
    b.upload().decode()
    outs = b.download()
    statuses, br = b.results()
    assert all(s == 0 for s in statuses)
    for i, f in enumerate(files):
        got = b.coefficients(i)
        for a, w in zip(got, gts[i]):
            assert np.array_equal(a[:len(w)], w[:len(a)]), f"image {i}: coefficients differ from the encoder's"
    worst = (0, 0.0)
    for i in range(0, len(files), oracle_every):
        o = O.decode(files[i], layout=LAYOUT_SPEC)
        assert br[i] == o.bytes_read
        m = assert_samples(outs[i], o.rgb, f"image {i}")
        worst = (max(worst[0], m[0]), max(worst[1], m[1]))
    b.close()
    return info, worst


def test_three_group_decode_of_a_mixed_corpus():
    """208 images of mixed size and sampling (>= 192: three groups, default planner): every image's coefficients against
    the encoder's, every 13th image's samples against the oracle."""
    subs = ["420", "444", "422", "gray", "440"]
    made = [synth.synth_jpeg(3000 + i, 96 + 16 * (i % 11), 64 + 8 * (i % 7), subs[i % 5], want_coefs=True) for i in range(208)]
    info, worst = _grouped_batch_check([m[0] for m in made], [m[1] for m in made], 3, 13)
    print("208 mixed images, 3 groups:", info, "max/mean |delta|:", worst)


def test_three_group_decode_of_1024_small_images():
    """1024 small 4:2:0 images through the three-group path (the benchmark's image count)."""
    made = [synth.synth_jpeg(4000 + i, 64 + 16 * (i % 4), 48 + 16 * (i % 3), "420", want_coefs=True) for i in range(64)]
    files = [made[i % 64][0] for i in range(1024)]
    gts = [made[i % 64][1] for i in range(1024)]
    info, worst = _grouped_batch_check(files, gts, 3, 97)
    print("1024 small images, 3 groups:", info, "max/mean |delta|:", worst)


def test_benchmark_planner_settings_on_full_size_images():
    """256 x 1080p 4:2:0 (16 distinct): three groups, 4096-bit subsequences - and the same images with the planner
    forced to what it picks for the 1024-image benchmark batch (8192-bit subsequences, 1024 bits of look-back, three
    groups).  All coefficients against the encoder's; four images against the oracle."""
    import os
    made = [synth.synth_jpeg(5000 + i, 1920, 1080, "420", want_coefs=True) for i in range(16)]
    files = [made[i % 16][0] for i in range(256)]
    gts = [made[i % 16][1] for i in range(256)]
    info, worst = _grouped_batch_check(files, gts, 3, 64)
    print("256 x 1080p default plan:", info, worst)
    keep = {k: os.environ.get(k) for k in ("JPGPU_SUBSEQ_BITS", "JPGPU_LOOKBACK_BITS", "JPGPU_GROUPS")}
    os.environ.update(JPGPU_SUBSEQ_BITS="8192", JPGPU_LOOKBACK_BITS="1024", JPGPU_GROUPS="3")
    try:
        info, worst = _grouped_batch_check(files, gts, 3, 64)
        assert info["sub_bits"] == 8192 and info["lookback_bits"] == 1024
        print("256 x 1080p benchmark plan:", info, worst)
    finally:
        for k, v in keep.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_jpegdecoder_builder_decodes_on_the_gpu():
    """the reference's builder sequence (mod.rs:388-415 / decoder.rs:55-162) on the Python mirror -
    JPEGDecoder.new(raw).frame_header(..).scan_header(..).dimensions(..), the three table setters, .decode() - must give
    what JPEGImage.parse gives for the same file, and what the oracle gives."""
    from jpeg_rust_b200.jpeg import (FrameComponentHeader, FrameHeader, HuffmanTable, JPEGDecoder, ScanComponentHeader,
                                     ScanHeader, parse_descriptor)
    for name, ext in (("lena.jpeg", EXT_NONE), ("lena-bw.jpeg", EXT_NONE), ("2x2-chroma.jpeg", EXT_NONE)):
        data = fixture_bytes(name)
        st, d, buf = parse_descriptor(data, ext, LAYOUT_REF)
        assert st == 0
        raw = np.ctypeslib.as_array((_ffi.C.c_uint8 * d.scan_len).from_address(d.scan)).copy()
        fh = FrameHeader(8, d.height, d.width, d.ncomp,
                         [FrameComponentHeader(d.comp[c].id, d.comp[c].h, d.comp[c].v, d.comp[c].tq) for c in range(d.ncomp)])
        sh = ScanHeader(d.ncomp, [ScanComponentHeader(d.comp[c].id, d.comp[c].td, d.comp[c].ta) for c in range(d.ncomp)])
        dec = JPEGDecoder.new(raw, layout=LAYOUT_REF).frame_header(fh).scan_header(sh).dimensions((d.width, d.height))
        for t in range(4):
            if d.ac_present[t]:
                dec.huffman_ac_tables(t, HuffmanTable.from_size_data_tables(bytes(d.ac_bits[t]), bytes(d.ac_vals[t][:d.ac_nvals[t]])))
            if d.dc_present[t]:
                dec.huffman_dc_tables(t, HuffmanTable.from_size_data_tables(bytes(d.dc_bits[t]), bytes(d.dc_vals[t][:d.dc_nvals[t]])))
            if d.qt_present[t]:
                dec.quantization_table(t, list(d.qt[t]))
        pixels, bytes_read = dec.decode()
        img = JPEGImage.parse(data, ext=ext, layout=LAYOUT_REF)
        o = O.decode(data, layout=LAYOUT_REF, ext=ext)
        assert pixels.shape == (d.width * d.height, 3)
        assert np.array_equal(pixels, img.image_data()) and bytes_read == img.bytes_read == o.bytes_read
        assert_samples(pixels.reshape(d.height, d.width, 3), o.rgb, name)


def test_position_saturates_on_a_flooded_scan():
    """see tests/test_sim_parity.py; 8.5 MB and 17 MB of zeros behind an 8x8 image, next to good
    images in the same batch.  No wild store, the declared block decodes, the neighbours are untouched."""
    good = synth.synth_jpeg(601, 64, 64, "420")
    files = [good, synth.crafted_flood_jpeg(8_500_000), good, synth.crafted_flood_jpeg(17_000_000), good]
    outs, statuses, br, coefs, _ = run_batch(files)
    assert statuses == [0, 0, 0, 0, 0]
    assert br[1] == 1 and br[3] == 1 and (outs[1] == 128).all() and (outs[3] == 128).all()
    assert np.array_equal(outs[0], outs[2]) and np.array_equal(outs[0], outs[4])
    assert_samples(outs[0], O.decode(good, layout=LAYOUT_SPEC).rgb)


def test_truncated_image_reads_as_zeros_past_its_data():
    """Arenas are reused wave after wave and never cleared: an image whose data ends early must not show the previous
    wave's coefficients in the blocks it never reached.  Decode a full image, then - same batch
    object, same shape - its truncated twin: the tail is mid-gray (all-zero blocks), twice the same bytes."""
    full = synth.synth_jpeg(7001, 256, 256, "420")
    other = synth.synth_jpeg(7002, 256, 256, "420")
    cut = other[:len(other) * 2 // 3]
    b = Batch([full], layout=LAYOUT_SPEC)
    b.upload().decode()
    ref_full = b.download()[0].copy()
    b.ctx.sync()
    runs = []
    for _ in range(2):
        b.replan([cut])
        b.upload().decode()
        out = b.download()[0].copy()
        st, _ = b.results()
        assert st[0] == _ffi.ERR_TRUNCATED
        runs.append(out)
        b.replan([full])
        b.upload().decode()
        b.results()
    assert np.array_equal(runs[0], runs[1])
    assert (runs[0][-16:] == 128).all(), "rows past the end of the data must come from zero blocks"
    assert not np.array_equal(runs[0][-16:], ref_full[-16:])
    b.close()


def test_replan_under_an_external_output_arena():
    """a replan to a larger plan while a caller-owned output arena is set falls back to the
    batch's own arena, sized for the new plan; restoring the own arena after that never leaves it too small."""
    import torch
    small = [synth.synth_jpeg(7100, 64, 64, "420")]
    big = [synth.synth_jpeg(7101 + i, 640, 480, "420") for i in range(3)]
    b = Batch(small, layout=LAYOUT_SPEC)
    arena = torch.empty(b.output_bytes() + 256, dtype=torch.uint8, device="cuda")
    base = (arena.data_ptr() + 255) // 256 * 256
    b.set_device_output(base, b.output_bytes())
    b.upload().decode()
    assert b.device_rgb(0)[0] == base
    b.replan(big)                      # does not fit the caller's arena any more
    assert b.device_rgb(0)[0] != base
    b.upload().decode()
    outs = b.download()
    statuses, _ = b.results()
    assert all(s == 0 for s in statuses)
    b.set_device_output(None, 0)
    b.upload().decode()
    outs2 = b.download()
    b.results()
    for a, c, f in zip(outs, outs2, big):
        assert np.array_equal(a, c)
        assert_samples(a, O.decode(f, layout=LAYOUT_SPEC).rgb)
    b.close()


def test_pipeline_host_to_host_matches_the_batch_path():
    """jpgpu_pipeline_* (SURVEY 8(f) row 3): files in one pinned buffer in, pixels in one pinned buffer out, chunks
    alternating between two stream sets, every transfer a single copy - byte-identical to upload/decode/download."""
    from jpeg_rust_b200 import Pipeline
    subs = ["420", "444", "422", "gray"]
    files = [synth.synth_jpeg(7200 + i, 80 + 16 * (i % 5), 56 + 8 * (i % 3), subs[i % 4]) for i in range(23)]
    files[7] = files[7][:len(files[7]) // 2]
    ref, ref_st, ref_br = run_batch(files)[:3]
    p = Pipeline(files, chunk=5)
    for _ in range(2):            # a second run on the same plans gives the same bytes
        p.run().sync()
        st, br = p.results()
        assert st == ref_st and st[7] != 0
        for i in range(len(files)):
            if st[i] == 0:
                assert br[i] == ref_br[i]
                assert np.array_equal(p.image(i), ref[i]), i
    assert p.elapsed_ms() > 0 and p.launch_count() > 0
    p.close()


def test_single_copy_transfers():
    """jpgpu_batch_upload_from / jpgpu_batch_download_contiguous: one copy each way instead of one per image."""
    from jpeg_rust_b200 import pack_files, parse_packed
    files = [synth.synth_jpeg(7300 + i, 64 + 16 * (i % 4), 64, "420") for i in range(9)]
    ref = run_batch(files)[0]
    buf, offs, owner = pack_files(files)
    descs, pst = parse_packed(buf, offs, [len(f) for f in files])
    assert not any(pst)
    b = Batch(descs=descs, keepalive=owner)
    out = np.zeros(b.output_bytes(), np.uint8)
    b.upload_from(buf).decode().download_contiguous(out)
    st, _ = b.results()
    assert all(s == 0 for s in st)
    for i in range(len(files)):
        off, nb = b.rgb_offset(i)
        assert np.array_equal(out[off:off + nb].reshape(ref[i].shape), ref[i])
    b.close()


def test_multi_device_handle_matches_one_device():
    """jpgpu_multi_*: one process, one context + stream set + worker thread per device, contiguous ranges balanced by
    scan bytes, no collective.  On every device count available (1, and 2+ when the box has them) the result is
    byte-identical to the single-device batch."""
    import torch
    from jpeg_rust_b200 import MultiDevice
    subs = ["420", "444", "422", "gray"]
    made = [synth.synth_jpeg(7400 + i, 96 + 16 * (i % 6), 64 + 8 * (i % 4), subs[i % 4], want_coefs=True) for i in range(37)]
    files = [m[0] for m in made]
    ref, ref_st, ref_br = run_batch(files)[:3]
    ndev = torch.cuda.device_count()
    for devices in ([0], list(range(min(ndev, 2))), list(range(ndev))):
        m = MultiDevice(devices)
        m.plan(files).upload().decode()
        outs = m.download()
        st, br = m.results()
        rngs = m.ranges()
        assert rngs[0][1] == 0 and sum(r[2] for r in rngs) == len(files) and [r[0] for r in rngs] == devices
        assert st == ref_st and br == ref_br
        for i in range(len(files)):
            assert np.array_equal(outs[i], ref[i]), (devices, i)
        for i in (0, 18, 36):
            for a, w in zip(m.coefficients(i), made[i][1]):
                assert np.array_equal(a[:len(w)], w[:len(a)])
        ms = m.time_decode(2)
        assert len(ms) == len(devices) and all(x > 0 for x in ms)
        o1, s1, b1 = m.decode_batch(files[:1])                # fewer images than devices: the other ranges are empty
        assert s1 == ref_st[:1] and b1 == ref_br[:1] and np.array_equal(o1[0], ref[0])
        outs2, st2, br2 = m.decode_batch(files[:11])          # the one-call form, replanning the same handle
        assert st2 == ref_st[:11] and br2 == ref_br[:11]
        for i in range(11):
            assert np.array_equal(outs2[i], ref[i])
        m.close()


def test_c_program_decodes_through_the_abi():
    """tests/c/abi_smoke: a C program that knows nothing but include/jpgpu.h decodes the fixtures through
    jpgpu_decode_file, the batch calls and the host pipeline; its output hash equals the Python binding's, its
    dimensions and bytes_read the oracle's."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "c", "abi_smoke")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", root, "tests/c/abi_smoke"])

    def fnv1a(buf):
        h = 1469598103934665603
        for chunk in np.frombuffer(buf, np.uint8).tolist():
            h = ((h ^ chunk) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h
    for name, ext in (("lena.jpeg", 0), ("lena-bw.jpeg", 0), ("huff_simple0.jpg", 1)):
        path = os.path.join(root, "tests", "golden", "fixtures", name)
        r = subprocess.run([exe, path, str(ext), "0"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        w, h, br, digest, batch_same, pipe_same = r.stdout.split()
        o = O.decode(fixture_bytes(name), layout=LAYOUT_REF, ext=ext)
        img = JPEGImage.parse(fixture_bytes(name), ext=ext, layout=LAYOUT_REF)
        assert (int(w), int(h), int(br)) == (o.width, o.height, o.bytes_read)
        assert int(digest, 16) == fnv1a(img.image_data().tobytes())
        assert batch_same == "1" and pipe_same == "1"


def test_f32_normalised_planar_output():
    """JPGPU_OUT_F32_PLANAR - the same 8-bit samples as three float32 planes, sample * scale + bias,
    converted inside the IDCT/colour kernel.  Default scale 1/255; an ImageNet-style normalisation; every sampling mode,
    ragged sizes (scalar copy-out) and the gather path (REF layout)."""
    import torch
    subs = ["420", "444", "422", "gray", "440"]
    files = [synth.synth_jpeg(7500 + i, 128 + 16 * i, 96 + 8 * i, subs[i % 5]) for i in range(5)]
    files += [synth.synth_jpeg(7510, 131, 77, "420"), synth.synth_jpeg(7511, 1920, 1080, "420")]
    for layout in (LAYOUT_SPEC, LAYOUT_REF):
        ref, ref_st = run_batch(files, layout)[:2]
        b = Batch(files, layout=layout)
        b.set_output_format(_ffi.OUT_F32_PLANAR)
        b.upload().decode()
        outs = b.download()
        st, _ = b.results()
        assert st == ref_st and sum(s == 0 for s in st) >= 5    # REF placement panics on some shapes (decoder.rs:372), like the reference
        ok = [i for i in range(len(files)) if st[i] == 0]
        for i in ok:
            assert outs[i].dtype == np.float32 and outs[i].shape == (3,) + ref[i].shape[:2]
            want = ref[i].transpose(2, 0, 1).astype(np.float32) * np.float32(1.0 / 255.0)
            assert np.allclose(outs[i], want, rtol=0, atol=1e-6), i
            t = b.device_tensor(i)
            assert t.dtype == torch.float32 and np.array_equal(t.cpu().numpy(), outs[i])
        files_ok = ok
        mean, std = np.array([0.485, 0.456, 0.406], np.float32), np.array([0.229, 0.224, 0.225], np.float32)
        b.set_normalisation(1.0 / (255.0 * std), -mean / std)
        b.decode()
        outs = b.download()
        b.results()
        for i in files_ok:
            want = (ref[i].transpose(2, 0, 1).astype(np.float32) / 255.0 - mean[:, None, None]) / std[:, None, None]
            assert np.allclose(outs[i], want, rtol=0, atol=2e-5), i
        b.set_output_format(_ffi.OUT_RGB_INTERLEAVED)      # and back: the u8 path is untouched
        b.decode()
        outs = b.download()
        b.results()
        for i in files_ok:
            assert np.array_equal(outs[i], ref[i])
        b.close()


def test_ppm_writer_matches_main_rs(tmp_path):
    """main.rs:34-39: "P3\\n{w} {h}\\n255\\n" then one "r g b\\n" line per pixel, row-major.  JPEGImage.write_ppm against
    a restatement of those lines fed with the oracle's pixels (exact where the GPU and the oracle agree exactly)."""
    for name in ("lena-bw.jpeg", "lena.jpeg"):
        data = fixture_bytes(name)
        img = JPEGImage.parse(data, layout=LAYOUT_REF)
        path = tmp_path / (name + ".ppm")
        img.write_ppm(str(path))
        text = path.read_text()
        lines = text.split("\n")
        assert lines[0] == "P3" and lines[1] == f"{img.width()} {img.height()}" and lines[2] == "255" and lines[-1] == ""
        assert len(lines) == 3 + img.width() * img.height() + 1
        o = O.decode(data, layout=LAYOUT_REF)
        want = "".join(f"{r} {g} {b}\n" for r, g, b in o.rgb.reshape(-1, 3).tolist())     # main.rs:36-38
        got_px = np.array([ln.split() for ln in lines[3:-1]], dtype=np.int16)
        assert np.abs(got_px - o.rgb.reshape(-1, 3).astype(np.int16)).max() <= 1
        if np.array_equal(img.image_data().reshape(-1, 3), o.rgb.reshape(-1, 3)):
            assert text == f"P3\n{o.width} {o.height}\n255\n" + want


@pytest.mark.parametrize("multi", ["0", "1"])
def test_sync_pass_single_and_multi_symbol(multi, monkeypatch):
    """Both forms of the synchronisation pass (JPGPU_SYNC_MULTI=0: sync_kernel, symbol by symbol; 1: sync_multi_kernel,
    multi-symbol tables) on mixed images incl. restart intervals, optimised tables and a dense image: coefficients
    against the oracle's, identical bytes_read."""
    monkeypatch.setenv("JPGPU_SYNC_MULTI", multi)
    files = [synth.synth_jpeg(7600 + i, 640 + 32 * i, 400, s) for i, s in enumerate(["420", "444", "422", "gray", "440"])]
    files += [synth.synth_jpeg(7610, 512, 384, "420", optimize=True), synth.synth_jpeg(7611, 800, 600, "420", quality=97, noise_sigma=25.0),
              fixture_bytes("lena.jpeg"), fixture_bytes("2x2-chroma.jpeg")]
    compare_with_oracle(files, LAYOUT_SPEC)
    dri = [synth.synth_jpeg(7620 + i, 640, 480, "420", restart_interval=ri) for i, ri in enumerate([40, 400])]
    compare_with_oracle(dri, LAYOUT_SPEC, EXT_DRI)


def test_repeated_decodes_replay_a_cuda_graph():
    """jpgpu_batch_decode: first call kernel by kernel, second call captured (all groups, forks and joins over the auxiliary
    streams), later calls one graph launch.  Same bytes every time; a changed output format or a replan drops the graph."""
    made = [synth.synth_jpeg(7700 + i, 160 + 16 * (i % 5), 120, ["420", "444", "gray"][i % 3], want_coefs=True) for i in range(200)]
    files = [m[0] for m in made]
    b = Batch(files, layout=LAYOUT_SPEC)       # 200 images: three groups on auxiliary streams
    b.upload()
    runs = []
    for k in range(4):
        b.decode()
        runs.append(b.download())
        st, _ = b.results()
        assert all(s == 0 for s in st)
    for k in range(1, 4):
        for a, c in zip(runs[0], runs[k]):
            assert np.array_equal(a, c)
    for i in (0, 77, 199):
        for a, w in zip(b.coefficients(i), made[i][1]):
            assert np.array_equal(a[:len(w)], w[:len(a)])
    n0 = b.launch_count()
    b.decode()
    assert b.launch_count() - n0 >= 3 * 7            # a replay still accounts for the kernels it runs
    b.set_output_format(_ffi.OUT_RGB_PLANAR)
    for k in range(3):
        b.decode()
        outs = b.download()
        b.results()
        for a, c in zip(runs[0], outs):
            assert np.array_equal(a.transpose(2, 0, 1), c)
    b.replan(files[:50])
    b.upload()
    for k in range(3):
        b.decode()
        outs = b.download()
        b.results()
        for a, c in zip(runs[0][:50], outs):
            assert np.array_equal(a.transpose(2, 0, 1), c)
    b.close()


from fancy_ref import fancy_reference as _fancy_reference  # noqa: E402


def test_fancy_upsampling_layout():
    """SURVEY 8(f) row 4 (a feature the reference lacks, behind JPGPU_LAYOUT_SPEC_FANCY): sub-sampled chroma interpolated
    with libjpeg's triangle filter instead of replicated.  Against the numpy restatement above (+-1: fma contraction), and
    as a sanity check against libjpeg itself (PIL), which this layout must approach much closer than box replication does."""
    import io
    from PIL import Image
    from jpeg_rust_b200 import LAYOUT_SPEC_FANCY
    cases = [("420", 640, 480, 0), ("422", 333, 217, 0), ("440", 320, 200, 0), ("420", 131, 77, 5), ("420", 1920, 1080, 0),
             ("444", 200, 100, 0), ("gray", 200, 100, 0)]
    files = [synth.synth_jpeg(7800 + i, w, h, s, restart_interval=ri) for i, (s, w, h, ri) in enumerate(cases)]
    files += [fixture_bytes("lena.jpeg"), fixture_bytes("2x2-chroma.jpeg")]
    fancy, st, br = run_batch(files, LAYOUT_SPEC_FANCY, EXT_DRI)[:3]
    box, st_box, br_box = run_batch(files, LAYOUT_SPEC, EXT_DRI)[:3]
    assert st == st_box == [0] * len(files) and br == br_box
    for i, f in enumerate(files):
        want = _fancy_reference(f, EXT_DRI)
        d = np.abs(fancy[i].astype(np.int16) - want.astype(np.int16))
        assert d.max() <= 1 and d.mean() < 0.01, (i, d.max(), d.mean())
        pil = np.asarray(Image.open(io.BytesIO(f)).convert("RGB")).astype(np.int16)
        e_fancy = np.abs(fancy[i].astype(np.int16) - pil).mean()
        e_box = np.abs(box[i].astype(np.int16) - pil).mean()
        if i < 5 or i >= 7:      # sub-sampled files: interpolation is what libjpeg does
            assert e_fancy < 0.8 and e_fancy < e_box, (i, e_fancy, e_box)
        else:                    # nothing to interpolate: identical to the SPEC layout
            assert np.array_equal(fancy[i], box[i])
    # every output format goes through the same kernel
    b = Batch(files[:2], layout=LAYOUT_SPEC_FANCY)
    b.set_output_format(_ffi.OUT_RGB_PLANAR)
    b.upload().decode()
    planar = b.download()
    b.results()
    b.set_output_format(_ffi.OUT_F32_PLANAR)
    b.decode()
    f32 = b.download()
    b.results()
    for i in range(2):
        assert np.array_equal(planar[i], fancy[i].transpose(2, 0, 1))
        assert np.allclose(f32[i], fancy[i].transpose(2, 0, 1).astype(np.float32) / 255.0, atol=1e-6)
    b.close()


@pytest.mark.parametrize("layout_name", ["SPEC", "SPEC_FANCY"])
def test_non_interleaved_scans(layout_name):
    """SURVEY 8(f) row 4 (a feature the reference lacks: it returns after the first scan, mod.rs:416-417): files with one
    non-interleaved scan per component (jpgpu_parse_scans).  Every scan is entropy-decoded as the one-component image it
    is; the compose path puts the frame together.  The same coefficients coded as one interleaved scan must give the same
    pixels (+-1: another IDCT instantiation), the coefficients of every scan equal the encoder's, and libjpeg agrees."""
    import io
    from PIL import Image
    from jpeg_rust_b200 import EXT_MULTISCAN, LAYOUT_SPEC_FANCY, decode_scans
    layout = LAYOUT_SPEC if layout_name == "SPEC" else LAYOUT_SPEC_FANCY
    cases = [("420", 640, 480, 0), ("422", 333, 217, 0), ("444", 200, 100, 0), ("420", 131, 77, 5), ("440", 96, 160, 0),
             ("420", 1920, 1080, 0), ("420", 64, 64, 1)]
    planar, gts, inter = [], [], []
    for i, (s, w, h, ri) in enumerate(cases):
        f, g = synth.synth_jpeg(7900 + i, w, h, s, restart_interval=ri, planar_scans=True, want_coefs=True)
        planar.append(f); gts.append(g)
        inter.append(synth.synth_jpeg(7900 + i, w, h, s, restart_interval=ri))
    inter.append(synth.synth_jpeg(7950, 320, 240, "420"))          # an ordinary file in the same call
    outs, st, coefs = decode_scans(planar + [inter[-1]], ext=EXT_DRI, layout=layout, want_coefs=True)
    assert st == [0] * (len(cases) + 1)
    ref = run_batch(inter, layout, EXT_DRI)[0]
    for i in range(len(cases)):
        for a, w in zip(coefs[i], gts[i]):
            assert np.array_equal(a, w), f"case {i}: coefficients of a scan differ from the encoder's"
        d = np.abs(outs[i].astype(np.int16) - ref[i].astype(np.int16))
        assert d.max() <= 1 and d.mean() < 0.01, (i, d.max(), d.mean())
        pil = np.asarray(Image.open(io.BytesIO(planar[i])).convert("RGB")).astype(np.int16)
        # libjpeg itself (integer IDCT and colour conversion, interpolated chroma): comparable where the up-sampling is the same
        if layout_name == "SPEC_FANCY" or cases[i][0] == "444":
            assert np.abs(outs[i].astype(np.int16) - pil).mean() < 1.0, (i, np.abs(outs[i].astype(np.int16) - pil).mean())
    assert np.array_equal(outs[-1], ref[-1])
    # the single-file entry point
    img = JPEGImage.parse(planar[0], ext=EXT_DRI | EXT_MULTISCAN, layout=layout)
    assert (img.width(), img.height()) == (640, 480) and np.array_equal(img.rgb(), outs[0])
    # a damaged chroma scan fails the frame, not its neighbours
    bad = bytearray(planar[1])
    last_sos = bytes(bad).rindex(b"\xff\xda")
    del bad[last_sos + 40:]
    outs2, st2 = decode_scans([planar[0], bytes(bad), planar[2]], ext=EXT_DRI, layout=layout)
    assert st2[0] == 0 and st2[2] == 0 and st2[1] != 0
    assert np.array_equal(outs2[0], outs[0]) and np.array_equal(outs2[2], outs[2])
