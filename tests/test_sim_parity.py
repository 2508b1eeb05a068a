"""CPU simulation of the kernels' algorithm (tests/sim/jpsim.cpp: same per-thread code, same planner,
barrier structure replayed serially) against the oracle.  This is what the not-gpu suite can say about
the parallel decode: coefficients bit-exact, samples within +-1, bytes_read identical."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_ffi as O
import sim_ffi as S
from conftest import fixture_bytes
from jpeg_rust_b200 import _ffi, synth

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "known_answers.json")))


def check(files, layout=1, ext=0, gts=None):
    rs, diag = S.decode_batch(files, layout=layout, ext=ext)
    for k, (f, r) in enumerate(zip(files, rs)):
        o = O.decode(f, layout=layout, ext=ext)
        assert o.status == 0 and r.status == 0, (k, r.status, o.msg)
        assert len(r.coefs) == len(o.coefs)
        for a, b in zip(r.coefs, o.coefs):
            assert np.array_equal(a, b), f"image {k}: coefficients differ"
        if gts:
            for a, b in zip(r.coefs, gts[k]):
                assert np.array_equal(a, b)
        if not (ext & 2):
            assert r.bytes_read == o.bytes_read
        d = np.abs(r.rgb.astype(int) - o.rgb.astype(int))
        assert d.max() <= 1 and d.mean() < 0.01
    return diag


@pytest.mark.parametrize("name,layout,ext", [("lena.jpeg", 0, 0), ("lena-bw.jpeg", 0, 0), ("huff_simple0.jpg", 0, 1),
                                             ("2x2-chroma.jpeg", 1, 0), ("lena.jpeg", 1, 0)])
def test_fixtures(name, layout, ext):
    data = fixture_bytes(name)
    check([data], layout, ext)
    (r,), _ = S.decode_batch([data], layout=layout, ext=ext)
    key = "SPEC" if layout else "REF"
    stream = b"".join(c.astype("<i2").tobytes() for c in r.coefs)
    assert hashlib.sha256(stream).hexdigest() == GOLD["oracle"]["fixtures"][name][key]["coef_sha256"]


@pytest.mark.parametrize("sub", ["420", "422", "444", "440", "gray"])
def test_shapes_mixed_batch(sub):
    files, gts = [], []
    for i, (w, h) in enumerate([(64, 64), (17, 9), (250, 131), (1, 1), (333, 200), (8, 8), (128, 16)]):
        f, g = synth.synth_jpeg(100 + i, w, h, sub, want_coefs=True)
        files.append(f)
        gts.append(g)
    check(files, gts=gts)


@pytest.mark.parametrize("sub,ri", [("444", 1), ("444", 7), ("420", 3), ("gray", 1), ("422", 16), ("444", 80)])
def test_restart_intervals(sub, ri):
    files, gts = [], []
    for i, (w, h) in enumerate([(320, 240), (250, 131), (64, 64)]):
        f, g = synth.synth_jpeg(200 + i, w, h, sub, restart_interval=ri, want_coefs=True)
        files.append(f)
        gts.append(g)
    check(files, ext=2, gts=gts)


def test_multi_sequence_image():
    """An image longer than one sequence (CTA of subsequences): chain verification across CTA boundaries."""
    f, g = synth.synth_jpeg(0, 1280, 720, "420", quality=92, want_coefs=True)
    diag = check([f], gts=[g])
    assert diag["sync_decodes"] > 128


@pytest.mark.parametrize("lookback", [0, 64, 512])
def test_short_lookback_is_repaired(lookback, monkeypatch):
    """With (almost) no look-back most cold starts are wrong: the verify/repair pass must mend every link."""
    monkeypatch.setenv("JPGPU_LOOKBACK_BITS", str(lookback))
    files = [synth.synth_jpeg(900 + i, 320, 240, s, want_coefs=True) for i, s in enumerate(["420", "444", "gray"])]
    diag = check([f for f, _ in files], gts=[g for _, g in files])
    if lookback == 0:
        assert diag["repairs"] > 0


def test_ref_layout_classes_equal_to_spec():
    check([synth.synth_jpeg(300, 320, 240, "422"), synth.synth_jpeg(301, 160, 120, "444"),
           synth.synth_jpeg(302, 200, 96, "gray")], layout=0)


REF_SHAPES = [("420", 64, 64), ("420", 48, 32), ("420", 752, 592), ("420", 250, 131), ("420", 512, 64), ("444", 61, 45),
              ("gray", 33, 17), ("422", 100, 40), ("440", 64, 48), ("420", 16, 16), ("420", 8, 8), ("444", 7, 5)]


@pytest.mark.parametrize("sub,w,h", REF_SHAPES)
def test_ref_layout_gather_path(sub, w, h):
    """REF placement where it differs from SPEC (4:2:0, ragged widths; decoder.rs:259-312, 347-379): the
    host-built placement map + gather arithmetic must reproduce the oracle's REF output, or raise the same panic."""
    f = synth.synth_jpeg(700 + w + h, w, h, sub)
    o = O.decode(f, layout=0)
    (r,), _ = S.decode_batch([f], layout=0)
    if o.status != 0:
        assert r.status == o.status, (r.status, o.msg)
        return
    assert r.status == 0, r.status
    for a, b in zip(r.coefs, o.coefs):
        assert np.array_equal(a, b)
    assert r.bytes_read == o.bytes_read
    d = np.abs(r.rgb.astype(int) - o.rgb.astype(int))
    assert d.max() <= 1 and d.mean() < 0.01


def test_ref_layout_random_shapes():
    """100 seeded shapes (1..150 pixels a side, every sampling mode) under the reference's own placement: about a third of
    them panic in the reference (decoder.rs:303-312 indexes past the image when the sub-sampled planes do not tile it);
    the planner must report exactly those, and reproduce the pixels of the others."""
    import random
    rng = random.Random(5)
    panics = 0
    for it in range(100):
        sub = rng.choice(["420", "420", "422", "440", "444", "gray"])
        w, h = rng.randint(1, 150), rng.randint(1, 150)
        f = synth.synth_jpeg(1000 + it, w, h, sub)
        o = O.decode(f, layout=0)
        (r,), _ = S.decode_batch([f], layout=0)
        if o.status != 0:
            panics += 1
            assert r.status == o.status, (sub, w, h, r.status, o.status, o.msg)
            continue
        assert r.status == 0, (sub, w, h, r.status)
        for a, b in zip(r.coefs, o.coefs):
            assert np.array_equal(a, b), (sub, w, h)
        assert r.bytes_read == o.bytes_read
        assert np.abs(r.rgb.astype(int) - o.rgb.astype(int)).max() <= 1, (sub, w, h)
    assert 10 < panics < 60


def test_spec_layout_random_shapes_restart_intervals_and_qualities():
    """120 seeded files: 1..200 pixels a side, every sampling mode, restart intervals of 1..16 MCUs (intervals of fewer
    than four bytes at quality 10 included), qualities 10..100, a fifth with optimised tables - decoded in batches of 40."""
    import random
    rng = random.Random(9)
    files = []
    for it in range(120):
        sub = rng.choice(["420", "420", "422", "440", "444", "gray"])
        w, h = rng.randint(1, 200), rng.randint(1, 200)
        files.append(synth.synth_jpeg(2000 + it, w, h, sub, quality=rng.choice([10, 50, 90, 100]),
                                      restart_interval=rng.choice([0, 0, 1, 2, 3, 7, 16]), optimize=rng.random() < 0.2))
    compared = 0
    for lo in range(0, len(files), 40):
        rs, _ = S.decode_batch(files[lo:lo + 40], layout=_ffi.LAYOUT_SPEC, ext=_ffi.EXT_DRI)
        for f, r in zip(files[lo:lo + 40], rs):
            o = O.decode(f, layout=_ffi.LAYOUT_SPEC, ext=_ffi.EXT_DRI)
            if o.status in (9, 10):     # one-bit codes of an optimised table: the reference cannot decode them
                continue
            assert r.status == o.status == 0, (r.status, o.status, o.msg)
            for a, b in zip(r.coefs, o.coefs):
                assert np.array_equal(a, b)
            assert r.bytes_read == o.bytes_read
            assert np.abs(r.rgb.astype(int) - o.rgb.astype(int)).max() <= 1
            compared += 1
    assert compared >= 100


def test_ref_layout_2x2_chroma_fixture():
    """configs[2]: 2x2-chroma.jpeg with the reference's own (buggy) 4:2:0 placement, 1763 of 1786 MCUs."""
    data = fixture_bytes("2x2-chroma.jpeg")
    check([data], layout=0)
    (r,), _ = S.decode_batch([data], layout=0)
    stream = b"".join(c.astype("<i2").tobytes() for c in r.coefs)
    assert hashlib.sha256(stream).hexdigest() == GOLD["oracle"]["fixtures"]["2x2-chroma.jpeg"]["REF"]["coef_sha256"]
    assert r.bytes_read == 144537


def test_spec_layout_generic_sampling_goes_through_the_gather_path():
    """Sampling the fused kernels have no variant for (gray declared as 2x2 in REF layout)."""
    f = bytearray(synth.synth_jpeg(800, 40, 24, "gray"))
    sof = bytes(f).index(b"\xff\xc0")
    f[sof + 11] = 0x22
    o = O.decode(bytes(f), layout=0)
    (r,), _ = S.decode_batch([bytes(f)], layout=0)
    if o.status in (9, 10):      # huffman.rs:156/162 lookup panics are reported as JPGPU_ERR_BAD_CODE
        assert r.status in (38, 37)
    else:
        assert r.status == o.status
    if o.status == 0:
        assert np.abs(r.rgb.astype(int) - o.rgb.astype(int)).max() <= 1


@pytest.mark.parametrize("q", [5, 100])
def test_quality_extremes(q):
    check([synth.synth_jpeg(400 + i, 160, 120, "420", quality=q) for i in range(2)])


def test_flat_images():
    flat = np.full((128, 128, 3), 128, np.uint8)
    check([synth.encode(flat, "420"), synth.encode(flat, "gray"), synth.encode(np.zeros((256, 256, 3), np.uint8), "444")])


def test_truncated_and_unsupported_inputs_get_a_status():
    good = synth.synth_jpeg(600, 64, 64, "420")
    rs, _ = S.decode_batch([good, good[:len(good) // 2], good])
    assert rs[0].status == 0 and rs[2].status == 0 and rs[1].status != 0
    assert np.array_equal(rs[0].rgb, rs[2].rgb)


def test_idct_factorisation_matches_the_direct_form():
    rng = np.random.default_rng(0)
    worst = 0.0
    for _ in range(100):
        blk = np.zeros(64, np.float32)
        k = rng.integers(1, 24)
        blk[rng.choice(64, k, replace=False)] = rng.integers(-400, 400, k)
        worst = max(worst, float(np.abs(O.idct_8x8(blk.reshape(8, 8)) - S.idct_8x8(blk.reshape(8, 8))).max()))
    assert worst < 2e-3


def test_corrupted_scans_never_derail_the_batch():
    """Random damage inside the entropy-coded data: every image gets a status, undamaged neighbours are unaffected,
    and where the oracle still decodes (damage that keeps the stream decodable) the coefficients agree."""
    rng = np.random.default_rng(7)
    good = [synth.synth_jpeg(950 + i, 96, 64, s, restart_interval=ri) for i, (s, ri) in enumerate([("420", 0), ("444", 3), ("gray", 0)])]
    files = []
    for k in range(24):
        f = bytearray(good[k % 3])
        sos = bytes(f).index(b"\xff\xda") + 14
        for _ in range(1 + k % 4):
            pos = int(rng.integers(sos, len(f) - 2))
            f[pos] ^= 1 << int(rng.integers(0, 8))
        files.append(bytes(f))
    files += good
    rs, _ = S.decode_batch(files, ext=2)
    for r, f in zip(rs[-3:], good):
        o = O.decode(f, layout=1, ext=2)
        assert r.status == 0 and np.abs(r.rgb.astype(int) - o.rgb.astype(int)).max() <= 1
    for r, f in zip(rs[:-3], files[:-3]):
        o = O.decode(f, layout=1, ext=2)
        if r.status == 0 and o.status == 0:
            assert all(np.array_equal(a, b) for a, b in zip(r.coefs, o.coefs))


def test_image_specific_huffman_tables_and_16bit_dqt():
    """Files the reference accepts but the Annex-K corpus does not cover: optimised (T.81 K.2) Huffman tables, a
    different set per image (many LUT sets in one batch), and 16-bit quantisation tables (mod.rs:245-256)."""
    files, gts = [], []
    for i, (sub, q, opt, wide) in enumerate([("420", 85, True, False), ("444", 30, True, False), ("gray", 95, True, False),
                                             ("422", 8, False, True), ("420", 3, True, True), ("440", 60, True, False)]):
        f, g = synth.synth_jpeg(1000 + i, 200 + 8 * i, 136, sub, quality=q, want_coefs=True, optimize=opt, dqt16=wide)
        files.append(f)
        gts.append(g)
    rs, _ = S.decode_batch(files)
    n_oracle = 0
    for f, g, r in zip(files, gts, rs):
        assert r.status == 0
        assert all(np.array_equal(a, b) for a, b in zip(r.coefs, g))
        o = O.decode(f, layout=1)
        if o.status == 0:               # within the reference's subset (no 1-bit code in any table)
            n_oracle += 1
            assert all(np.array_equal(a, b) for a, b in zip(r.coefs, o.coefs)) and r.bytes_read == o.bytes_read
            assert np.abs(r.rgb.astype(int) - o.rgb.astype(int)).max() <= 1
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
        (r,), _ = S.decode_batch([f])
        (rp,), _ = S.decode_batch([plain])
        assert r.status == 0 and rp.status == 0
        assert all(np.array_equal(a, b) for a, b in zip(r.coefs, g))
        assert np.array_equal(r.rgb, rp.rgb)


@pytest.mark.parametrize("parts", [2, 4, 8])
def test_write_pass_units(parts, monkeypatch):
    """The write pass cuts every subsequence into units at checkpoint boundaries and starts each unit from the state
    of its subsequence at A carried over the checkpoint records before it (decode_write_kernel): same coefficients
    for every cut — repaired links (short look-back), restart intervals on the general path and on the interval path."""
    monkeypatch.setenv("JPGPU_WRITE_PARTS", str(parts))
    monkeypatch.setenv("JPGPU_LOOKBACK_BITS", "128")
    files = [fixture_bytes("lena.jpeg"), synth.synth_jpeg(31, 640, 360, "420", quality=92),
             synth.synth_jpeg(32, 320, 240, "444", restart_interval=40), synth.synth_jpeg(33, 320, 240, "422", restart_interval=2)]
    rs, diag = S.decode_batch(files, layout=1, ext=2, sub_bits=8192)
    assert diag["repairs"] > 0
    for k, (f, r) in enumerate(zip(files, rs)):
        o = O.decode(f, layout=1, ext=2)
        assert o.status == 0 and r.status == 0, (k, r.status, o.msg)
        for a, b in zip(r.coefs, o.coefs):
            assert np.array_equal(a, b), f"image {k}: coefficients differ"
        assert np.abs(r.rgb.astype(int) - o.rgb.astype(int)).max() <= 1


@pytest.mark.parametrize("lookback", ["1024", None])
def test_interval_start_on_a_subsequence_boundary(lookback, monkeypatch):
    """Regression: in this image restart interval 678 starts exactly on a subsequence boundary and its predecessor's last
    symbol ends on the interval's last bit (no pad bits), so the predecessor stops without having crossed.  The
    subsequence starting there stands on the first bit of an interval — an absolute state — and must record it as such,
    or every DC predictor up to the next marker is off by the previous interval's sums (coefficients against the
    encoder's own; the oracle's O(N^4) IDCT would take minutes on 16 Mpixel)."""
    if lookback:
        monkeypatch.setenv("JPGPU_LOOKBACK_BITS", lookback)
    f, gt = synth.synth_jpeg(77, 4096, 4096, "420", restart_interval=64, want_coefs=True)
    rs, diag = S.decode_batch([f], layout=1, ext=2)
    assert rs[0].status == 0
    for a, g in zip(rs[0].coefs, gt):
        assert np.array_equal(a, g)


def test_randomised_configurations(monkeypatch):
    """A fixed-seed sweep over what the planner can vary (subsequence length, look-back, write-pass units, interval
    mode) and what the input can (sampling, size, quality, restart interval): coefficients against the encoder's own.
    (A 10 000-image run of the same generator is how the interval-start fix above was validated.)"""
    import random
    rng = random.Random(20261017)
    for it in range(60):
        sub = rng.choice(["420", "420", "422", "444", "440", "gray"])
        w = rng.choice([rng.randint(8, 300), rng.randint(300, 1100)])
        h = rng.choice([rng.randint(8, 300), rng.randint(300, 900)])
        ri = rng.choice([0, 0, 1, 2, 3, 5, 8, 16, 33, 64, 100, 256, 1000])
        q = rng.choice([30, 60, 85, 95])
        sb = rng.choice([0, 1024, 2048, 4096, 8192])
        monkeypatch.setenv("JPGPU_LOOKBACK_BITS", str(rng.choice([64, 256, 1024, 1024, 4096])))
        monkeypatch.setenv("JPGPU_WRITE_PARTS", str(rng.choice([1, 1, 2, 4])))
        if rng.random() < 0.3:
            monkeypatch.setenv("JPGPU_INTERVAL_MODE", str(rng.choice([0, 1])))
        else:
            monkeypatch.delenv("JPGPU_INTERVAL_MODE", raising=False)
        seed = rng.randint(0, 10 ** 6)
        f, gt = synth.synth_jpeg(seed, w, h, sub, quality=q, restart_interval=ri, want_coefs=True)
        rs, _ = S.decode_batch([f], layout=1, ext=2, sub_bits=sb)
        assert rs[0].status == 0, (it, seed, w, h, sub, ri, q, sb)
        for a, g in zip(rs[0].coefs, gt):
            assert np.array_equal(a, g), (it, seed, w, h, sub, ri, q, sb)


def test_position_saturates_on_a_flooded_scan():
    """an 8x8 image with 1-bit DC/EOB codes followed by > 8.39 MB of zero scan bytes takes the
    running coefficient position of the synchronisation pass past 2^31.  The position saturates (fold_advance), the
    declared block decodes, the flood behind it is ignored."""
    good = synth.synth_jpeg(601, 64, 64, "420")
    rs, _ = S.decode_batch([good, synth.crafted_flood_jpeg(8_500_000), good, synth.crafted_flood_jpeg(17_000_000)])
    assert [r.status for r in rs] == [0, 0, 0, 0]
    for k in (1, 3):
        assert rs[k].bytes_read == 1 and (rs[k].rgb == 128).all()
    assert np.array_equal(rs[0].rgb, rs[2].rgb)


def test_header_claim_is_tied_to_the_scan_bytes():
    """a small file whose SOF0 declares 30001 x 30001 must not make the planner size anything
    by its header (10 GB of placement map on the REF gather path): it cannot hold that many blocks and is turned away
    as truncated at plan time."""
    import time
    f = bytearray(synth.synth_jpeg(602, 64, 64, "420"))
    sof = bytes(f).index(b"\xff\xc0")
    f[sof + 5:sof + 9] = (30001).to_bytes(2, "big") + (30001).to_bytes(2, "big")
    t0 = time.time()
    rs, _ = S.decode_batch([bytes(f)], layout=0)
    assert rs[0].status == 37 and time.time() - t0 < 2.0     # JPGPU_ERR_TRUNCATED, without the 10 GB detour


def test_multi_symbol_sync_records_equal_the_single_symbol_ones():
    """The synchronisation pass through the multi-symbol tables (sync_multi_kernel / multi_symbol()) must write down
    exactly the states the symbol-by-symbol pass writes - every SubInfo and SegRec, not just the final pixels - because
    the chain verification and the repair walks (symbol by symbol) compare against them."""
    rng = np.random.default_rng(11)
    cases = []
    for name, layout, ext in (("lena.jpeg", 0, 0), ("lena-bw.jpeg", 0, 0), ("huff_simple0.jpg", 0, 1), ("2x2-chroma.jpeg", 1, 0)):
        cases.append(([fixture_bytes(name)], layout, ext))
    cases.append(([synth.synth_jpeg(9000 + i, 320 + 16 * i, 200 + 8 * i, s) for i, s in enumerate(["420", "444", "422", "gray", "440"])], 1, 0))
    cases.append(([synth.synth_jpeg(9010 + i, 256, 192, "420", restart_interval=ri) for i, ri in enumerate([1, 5, 64])], 1, 2))
    cases.append(([synth.synth_jpeg(9020, 384, 256, "420", optimize=True), synth.synth_jpeg(9021, 384, 256, "444", optimize=True, dqt16=True),
                   synth.synth_jpeg(9022, 640, 480, "420", quality=98, noise_sigma=25.0), synth.synth_jpeg(9023, 640, 480, "420", quality=20)], 1, 0))
    cases.append(([synth.encode(rng.integers(0, 256, (96, 128, 3), dtype=np.uint8), "444", quality=100)], 1, 0))
    cases.append(([synth.crafted_flood_jpeg(300_000)], 1, 0))
    try:
        for files, layout, ext in cases:
            for sub_bits in (0, 1024):
                S.set_sync_multi(True)
                a, da = S.decode_batch(files, layout=layout, ext=ext, sub_bits=sub_bits)
                S.set_sync_multi(False)
                b, db = S.decode_batch(files, layout=layout, ext=ext, sub_bits=sub_bits)
                assert da["sync_digest"] == db["sync_digest"] and da["repairs"] == db["repairs"]
                for x, y in zip(a, b):
                    assert x.status == y.status and x.bytes_read == y.bytes_read and np.array_equal(x.rgb, y.rgb)
    finally:
        S.set_sync_multi(True)


def test_multi_symbol_table_entries():
    """build_multi_lut against a brute-force decode of every window with the Annex-K luminance AC table."""
    import ctypes as C
    from jpeg_rust_b200 import parse_descriptor
    st, d, _ = parse_descriptor(synth.synth_jpeg(1, 16, 16, "gray"))
    assert st == 0
    bits, vals = list(d.ac_bits[0]), list(d.ac_vals[0])
    codes, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            codes[(length, code)] = vals[k]
            code += 1
            k += 1
        code <<= 1
    out = (C.c_uint32 * (1 << 16))()
    n = S.lib().jpsim_build_multi_lut((C.c_uint8 * 16)(*bits), (C.c_uint8 * 256)(*(vals + [0] * (256 - len(vals)))), 0, out, 1 << 16)
    K = n.bit_length() - 1          # kMultiBitsAc
    assert n == 1 << K and 9 <= K <= 14
    rng = np.random.default_rng(5)
    for w in [0, 1, n - 1, n - 2, 0x555 & (n - 1), 0xAAA & (n - 1)] + [int(x) for x in rng.integers(0, n, 300)]:
        pos, syms = 0, []
        while pos < K:
            hit = None
            for length in range(1, 17):
                if pos + length > K:
                    break
                c = (w >> (K - pos - length)) & ((1 << length) - 1)
                if (length, c) in codes:
                    hit = (length, codes[(length, c)])
                    break
            if hit is None:
                break
            length, sym = hit
            size = 0 if sym in (0x00, 0xF0) else sym & 15
            adv = 64 if sym == 0 else (16 if sym == 0xF0 else (sym >> 4) + 1)
            if syms and sum(a for _, a in syms) + (0 if sym == 0 else adv) > 63:
                break
            syms.append((length + size, adv))
            pos += length + size
            if sym == 0:
                break
        e = out[w]
        if not syms:
            assert e == 0
            continue
        assert e & 31 == sum(t for t, _ in syms)
        assert (e >> 5) & 127 == sum(a for _, a in syms)
        assert (e >> 12) & 63 == sum(a for _, a in syms[:-1])
        assert (e >> 18) & 31 == syms[0][0] and (e >> 23) & 127 == syms[0][1]


def test_compose_path_fancy_upsampling_and_non_interleaved_scans():
    """The compose path in the CPU simulation (sim_block_idct_all + sim_compose = block_idct_kernel + compose_colour_kernel):
    fancy chroma up-sampling against the numpy restatement on the oracle's planes, and files of non-interleaved scans
    (jpgpu_parse_scans) against the same coefficients coded as one interleaved scan - every scan's coefficients equal the
    encoder's, the pixels identical, a damaged scan fails its frame alone."""
    from fancy_ref import fancy_reference
    cases = [("420", 131, 77, 5), ("422", 200, 100, 0), ("440", 96, 80, 0), ("420", 320, 240, 0)]
    for sub, w, h, ri in cases:
        inter = synth.synth_jpeg(40, w, h, sub, restart_interval=ri)
        planar, gts = synth.synth_jpeg(40, w, h, sub, restart_interval=ri, planar_scans=True, want_coefs=True)
        want = fancy_reference(inter, 2)
        rs, _ = S.decode_batch([inter], layout=2, ext=2)
        assert rs[0].status == 0
        d = np.abs(rs[0].rgb.astype(int) - want.astype(int))
        assert d.max() <= 1 and d.mean() < 0.01, (sub, d.max(), d.mean())
        for layout in (1, 2):
            outs, st, parts = S.decode_scans([planar], layout=layout, ext=2)
            ref, _ = S.decode_batch([inter], layout=layout, ext=2)
            # (chroma scans are decoded as one-component images, level shift +128 included, and lose it again in the compose
            # step: the float round trip flips a truncation in a pixel or two)
            dd = np.abs(outs[0].astype(int) - ref[0].rgb.astype(int))
            assert st == [0] and dd.max() <= 1 and dd.mean() < 1e-3
            for part, g in zip(parts[0], gts):
                assert np.array_equal(part.coefs[0], g)
    good = synth.synth_jpeg(41, 96, 64, "420", planar_scans=True)
    bad = bytearray(synth.synth_jpeg(42, 96, 64, "420", planar_scans=True))
    del bad[bytes(bad).rindex(b"\xff\xda") + 30:]
    outs, st, _ = S.decode_scans([good, bytes(bad), good], layout=1)
    assert st[0] == 0 and st[2] == 0 and st[1] != 0 and np.array_equal(outs[0], outs[2])
