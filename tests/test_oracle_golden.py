"""The oracle against every pin available for this path: the survey's known-answer
table, its own recorded outputs (tests/golden/known_answers.json, made by tests/golden/make_golden.py),
exact scan consumption, and libjpeg as a loose sanity bound."""
import hashlib
import io
import json
import os

import numpy as np
import pytest

import oracle_ffi as O
from conftest import fixture_bytes

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "known_answers.json")))


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.mark.parametrize("key", sorted(GOLD["survey"]))
def test_survey_known_answers(key):
    e = GOLD["survey"][key]
    name = key.split("#")[0]
    layout = O.LAYOUT_REF if e["layout"] == "REF" else O.LAYOUT_SPEC
    r = O.decode(fixture_bytes(name), layout=layout, ext=e["ext"])
    assert r.status == 0, r.msg
    assert r.mcus_read == e["mcus"]
    assert [len(c) for c in r.coefs] == e["nblocks"]
    assert r.bytes_read == e["bytes_read"]
    if "scan_len" in e:
        assert r.scan_len == e["scan_len"]
    assert sha(r.coefficient_stream()) == e["coef_sha256"]
    if "first_block" in e:
        assert list(r.coefs[0][0][:len(e["first_block"])]) == e["first_block"]
    for pos, want in e.get("pixels", {}).items():
        y, x = map(int, pos.split(","))
        assert np.abs(r.rgb[y, x].astype(int) - np.array(want)).max() <= 1   # survey values are +-1 (f64 model)
    if "means" in e:
        assert np.allclose(r.rgb.reshape(-1, 3).mean(0), e["means"], atol=0.02)


def test_exact_scan_consumption():
    """A correct entropy decode ends exactly 2 bytes (EOI) before the end of the unstuffed data."""
    for name in ("lena.jpeg", "lena-bw.jpeg"):
        r = O.decode(fixture_bytes(name))
        assert r.scan_len - r.bytes_read == 2
    r = O.decode(fixture_bytes("2x2-chroma.jpeg"), layout=O.LAYOUT_SPEC)
    assert r.scan_len - r.bytes_read == 2
    r = O.decode(fixture_bytes("huff_simple0.jpg"), ext=O.EXT_SKIP_APPN)
    assert r.scan_len - r.bytes_read == 2


def test_reference_panics_are_reproduced():
    r = O.decode(fixture_bytes("huff_simple0.jpg"))
    assert r.status == 3 and "ApplicationSegment12" in r.msg        # mod.rs:446


@pytest.mark.parametrize("name", sorted(GOLD["oracle"]["fixtures"]))
def test_recorded_oracle_outputs(name):
    ext = O.EXT_SKIP_APPN if name == "huff_simple0.jpg" else O.EXT_NONE
    for lname, layout in (("REF", O.LAYOUT_REF), ("SPEC", O.LAYOUT_SPEC)):
        e = GOLD["oracle"]["fixtures"][name][lname]
        r = O.decode(fixture_bytes(name), layout=layout, ext=ext)
        assert r.status == e["status"]
        assert sha(r.rgb.tobytes()) == e["rgb_sha256"]
        assert sha(r.coefficient_stream()) == e["coef_sha256"]
        assert r.bytes_read == e["bytes_read"] and r.mcus_read == e["mcus"]


def test_ref_equals_spec_where_the_survey_says_so():
    for name, ext in (("lena.jpeg", 0), ("lena-bw.jpeg", 0), ("huff_simple0.jpg", 1)):
        e = GOLD["oracle"]["fixtures"][name]
        assert e["REF"]["rgb_sha256"] == e["SPEC"]["rgb_sha256"]
    e = GOLD["oracle"]["fixtures"]["2x2-chroma.jpeg"]
    assert e["REF"]["rgb_sha256"] != e["SPEC"]["rgb_sha256"]    # the reference's H2V2 placement bugs


def test_cos_table_mode_is_bit_identical_to_cosf_per_term():
    data = fixture_bytes("lena-bw.jpeg")
    a = O.decode(data, cos_mode=O.COS_TABLE)
    b = O.decode(data, cos_mode=O.COS_CALL)
    assert np.array_equal(a.rgb, b.rgb)
    assert all(np.array_equal(x, y) for x, y in zip(a.planes, b.planes))


def test_libjpeg_sanity_bounds():
    """Not the parity gate: expected gaps reference<->libjpeg from SURVEY.md §4."""
    from PIL import Image
    g = np.asarray(Image.open(io.BytesIO(fixture_bytes("lena-bw.jpeg"))).convert("RGB")).astype(int)
    r = O.decode(fixture_bytes("lena-bw.jpeg")).rgb.astype(int)
    d = np.abs(g - r)
    assert d.max() <= 1 and d.mean() < 0.6          # reference truncates, libjpeg rounds
    c = np.asarray(Image.open(io.BytesIO(fixture_bytes("lena.jpeg"))).convert("RGB")).astype(int)
    r = O.decode(fixture_bytes("lena.jpeg")).rgb.astype(int)
    assert np.abs(c - r).mean() < 1.5               # libjpeg uses fancy upsampling


@pytest.mark.parametrize("e", GOLD["oracle"]["synthetic"], ids=lambda e: f"{e['index']}-{e['subsampling']}")
def test_recorded_synthetic_corpus(e):
    from jpeg_rust_b200 import synth
    data, gt = synth.synth_jpeg(e["index"], e["width"], e["height"], e["subsampling"], e["quality"],
                                e["restart_interval"], want_coefs=True)
    assert len(data) == e["file_len"] and sha(data) == e["file_sha256"]       # generator + encoder are deterministic
    assert sha(b"".join(c.astype("<i2").tobytes() for c in gt)) == e["coef_sha256"]
    if e["width"] * e["height"] > 700000:
        return                                                                # keep the CPU suite short
    r = O.decode(data, layout=O.LAYOUT_SPEC, ext=O.EXT_DRI if e["restart_interval"] else 0)
    assert sha(r.rgb.tobytes()) == e["rgb_sha256_spec"]
    assert sha(r.coefficient_stream()) == e["coef_sha256"]


def _libjpeg_planes(data):
    """libjpeg (via PIL) without its colour conversion: the decoder's own Y / Cb / Cr planes (draft mode YCbCr)."""
    from PIL import Image
    im = Image.open(io.BytesIO(data))
    if im.mode != "L":
        im.draft("YCbCr", im.size)
    g = np.asarray(im).astype(int)
    return g[:, :, None] if g.ndim == 2 else g


CASES = [("gray", 512, 512, 85, 0, False), ("gray", 1917, 1075, 85, 0, False), ("gray", 640, 480, 50, 0, False),
         ("gray", 640, 480, 98, 0, True), ("444", 512, 512, 85, 0, False), ("444", 333, 217, 85, 0, False),
         ("444", 640, 480, 98, 0, False), ("444", 640, 480, 70, 0, True), ("444", 640, 480, 30, 0, False), ("gray", 640, 480, 85, 7, False),
         ("444", 640, 480, 85, 16, False), ("420", 640, 480, 85, 0, False), ("422", 640, 480, 85, 0, False),
         ("420", 333, 217, 85, 5, False)]


@pytest.mark.parametrize("sub,w,h,q,ri,opt", CASES, ids=lambda v: str(v))
def test_libjpeg_second_opinion_per_component(sub, w, h, q, ri, opt):
    """VERDICT round 1 (weak #2): the one independent decoder in the image as a second opinion on the oracle's whole
    Huffman -> dequantise -> de-zigzag -> IDCT -> placement path.  Compared BEFORE colour conversion and truncation,
    plane by plane: the oracle's float samples, rounded, against libjpeg's integer-IDCT samples.  They agree within
    libjpeg's own rounding (max 1, mean |delta| ~ 0.01) on gray and 4:4:4 - every component - with and without restart
    intervals and optimised tables; on sub-sampled files libjpeg interpolates chroma, so only the luma plane is compared."""
    from jpeg_rust_b200 import synth
    data = synth.synth_jpeg(30 + w % 97, w, h, sub, q, ri, optimize=opt)
    g = _libjpeg_planes(data)
    o = O.decode(data, layout=O.LAYOUT_SPEC, ext=O.EXT_DRI if ri else O.EXT_NONE)
    if opt and o.status in (9, 10):
        pytest.skip("image-specific tables with a 1-bit code: the reference (and so the oracle) cannot decode them, huffman.rs:61")
    assert o.status == 0
    ncmp = len(o.planes) if sub in ("gray", "444") else 1
    for c in range(ncmp):
        want = np.clip(np.floor(o.planes[c].reshape(h, w) + 128.5), 0, 255).astype(int)
        d = np.abs(want - g[:, :, c])
        assert d.max() <= 1 and d.mean() < 0.03, (c, d.max(), d.mean())


def test_libjpeg_second_opinion_on_fixtures():
    for name, ext, ncmp in (("lena-bw.jpeg", 0, 1), ("huff_simple0.jpg", 1, 3), ("lena.jpeg", 0, 1), ("2x2-chroma.jpeg", 0, 1)):
        data = fixture_bytes(name)
        g = _libjpeg_planes(data)
        o = O.decode(data, layout=O.LAYOUT_SPEC, ext=ext)
        for c in range(ncmp):
            want = np.clip(np.floor(o.planes[c].reshape(o.height, o.width) + 128.5), 0, 255).astype(int)
            d = np.abs(want - g[:, :, c])
            assert d.max() <= 1 and d.mean() < 0.03, (name, c, d.max(), d.mean())
