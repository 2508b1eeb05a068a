"""The in-repo encoder: ground-truth coefficients, standard tables, libjpeg readability."""
import io

import numpy as np
import pytest

import oracle_ffi as O
from jpeg_rust_b200 import synth


@pytest.mark.parametrize("sub", ["420", "422", "444", "440", "gray"])
def test_ground_truth_coefficients_equal_the_oracle(sub):
    data, gt = synth.synth_jpeg(11, 200, 120, sub, want_coefs=True)
    r = O.decode(data, layout=O.LAYOUT_SPEC)
    assert r.status == 0, r.msg
    assert len(gt) == len(r.coefs)
    for a, b in zip(gt, r.coefs):
        assert np.array_equal(a, b)
    assert r.scan_len - r.bytes_read == 2


def test_libjpeg_reads_the_files_and_agrees_roughly():
    from PIL import Image
    rgb = synth.synth_rgb(3, 160, 96)
    for sub in ("444", "420", "gray"):
        data = synth.encode(rgb, sub, quality=90)
        im = np.asarray(Image.open(io.BytesIO(data)).convert("RGB")).astype(int)
        ref = rgb.astype(int) if sub != "gray" else None
        if ref is not None:
            assert np.abs(im - ref).mean() < 6
        o = O.decode(data, layout=O.LAYOUT_SPEC).rgb.astype(int)
        assert np.abs(im - o).mean() < (4.0 if sub == "420" else 1.5)   # libjpeg's fancy upsampling on noisy chroma


def test_only_reference_accepted_markers_unless_dri():
    data = synth.synth_jpeg(1, 64, 48, "420")
    i, seen = 2, []
    while i < len(data):
        assert data[i] == 0xff
        m = data[i + 1]
        seen.append(m)
        if m == 0xda:
            break
        i += 2 + ((data[i + 2] << 8) | data[i + 3])
    assert set(seen) <= {0xe0, 0xdb, 0xc0, 0xc4, 0xda}       # mod.rs:157-181 minus the panicking ones
    assert data[-2:] == b"\xff\xd9" and O.decode(data).status == 0
    assert O.decode(synth.synth_jpeg(1, 64, 48, "420", restart_interval=2)).status == 2     # mod.rs:427


def test_dri_invariance_in_the_oracle():
    a = O.decode(synth.synth_jpeg(9, 200, 120, "420"), layout=O.LAYOUT_SPEC)
    b = O.decode(synth.synth_jpeg(9, 200, 120, "420", restart_interval=5), layout=O.LAYOUT_SPEC, ext=O.EXT_DRI)
    assert b.status == 0
    assert all(np.array_equal(x, y) for x, y in zip(a.coefs, b.coefs)) and np.array_equal(a.rgb, b.rgb)
