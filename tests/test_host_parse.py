"""jpgpu_parse (the host parser behind the C ABI) against the oracle's restatement of JPEGImage::parse:
same fields on good files, same panic on bad ones.  Runs without a GPU."""
import ctypes as C

import numpy as np
import pytest

import oracle_ffi as O
from conftest import fixture_bytes
from jpeg_rust_b200 import EXT_DRI, EXT_NONE, EXT_SKIP_APPN, LAYOUT_REF, LAYOUT_SPEC, _ffi, parse_descriptor, synth


@pytest.mark.parametrize("name,ext", [("lena.jpeg", 0), ("2x2-chroma.jpeg", 0), ("lena-bw.jpeg", 0), ("huff_simple0.jpg", 1)])
def test_fields_match_the_oracle(name, ext):
    data = fixture_bytes(name)
    st, d, _ = parse_descriptor(data, ext, LAYOUT_REF)
    o = O.decode(data, ext=ext)
    assert st == 0
    assert (d.width, d.height, d.ncomp) == (o.width, o.height, o.ncomp)
    assert [d.comp[i].h for i in range(d.ncomp)] == o.hs[:o.ncomp]
    assert [d.comp[i].v for i in range(d.ncomp)] == o.vs[:o.ncomp]
    mcus, bpm, nb = C.c_uint32(), C.c_uint32(), (C.c_uint32 * 4)()
    assert _ffi.lib().jpgpu_geometry(C.byref(d), C.byref(mcus), C.byref(bpm), nb) == 0
    assert mcus.value == o.mcus_read                                  # decoder.rs:191-192
    assert list(nb)[:o.ncomp] == [len(c) for c in o.coefs]
    # the raw scan unstuffs to the oracle's data vector length (mod.rs:371-385)
    raw = bytes((C.c_uint8 * d.scan_len).from_address(d.scan))
    out = C.create_string_buffer(len(raw))
    assert O.lib().oracle_unstuff(raw, len(raw), out) == o.scan_len


def test_spec_geometry_uses_true_mcu_count():
    st, d, _ = parse_descriptor(fixture_bytes("2x2-chroma.jpeg"), 0, LAYOUT_SPEC)
    mcus = C.c_uint32()
    assert _ffi.lib().jpgpu_geometry(C.byref(d), C.byref(mcus), None, None) == 0
    assert mcus.value == 1786


def _seg(marker, payload):
    n = len(payload) + 2
    return bytes([0xff, marker, n >> 8, n & 255]) + payload


def mutations():
    good = synth.synth_jpeg(2, 48, 32, "420")
    sos = good.index(b"\xff\xda")
    sof = good.index(b"\xff\xc0")
    yield "dri", good[:sos] + _seg(0xdd, b"\x00\x04") + good[sos:]
    yield "app12", good[:sos] + _seg(0xec, b"Ducky\x00") + good[sos:]
    yield "app14", good[:sos] + _seg(0xee, b"Adobe\x00") + good[sos:]
    yield "app1-unknown-marker", good[:sos] + _seg(0xe1, b"Exif\x00\x00") + good[sos:]
    yield "sof2-unknown-marker", good[:sof] + b"\xff\xc2" + good[sof + 2:]
    yield "garbage-byte", good[:sos] + b"\x12" + good[sos:]
    yield "no-sos", good[:sos]
    yield "sos-before-sof", good[:sof] + good[sos:]
    bad = bytearray(good); bad[sof + 11] = 0x31
    yield "sampling-factor-3", bytes(bad)
    dqt = good.index(b"\xff\xdb")
    bad = bytearray(good); bad[dqt + 4] = 0x20
    yield "dqt-precision-2", bytes(bad)
    yield "truncated-header", good[:sof + 6]
    yield "ends-in-ff", good[:-1]
    yield "zero-length-segment", good[:sos] + b"\xff\xfe\x00\x01" + good[sos:]
    yield "dqt-16bit", good[:sos] + _seg(0xdb, bytes([0x12]) + bytes(range(1, 129))) + good[sos:]
    yield "empty", b""
    yield "soi-only", b"\xff\xd8"


@pytest.mark.parametrize("name,data", list(mutations()), ids=[n for n, _ in mutations()])
def test_panic_parity_on_malformed_input(name, data):
    st, d, _ = parse_descriptor(data, EXT_NONE, LAYOUT_REF)
    o = O.decode(data)
    if o.status == 0:
        assert st == 0
    elif o.status in (8, 9, 10, 11, 12):      # decode-stage panics: parse must have succeeded
        assert st == 0
    else:
        assert st == o.status, (name, st, o.status, o.msg)


def test_extensions_are_opt_in():
    good = synth.synth_jpeg(2, 48, 32, "444", restart_interval=3)
    assert parse_descriptor(good, EXT_NONE)[0] == _ffi.PANIC_DRI
    st, d, _ = parse_descriptor(good, EXT_DRI)
    assert st == 0 and d.restart_interval == 3
    h = fixture_bytes("huff_simple0.jpg")
    assert parse_descriptor(h, EXT_NONE)[0] == _ffi.PANIC_APP12_14
    assert parse_descriptor(h, EXT_SKIP_APPN)[0] == 0


def test_builder_produces_the_same_descriptor_as_the_parser():
    """The JPEGDecoder builder mirror (decoder.rs:55-152) and jpgpu_parse agree field by field."""
    from jpeg_rust_b200 import (FrameComponentHeader, FrameHeader, HuffmanTable, JPEGDecoder, ScanComponentHeader,
                                ScanHeader)
    data = fixture_bytes("lena.jpeg")
    st, d, _ = parse_descriptor(data)
    raw = bytes((C.c_uint8 * d.scan_len).from_address(d.scan))
    dec = (JPEGDecoder.new(raw)
           .frame_header(FrameHeader(8, d.height, d.width, d.ncomp,
                                     [FrameComponentHeader(d.comp[i].id, d.comp[i].h, d.comp[i].v, d.comp[i].tq) for i in range(3)]))
           .scan_header(ScanHeader(3, [ScanComponentHeader(d.comp[i].id, d.comp[i].td, d.comp[i].ta) for i in range(3)]))
           .dimensions((d.width, d.height)))
    for t in range(4):
        if d.ac_present[t]:
            dec.huffman_ac_tables(t, HuffmanTable.from_size_data_tables(bytes(d.ac_bits[t]), bytes(d.ac_vals[t])[:d.ac_nvals[t]]))
        if d.dc_present[t]:
            dec.huffman_dc_tables(t, HuffmanTable.from_size_data_tables(bytes(d.dc_bits[t]), bytes(d.dc_vals[t])[:d.dc_nvals[t]]))
        if d.qt_present[t]:
            dec.quantization_table(t, list(d.qt[t]))
    d2 = dec.descriptor()
    for f in ("width", "height", "ncomp", "restart_interval", "scan_len"):
        assert getattr(d, f) == getattr(d2, f)
    assert bytes(d.qt) == bytes(d2.qt) and bytes(d.ac_vals) == bytes(d2.ac_vals) and bytes(d.dc_bits) == bytes(d2.dc_bits)
    assert bytes(d.comp) == bytes(d2.comp)


def test_parse_scans_of_non_interleaved_files():
    """jpgpu_parse_scans (host only): one descriptor per scan; a frame of non-interleaved scans comes as consecutive
    one-component descriptors of the component's own size, the ordinary file as the one descriptor jpgpu_parse gives."""
    from jpeg_rust_b200 import EXT_DRI, parse_scans, plan_info
    f = synth.synth_jpeg(5, 131, 77, "420", restart_interval=5, planar_scans=True)
    st, ds, _buf = parse_scans(f, EXT_DRI)
    assert st == 0 and [d.frame_part for d in ds] == [1, 2, 2]
    assert [(d.width, d.height) for d in ds] == [(131, 77), (66, 39), (66, 39)]
    assert all((d.frame_width, d.frame_height, d.frame_ncomp, d.ncomp, d.restart_interval) == (131, 77, 3, 1, 5) for d in ds)
    assert [(d.frame_comp, d.frame_h, d.frame_v, d.frame_hmax, d.frame_vmax) for d in ds] == [(0, 2, 2, 2, 2), (1, 1, 1, 2, 2), (2, 1, 1, 2, 2)]
    assert [d.comp[0].tq for d in ds] == [0, 1, 1] and [d.comp[0].td for d in ds] == [0, 1, 1]
    # the scans tile the file: each ends where the next SOS segment (or EOI) begins
    ends = [d.scan + d.scan_len for d in ds]
    assert ends[0] < ds[1].scan and ends[1] < ds[2].scan and ends[2] == _buf.ctypes.data + len(f) - 2
    plain = synth.synth_jpeg(5, 64, 64, "420")
    st, ds, _ = parse_scans(plain)
    assert st == 0 and len(ds) == 1 and ds[0].frame_part == 0 and ds[0].ncomp == 3 and ds[0].scan_len == len(plain) - 2 - (ds[0].scan - _.ctypes.data)
    # without DRI in the extension flags the restart-interval file is the reference's panic
    assert parse_scans(f)[0] == 2
    # the planner takes a frame's scans as images of one batch
    st, ds, _buf = parse_scans(f, EXT_DRI)
    info = plan_info(descs=ds)
    assert info["groups"] == 1 and info["warp_jobs"] >= 3


def test_seeded_header_mutations_keep_the_references_verdict():
    """800 seeded mutations of a small file's header region (byte changes, deletions, insertions, truncation): the host
    side must give the reference's verdict - at parse time the parser's panic (mod.rs), for what the reference only
    meets inside decode() the descriptor check every decode starts with (jpgpu_geometry: table selectors and missing
    tables in the order decoder.rs:197-198 / 222-224 meets them, component count).  Entropy-stage panics (9, 10, 12,
    reads past the data) belong to the GPU tests.  35 / 36 are the library's own refusals (sampling factors outside the
    decodable subset, BITS/HUFFVAL that are no prefix code), 16 is the oracle's (more than four scan components).
    One ordering the host cannot know: a component whose entropy data fails in the first MCU panics in the reference
    before a later component's selector / table / sampling panic; the mutations here leave the entropy data intact."""
    import random
    good = synth.synth_jpeg(2, 48, 32, "420")
    hdr_end = good.index(b"\xff\xda") + 14
    rng = random.Random(1)
    seen = set()
    for it in range(800):
        b = bytearray(good)
        for _ in range(rng.choice([1, 1, 2, 3])):
            op = rng.random()
            pos = rng.randrange(2, max(3, min(hdr_end, len(b))))
            if pos >= len(b):
                continue
            if op < 0.6:
                b[pos] = rng.choice([0, 1, 0xff, rng.randrange(256), b[pos] ^ (1 << rng.randrange(8))])
            elif op < 0.75:
                del b[pos:pos + rng.randrange(1, 8)]
            elif op < 0.9:
                b[pos:pos] = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 6)))
            else:
                b = b[:rng.randrange(2, len(b))]
        data = bytes(b)
        st, d, _ = parse_descriptor(data, EXT_NONE, LAYOUT_REF)
        o = O.decode(data)
        gs = _ffi.lib().jpgpu_geometry(C.byref(d), None, None, None) if st == 0 else -1
        seen.add((st, gs, o.status))
        if st in (35, 36) or gs == 35 or o.status == 16:
            continue
        if st != 0:
            assert st == o.status, (it, st, o.status, o.msg)
        elif o.status in (0, 9, 10, 12):
            assert gs == 0, (it, gs, o.status, o.msg)
        elif o.status in (8, 11):
            assert gs == o.status, (it, gs, o.status, o.msg)
        elif o.status == 6:    # "the len is 4": a table selector >= 4 (host); any other length: a read past the scan data (GPU)
            assert gs == (6 if "the len is 4" in o.msg else 0), (it, gs, o.msg)
        else:
            raise AssertionError((it, st, gs, o.status, o.msg))
    assert {s for s, _, _ in seen} >= {0, 1, 4, 5, 6, 15}, seen   # the campaign reaches the panics it is about
