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
