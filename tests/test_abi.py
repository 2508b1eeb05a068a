"""The C-ABI library loads on a box without a GPU, exports every symbol include/jpgpu.h declares, agrees with
the ctypes mirror on struct layout, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT, fixture_bytes
from jpeg_rust_b200 import JPEGImage, JpgpuError, _ffi

HEADER = os.path.join(ROOT, "include", "jpgpu.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jpgpu_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    L = _ffi.lib()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/jpgpu.h but not exported by libjpgpu.so"
    assert sorted(_ffi.EXPORTED_SYMBOLS) == names


def test_abi_version_and_status_strings():
    L = _ffi.lib()
    assert L.jpgpu_abi_version() == 3
    assert "restart interval" in _ffi.status_string(_ffi.PANIC_DRI)
    assert "no CPU fallback" in _ffi.status_string(_ffi.ERR_NO_DEVICE)


def test_struct_layout_matches_the_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "jpgpu.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(jpgpu_image_desc), offsetof(jpgpu_image_desc, qt), offsetof(jpgpu_image_desc, dc_bits),"
                   "offsetof(jpgpu_image_desc, ac_vals), offsetof(jpgpu_image_desc, restart_interval),"
                   "offsetof(jpgpu_image_desc, scan), offsetof(jpgpu_image_desc, frame_part), offsetof(jpgpu_image_desc, frame_hmax));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    D = _ffi.ImageDesc
    assert got == [C.sizeof(D), D.qt.offset, D.dc_bits.offset, D.ac_vals.offset, D.restart_interval.offset, D.scan.offset,
                   D.frame_part.offset, D.frame_hmax.offset]


def test_product_sources_never_reach_the_oracle_or_the_simulation():
    pkg = os.path.join(ROOT, "jpeg_rust_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_ffi" not in text and "liboracle" not in text and "jpsim" not in text, f
    mk = open(os.path.join(ROOT, "Makefile")).read()
    line = [l for l in mk.splitlines() if l.startswith("$(LIB)/libjpgpu.so:")][0]
    assert "oracle" not in line and "sim" not in line


def test_decode_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is for boxes without one")
    with pytest.raises(JpgpuError) as e:
        JPEGImage.parse(fixture_bytes("lena.jpeg"))
    assert e.value.status == _ffi.ERR_NO_DEVICE


def test_c_consumer_builds_and_fails_cleanly_without_a_gpu():
    """tests/c/abi_smoke.c: plain C against include/jpgpu.h + libjpgpu.so (what a cgo / JNI / Rust binding does).
    It must build with gcc alone; without a GPU it stops at jpgpu_create with JPGPU_ERR_NO_DEVICE - no CPU fallback."""
    import subprocess
    exe = os.path.join(ROOT, "tests", "c", "abi_smoke")
    subprocess.check_call(["make", "-s", "-C", ROOT, "tests/c/abi_smoke"])
    assert os.path.exists(exe)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "fixtures", "lena.jpeg")], capture_output=True, text=True)
        assert r.returncode == 33 and "no usable CUDA device" in r.stderr


def test_panic_messages_are_the_references_own():
    L = _ffi.lib()
    assert L.jpgpu_panic_message(_ffi.PANIC_DRI) == b"got to restart interval def"            # mod.rs:427
    assert L.jpgpu_panic_message(_ffi.PANIC_AC_LOOKUP) == b"ILLEGAL STATE!"                   # huffman.rs:162
    assert L.jpgpu_panic_message(_ffi.PANIC_READ_BITS_ASSERT) == b"Should not read more than 16 bits at a time!"
    assert L.jpgpu_panic_message(_ffi.PANIC_COMPONENT_COUNT) == b"asd"                        # decoder.rs:330
    assert L.jpgpu_panic_message(_ffi.PANIC_DQT_PRECISION) == b"Unknown precision of quantization table"
    assert L.jpgpu_panic_message(_ffi.OK) is None and L.jpgpu_panic_message(_ffi.ERR_TRUNCATED) is None
    for s in range(1, 16):
        if s != _ffi.NO_SCAN:
            assert L.jpgpu_panic_message(s), s
