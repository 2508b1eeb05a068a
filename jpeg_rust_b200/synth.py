"""Offline input generator: deterministic synthetic images encoded by the in-repo baseline
JPEG encoder (jpeg_rust_b200/csrc/jpgenc.cpp -> lib/libjpgenc.so).  Host-only; used by
tests and bench.py to build the corpus the north-star names (there is no network)."""
import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "lib", "libjpgenc.so")
SEED_BASE = 0x5EED0000
_lib = None

SUBSAMPLING = {"420": (0, 2, 2), "422": (0, 2, 1), "444": (0, 1, 1), "440": (0, 1, 2), "gray": (1, 1, 1)}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise RuntimeError(f"{_LIB} is missing: run `make`")
        L = C.CDLL(_LIB)
        L.jpgenc_max_size.restype = C.c_size_t
        L.jpgenc_max_size.argtypes = [C.c_int, C.c_int]
        L.jpgenc_synth_rgb.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.jpgenc_synth_rgb.restype = None
        L.jpgenc_encode.restype = C.c_size_t
        L.jpgenc_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        L.jpgenc_encode_ex.restype = C.c_size_t
        L.jpgenc_encode_ex.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        L.jpgenc_synth_encode_ex.restype = C.c_size_t
        L.jpgenc_synth_encode_ex.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                             C.c_void_p]
        L.jpgenc_synth_encode.restype = C.c_size_t
        L.jpgenc_synth_encode.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        _lib = L
    return _lib


def synth_rgb(index, width, height, noise_sigma=6.0):
    out = np.empty((height, width, 3), np.uint8)
    lib().jpgenc_synth_rgb(SEED_BASE + index, width, height, noise_sigma, out.ctypes.data)
    return out


def _coef_capacity(width, height):
    return ((width + 15) // 16) * ((height + 15) // 16) * 12 * 64


def _flags(optimize, dqt16, planar_scans=False):
    return (1 if optimize else 0) | (2 if dqt16 else 0) | (4 if planar_scans else 0)


def encode(rgb, subsampling="420", quality=85, restart_interval=0, want_coefs=False, optimize=False, dqt16=False,
           planar_scans=False):
    """Encode an HxWx3 uint8 array. Returns bytes, or (bytes, [per-component (nblocks,64) int16]) with want_coefs.
    optimize: image-specific Huffman tables (T.81 K.2; may hold 1-bit codes the reference cannot decode);
    dqt16: 16-bit quantisation tables with unclamped scaling."""
    gray, hy, vy = SUBSAMPLING[subsampling]
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w = rgb.shape[:2]
    cap = lib().jpgenc_max_size(w, h)
    out = np.empty(cap, np.uint8)
    nb = (C.c_size_t * 3)()
    coef = np.zeros(_coef_capacity(w, h), np.int16) if want_coefs else None
    n = lib().jpgenc_encode_ex(rgb.ctypes.data, w, h, gray, hy, vy, quality, restart_interval, _flags(optimize, dqt16, planar_scans),
                               out.ctypes.data, cap, coef.ctypes.data if want_coefs else None,
                               coef.size if want_coefs else 0, nb)
    if n == 0:
        raise RuntimeError("jpgenc_encode failed")
    data = out[:n].tobytes()
    if not want_coefs:
        return data
    comps, off = [], 0
    for c in range(1 if gray else 3):
        comps.append(coef[off:off + nb[c] * 64].reshape(-1, 64).copy())
        off += nb[c] * 64
    return data, comps


def synth_jpeg(index, width, height, subsampling="420", quality=85, restart_interval=0, noise_sigma=6.0,
               want_coefs=False, optimize=False, dqt16=False, planar_scans=False):
    """Synthetic image `index` (seed 0x5EED0000 + index) as a baseline JPEG.  planar_scans: one non-interleaved scan per
    component (a file the reference stops reading after its first scan); the coefficients then come in raster order."""
    gray, hy, vy = SUBSAMPLING[subsampling]
    cap = lib().jpgenc_max_size(width, height)
    out = np.empty(cap, np.uint8)
    nb = (C.c_size_t * 3)()
    coef = np.zeros(_coef_capacity(width, height), np.int16) if want_coefs else None
    n = lib().jpgenc_synth_encode_ex(SEED_BASE + index, width, height, noise_sigma, gray, hy, vy, quality,
                                     restart_interval, _flags(optimize, dqt16, planar_scans), out.ctypes.data, cap,
                                     coef.ctypes.data if want_coefs else None, coef.size if want_coefs else 0, nb)
    if n == 0:
        raise RuntimeError("jpgenc_synth_encode failed")
    data = out[:n].tobytes()
    if not want_coefs:
        return data
    comps, off = [], 0
    for c in range(1 if gray else 3):
        comps.append(coef[off:off + nb[c] * 64].reshape(-1, 64).copy())
        off += nb[c] * 64
    return data, comps


def synth_corpus(count, width, height, subsampling="420", quality=85, restart_interval=0, noise_sigma=6.0,
                 threads=None, first_index=0):
    """`count` distinct synthetic JPEGs, generated on all host cores (ctypes releases the GIL)."""
    threads = threads or min(32, os.cpu_count() or 1)
    with ThreadPoolExecutor(threads) as ex:
        return list(ex.map(lambda i: synth_jpeg(first_index + i, width, height, subsampling, quality,
                                                restart_interval, noise_sigma), range(count)))


def crafted_flood_jpeg(scan_bytes, width=8, height=8):
    """A hand-made gray baseline JPEG whose header declares `width` x `height` but whose scan goes on for `scan_bytes`
    zero bytes: with a 1-bit DC code (size 0) and a 1-bit EOB every two bits are one more "block", so the running
    coefficient position of anything that decodes the whole stream exceeds 2^31 beyond 8.39 MB (the
    position must saturate).  The declared blocks decode to 128-gray; the rest is trailing garbage a decoder ignores."""
    def seg(marker, payload):
        return bytes([0xFF, marker]) + (len(payload) + 2).to_bytes(2, "big") + payload
    dqt = seg(0xDB, bytes([0]) + bytes([1] * 64))
    sof = seg(0xC0, bytes([8]) + height.to_bytes(2, "big") + width.to_bytes(2, "big") + bytes([1, 1, 0x11, 0]))
    dht_dc = seg(0xC4, bytes([0x00]) + bytes([1] + [0] * 15) + bytes([0]))        # one code "0": size category 0
    dht_ac = seg(0xC4, bytes([0x10]) + bytes([1] + [0] * 15) + bytes([0x00]))     # one code "0": EOB
    sos = seg(0xDA, bytes([1, 1, 0x00, 0, 63, 0]))
    return b"\xff\xd8" + dqt + sof + dht_dc + dht_ac + sos + bytes(scan_bytes) + b"\xff\xd9"
