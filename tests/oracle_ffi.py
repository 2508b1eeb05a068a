"""ctypes binding of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs; never by the product package."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")

LAYOUT_REF, LAYOUT_SPEC = 0, 1
EXT_NONE, EXT_SKIP_APPN, EXT_DRI = 0, 1, 2
COS_CALL, COS_TABLE = 0, 1


class _Result(C.Structure):
    _fields_ = [
        ("status", C.c_int), ("msg", C.c_char * 160),
        ("width", C.c_int), ("height", C.c_int), ("ncomp", C.c_int),
        ("hs", C.c_int * 4), ("vs", C.c_int * 4),
        ("mcus_read", C.c_int),
        ("bytes_read", C.c_size_t), ("scan_len", C.c_size_t),
        ("rgb", C.POINTER(C.c_uint8)), ("rgb_len", C.c_size_t),
        ("coefs", C.POINTER(C.c_int16) * 4), ("nblocks", C.c_size_t * 4),
        ("planes", C.POINTER(C.c_float) * 4),
    ]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.oracle_decode_file.restype = C.POINTER(_Result)
        L.oracle_decode_file.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int]
        L.oracle_free.argtypes = [C.POINTER(_Result)]
        L.oracle_idct_8x8.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int]
        L.oracle_build_codes.restype = C.c_int
        L.oracle_build_codes.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_uint8),
                                         C.POINTER(C.c_uint16), C.POINTER(C.c_uint8)]
        L.oracle_ycbcr_to_rgb.argtypes = [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_uint8)]
        L.oracle_f32_to_u8.restype = C.c_uint8
        L.oracle_f32_to_u8.argtypes = [C.c_float]
        L.oracle_value_correction.restype = C.c_int16
        L.oracle_value_correction.argtypes = [C.c_uint16, C.c_int]
        L.oracle_zigzag_indices.restype = C.POINTER(C.c_int)
        L.oracle_unstuff.restype = C.c_size_t
        L.oracle_unstuff.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.oracle_time_decode.restype = C.c_double
        L.oracle_time_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib = L
    return _lib


class OracleImage:
    """Result of JPEGImage::parse + JPEGDecoder::decode as restated by the oracle."""

    def __init__(self, r):
        self.status = r.status
        self.msg = r.msg.decode("utf-8", "replace")
        self.width, self.height, self.ncomp = r.width, r.height, r.ncomp
        self.hs, self.vs = list(r.hs), list(r.vs)
        self.mcus_read = r.mcus_read
        self.bytes_read, self.scan_len = r.bytes_read, r.scan_len
        self.rgb = None
        if r.rgb and r.rgb_len:
            self.rgb = np.ctypeslib.as_array(r.rgb, shape=(r.rgb_len,)).copy().reshape(self.height, self.width, 3)
        self.coefs = []
        self.planes = []
        for k in range(4):
            n = r.nblocks[k]
            if n and r.coefs[k]:
                self.coefs.append(np.ctypeslib.as_array(r.coefs[k], shape=(n * 64,)).copy().reshape(n, 64))
            elif k < max(self.ncomp, 0):
                self.coefs.append(np.zeros((0, 64), np.int16))
            if r.planes[k]:
                self.planes.append(np.ctypeslib.as_array(r.planes[k], shape=(self.height * self.width,)).copy()
                                   .reshape(self.height, self.width))

    def coefficient_stream(self):
        """components in scan order, blocks in decode order, 64 LE i16 zigzag, absolute DC."""
        return b"".join(c.astype("<i2").tobytes() for c in self.coefs)


def decode(data, layout=LAYOUT_REF, ext=EXT_NONE, cos_mode=COS_TABLE):
    L = lib()
    p = L.oracle_decode_file(bytes(data), len(data), layout, ext, cos_mode)
    try:
        return OracleImage(p.contents)
    finally:
        L.oracle_free(p)


def idct_8x8(block, cos_mode=COS_TABLE):
    a = np.ascontiguousarray(block, np.float32).reshape(64)
    out = np.empty(64, np.float32)
    lib().oracle_idct_8x8(a.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float)), cos_mode)
    return out.reshape(8, 8)


def time_decode(data, layout=LAYOUT_REF, ext=EXT_NONE, cos_mode=COS_CALL, reps=1):
    return lib().oracle_time_decode(bytes(data), len(data), layout, ext, cos_mode, reps)
