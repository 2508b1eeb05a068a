"""ctypes binding of the C ABI declared in include/jpgpu.h (libjpgpu.so, built in-tree).

This is the only way the package reaches the decode path.  There is no CPU fallback:
if the library is missing or no sm_100 device is usable, calls raise JpgpuError.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libjpgpu.so")

LAYOUT_REF, LAYOUT_SPEC, LAYOUT_SPEC_FANCY = 0, 1, 2
EXT_NONE, EXT_SKIP_APPN, EXT_DRI, EXT_MULTISCAN = 0, 1, 2, 4
OUT_RGB_INTERLEAVED, OUT_RGB_PLANAR, OUT_F32_PLANAR = 0, 1, 2
MEMORY_HOST, MEMORY_DEVICE = 0, 1

OK = 0
PANIC_UNHANDLED_MARKER, PANIC_DRI, PANIC_APP12_14, PANIC_DQT_PRECISION = 1, 2, 3, 4
PANIC_SAMPLING_ASSERT, PANIC_INDEX_OOB, PANIC_NO_FRAME_HEADER, PANIC_MISSING_TABLE = 5, 6, 7, 8
PANIC_DC_LOOKUP, PANIC_AC_LOOKUP, PANIC_COMPONENT_COUNT, PANIC_READ_BITS_ASSERT = 9, 10, 11, 12
PANIC_SCAN_COMPONENT, NO_SCAN, PANIC_ARITH = 13, 14, 15
ERR_INVALID_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED = 32, 33, 34, 35
ERR_BAD_HUFFMAN_TABLE, ERR_TRUNCATED, ERR_BAD_CODE, ERR_RESTART, ERR_OOM = 36, 37, 38, 39, 40


class Component(C.Structure):
    _fields_ = [("id", C.c_uint8), ("h", C.c_uint8), ("v", C.c_uint8),
                ("tq", C.c_uint8), ("td", C.c_uint8), ("ta", C.c_uint8)]


class ImageDesc(C.Structure):
    """jpgpu_image_desc: what mod.rs:388-413 hands to the JPEGDecoder builder."""
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("ncomp", C.c_uint32),
        ("comp", Component * 4),
        ("qt", (C.c_uint16 * 64) * 4), ("qt_present", C.c_uint8 * 4),
        ("dc_bits", (C.c_uint8 * 16) * 4), ("dc_vals", (C.c_uint8 * 256) * 4),
        ("dc_nvals", C.c_uint16 * 4), ("dc_present", C.c_uint8 * 4),
        ("ac_bits", (C.c_uint8 * 16) * 4), ("ac_vals", (C.c_uint8 * 256) * 4),
        ("ac_nvals", C.c_uint16 * 4), ("ac_present", C.c_uint8 * 4),
        ("restart_interval", C.c_uint32), ("layout", C.c_uint32),
        ("scan", C.c_void_p), ("scan_len", C.c_size_t),
        ("frame_part", C.c_uint32), ("frame_width", C.c_uint32), ("frame_height", C.c_uint32),
        ("frame_ncomp", C.c_uint8), ("frame_comp", C.c_uint8), ("frame_h", C.c_uint8), ("frame_v", C.c_uint8),
        ("frame_hmax", C.c_uint8), ("frame_vmax", C.c_uint8), ("frame_pad", C.c_uint8 * 2),
    ]


class JpgpuError(RuntimeError):
    def __init__(self, status, what=""):
        self.status = status
        msg = status_string(status) if _lib is not None else str(status)
        super().__init__(f"jpgpu status {status}: {msg}" + (f" ({what})" if what else ""))


_lib = None


def lib():
    """Loads libjpgpu.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()); "
                               "there is no CPU fallback for the decode path")
        L = C.CDLL(LIB_PATH)
        vp, u8p, sz = C.c_void_p, C.POINTER(C.c_uint8), C.c_size_t
        L.jpgpu_abi_version.restype = C.c_int
        L.jpgpu_status_string.restype = C.c_char_p
        L.jpgpu_status_string.argtypes = [C.c_int]
        L.jpgpu_panic_message.restype = C.c_char_p
        L.jpgpu_panic_message.argtypes = [C.c_int]
        L.jpgpu_parse.argtypes = [vp, sz, C.c_uint32, C.c_uint32, C.POINTER(ImageDesc)]
        L.jpgpu_parse_scans.argtypes = [vp, sz, C.c_uint32, C.c_uint32, C.POINTER(ImageDesc), sz, C.POINTER(sz)]
        L.jpgpu_geometry.argtypes = [C.POINTER(ImageDesc), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                     C.POINTER(C.c_uint32)]
        L.jpgpu_plan_info.argtypes = [C.POINTER(ImageDesc), sz, C.POINTER(C.c_uint64)]
        L.jpgpu_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.jpgpu_destroy.argtypes = [vp]
        L.jpgpu_destroy.restype = None
        L.jpgpu_last_error.restype = C.c_char_p
        L.jpgpu_last_error.argtypes = [vp]
        L.jpgpu_set_stream.argtypes = [vp, vp]
        L.jpgpu_sync.argtypes = [vp]
        L.jpgpu_decode.argtypes = [vp, C.POINTER(ImageDesc), vp, C.POINTER(sz)]
        L.jpgpu_decode_file.argtypes = [vp, vp, sz, C.c_uint32, C.c_uint32, vp, sz, C.POINTER(C.c_uint32),
                                        C.POINTER(C.c_uint32), C.POINTER(sz)]
        L.jpgpu_batch_create.argtypes = [vp, C.POINTER(ImageDesc), sz, C.POINTER(vp)]
        L.jpgpu_batch_replan.argtypes = [vp, C.POINTER(ImageDesc), sz]
        L.jpgpu_batch_destroy.argtypes = [vp]
        L.jpgpu_batch_destroy.restype = None
        for name in ("upload", "entropy", "idct", "decode"):
            getattr(L, "jpgpu_batch_" + name).argtypes = [vp]
        L.jpgpu_batch_set_device_scans.argtypes = [vp, vp, C.POINTER(C.c_uint64)]
        L.jpgpu_batch_set_device_output.argtypes = [vp, vp, sz]
        L.jpgpu_batch_set_output_format.argtypes = [vp, C.c_uint32]
        L.jpgpu_batch_set_normalisation.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.jpgpu_batch_download.argtypes = [vp, C.POINTER(vp)]
        L.jpgpu_batch_device_rgb.restype = vp
        L.jpgpu_batch_device_rgb.argtypes = [vp, sz, C.POINTER(sz)]
        L.jpgpu_batch_output_bytes.restype = sz
        L.jpgpu_batch_output_bytes.argtypes = [vp]
        L.jpgpu_batch_rgb_offset.argtypes = [vp, sz, C.POINTER(sz), C.POINTER(sz)]
        L.jpgpu_batch_results.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
        L.jpgpu_batch_coefficients.argtypes = [vp, sz, vp, sz, C.POINTER(C.c_uint32)]
        L.jpgpu_batch_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.jpgpu_batch_profile.argtypes = [vp, C.POINTER(C.c_float)]
        L.jpgpu_batch_launch_count.restype = C.c_uint64
        L.jpgpu_batch_launch_count.argtypes = [vp]
        L.jpgpu_stream.restype = vp
        L.jpgpu_stream.argtypes = [vp]
        L.jpgpu_batch_upload_from.argtypes = [vp, vp, sz]
        L.jpgpu_batch_download_contiguous.argtypes = [vp, vp, sz]
        L.jpgpu_pipeline_create.argtypes = [C.c_int, C.POINTER(ImageDesc), sz, sz, C.POINTER(vp)]
        L.jpgpu_pipeline_destroy.argtypes = [vp]
        L.jpgpu_pipeline_destroy.restype = None
        L.jpgpu_pipeline_output_bytes.restype = sz
        L.jpgpu_pipeline_output_bytes.argtypes = [vp]
        L.jpgpu_pipeline_image_offset.argtypes = [vp, sz, C.POINTER(sz), C.POINTER(sz)]
        L.jpgpu_pipeline_run.argtypes = [vp, vp, sz, vp, sz]
        L.jpgpu_pipeline_sync.argtypes = [vp]
        L.jpgpu_pipeline_elapsed_ms.argtypes = [vp, C.POINTER(C.c_float)]
        L.jpgpu_pipeline_results.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
        L.jpgpu_pipeline_launch_count.restype = C.c_uint64
        L.jpgpu_pipeline_launch_count.argtypes = [vp]
        L.jpgpu_pipeline_last_error.restype = C.c_char_p
        L.jpgpu_pipeline_last_error.argtypes = [vp]
        L.jpgpu_decode_batch_host.argtypes = [C.c_int, C.POINTER(ImageDesc), sz, vp, sz, vp, sz, C.POINTER(sz),
                                              C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
        L.jpgpu_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
        L.jpgpu_multi_destroy.argtypes = [vp]
        L.jpgpu_multi_destroy.restype = None
        L.jpgpu_multi_device_count.argtypes = [vp]
        L.jpgpu_partition.argtypes = [C.POINTER(ImageDesc), sz, sz, C.POINTER(sz)]
        L.jpgpu_multi_plan.argtypes = [vp, C.POINTER(ImageDesc), sz]
        L.jpgpu_multi_range.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(sz), C.POINTER(sz)]
        for name in ("upload", "decode", "sync"):
            getattr(L, "jpgpu_multi_" + name).argtypes = [vp]
        L.jpgpu_multi_set_output_format.argtypes = [vp, C.c_uint32]
        L.jpgpu_multi_download.argtypes = [vp, C.POINTER(vp)]
        L.jpgpu_multi_results.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
        L.jpgpu_multi_device_rgb.restype = vp
        L.jpgpu_multi_device_rgb.argtypes = [vp, sz, C.POINTER(C.c_int), C.POINTER(sz)]
        L.jpgpu_multi_coefficients.argtypes = [vp, sz, vp, sz, C.POINTER(C.c_uint32)]
        L.jpgpu_multi_launch_count.restype = C.c_uint64
        L.jpgpu_multi_launch_count.argtypes = [vp]
        L.jpgpu_multi_time_decode.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
        L.jpgpu_multi_decode_batch.argtypes = [vp, C.POINTER(ImageDesc), sz, C.POINTER(vp), C.POINTER(C.c_int32),
                                               C.POINTER(C.c_uint64), C.c_uint32]
        _lib = L
    return _lib


def status_string(status):
    return lib().jpgpu_status_string(int(status)).decode()


def check(status, what=""):
    if status != OK:
        raise JpgpuError(status, what)


EXPORTED_SYMBOLS = [
    "jpgpu_parse", "jpgpu_parse_scans", "jpgpu_geometry", "jpgpu_status_string", "jpgpu_panic_message", "jpgpu_abi_version",
    "jpgpu_plan_info",
    "jpgpu_create", "jpgpu_destroy", "jpgpu_last_error", "jpgpu_set_stream", "jpgpu_sync",
    "jpgpu_decode", "jpgpu_decode_file",
    "jpgpu_batch_create", "jpgpu_batch_replan", "jpgpu_batch_destroy", "jpgpu_batch_upload", "jpgpu_batch_set_device_scans", "jpgpu_batch_set_device_output",
    "jpgpu_batch_set_output_format", "jpgpu_batch_set_normalisation",
    "jpgpu_batch_entropy", "jpgpu_batch_idct", "jpgpu_batch_decode", "jpgpu_batch_download",
    "jpgpu_batch_device_rgb", "jpgpu_batch_output_bytes", "jpgpu_batch_rgb_offset", "jpgpu_batch_results", "jpgpu_batch_coefficients", "jpgpu_batch_stats",
    "jpgpu_batch_profile", "jpgpu_batch_launch_count",
    "jpgpu_stream", "jpgpu_batch_upload_from", "jpgpu_batch_download_contiguous",
    "jpgpu_pipeline_create", "jpgpu_pipeline_destroy", "jpgpu_pipeline_output_bytes", "jpgpu_pipeline_image_offset",
    "jpgpu_pipeline_run", "jpgpu_pipeline_sync", "jpgpu_pipeline_elapsed_ms", "jpgpu_pipeline_results",
    "jpgpu_pipeline_launch_count", "jpgpu_pipeline_last_error", "jpgpu_decode_batch_host",
    "jpgpu_multi_create", "jpgpu_multi_destroy", "jpgpu_multi_device_count", "jpgpu_partition", "jpgpu_multi_plan", "jpgpu_multi_range",
    "jpgpu_multi_upload", "jpgpu_multi_decode", "jpgpu_multi_sync", "jpgpu_multi_set_output_format", "jpgpu_multi_download",
    "jpgpu_multi_results", "jpgpu_multi_device_rgb", "jpgpu_multi_coefficients", "jpgpu_multi_launch_count",
    "jpgpu_multi_time_decode", "jpgpu_multi_decode_batch",
]
