"""jpeg_rust_b200 — B200-native drop-in for the decode path of martinhath/jpeg-rust.

The package holds only what that path needs: the C ABI library (csrc/ -> lib/libjpgpu.so,
hand-written sm_100a kernels), its ctypes binding, a host-side mirror of the reference's
JPEGImage / JPEGDecoder interface, and the offline synthetic-input generator.
"""
from . import _ffi
from ._ffi import (EXT_DRI, EXT_MULTISCAN, EXT_NONE, EXT_SKIP_APPN, LAYOUT_REF, LAYOUT_SPEC, LAYOUT_SPEC_FANCY, OUT_RGB_INTERLEAVED, OUT_RGB_PLANAR, OUT_F32_PLANAR,
                   JpgpuError)
from .jpeg import (Batch, Context, MultiDevice, Pipeline, pack_files, parse_packed, FrameComponentHeader, FrameHeader, HuffmanTable, JPEGDecoder, JPEGImage,
                   JPEGPanic, ScanComponentHeader, ScanHeader, context, decode_batch, decode_scans, decode_waves, parse_descriptor, parse_scans,
                   plan_info, shard_range)

__all__ = ["_ffi", "EXT_DRI", "EXT_MULTISCAN", "EXT_NONE", "EXT_SKIP_APPN", "LAYOUT_REF", "LAYOUT_SPEC", "LAYOUT_SPEC_FANCY", "OUT_RGB_INTERLEAVED",
           "OUT_RGB_PLANAR", "OUT_F32_PLANAR", "JpgpuError", "Batch", "MultiDevice", "Pipeline", "pack_files", "parse_packed",
           "Context", "FrameComponentHeader", "FrameHeader", "HuffmanTable", "JPEGDecoder", "JPEGImage", "JPEGPanic",
           "ScanComponentHeader", "ScanHeader", "context", "decode_batch", "decode_scans", "decode_waves", "parse_descriptor", "parse_scans", "plan_info", "shard_range"]
