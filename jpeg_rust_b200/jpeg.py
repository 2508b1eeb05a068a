"""Host-side mirror of the reference's public interface for the decode path.

Reference (martinhath/jpeg-rust)            here
-------------------------------------------  -----------------------------------------
jpeg::JPEGImage::parse(Vec<u8>)  mod.rs:202   JPEGImage.parse(bytes)
  .width() .height() .image_data() 467-477    same names
jpeg::decoder::JPEGDecoder       decoder.rs:19 JPEGDecoder (same builder calls, same order)
  ::new / .frame_header / .scan_header /
  .dimensions / .huffman_ac_tables /
  .huffman_dc_tables / .quantization_table /
  .decode() -> (Vec<(u8,u8,u8)>, usize)
jpeg::huffman::HuffmanTable
  ::from_size_data_tables        huffman.rs:37 HuffmanTable.from_size_data_tables

Every reference panic becomes a `JPEGPanic` carrying the status code of include/jpgpu.h.
The image data never passes through Python arithmetic: parsing fills a C descriptor
(jpgpu_parse) and decode() calls the CUDA library through the C ABI.  Without the
library or without a B200 these calls raise — there is no CPU path.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import _ffi
from ._ffi import EXT_DRI, EXT_MULTISCAN, EXT_NONE, EXT_SKIP_APPN, LAYOUT_REF, LAYOUT_SPEC, LAYOUT_SPEC_FANCY  # noqa: F401


class JPEGPanic(_ffi.JpgpuError):
    """The reference would have panicked (or this library rejected the input)."""


def _check(status, what=""):
    if status != _ffi.OK:
        raise JPEGPanic(status, what)


# ----------------------------------------------------------------------------- device context
_contexts = {}


class Context:
    """One jpgpu_ctx per (process, device)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(_ffi.lib().jpgpu_create(device, C.byref(self._h)), f"jpgpu_create(device={device})")
        self.device = device

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream):
        _check(_ffi.lib().jpgpu_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def sync(self):
        self._ck(_ffi.lib().jpgpu_sync(self._h))

    def _ck(self, status, what=""):
        if status in (_ffi.ERR_CUDA, _ffi.ERR_OOM):
            what = (what + " " if what else "") + _ffi.lib().jpgpu_last_error(self._h).decode()
        _check(status, what)

    def close(self):
        if self._h:
            _ffi.lib().jpgpu_destroy(self._h)
            self._h = C.c_void_p()


def context(device=0):
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]


# ----------------------------------------------------------------------------- header types (mod.rs:89-139)
@dataclass
class FrameComponentHeader:
    component_id: int
    horizontal_sampling_factor: int
    vertical_sampling_factor: int
    quantization_selector: int


@dataclass
class FrameHeader:
    sample_precision: int = 8
    num_lines: int = 0
    samples_per_line: int = 0
    image_components: int = 0
    frame_components: List[FrameComponentHeader] = field(default_factory=list)


@dataclass
class ScanComponentHeader:
    component_id: int
    dc_table_selector: int
    ac_table_selector: int


@dataclass
class ScanHeader:
    num_components: int = 0
    scan_components: List[ScanComponentHeader] = field(default_factory=list)
    start_spectral_selection: int = 0
    end_spectral_selection: int = 63
    successive_approximation_bit_pos_high: int = 0
    successive_approximation_bit_pos_low: int = 0


class HuffmanTable:
    """huffman.rs:24-58: built from DHT BITS (`size_data`) and HUFFVAL (`data_table`)."""

    def __init__(self, size_data, data_table):
        self.size_data = bytes(size_data)
        self.data_table = bytes(data_table)
        if len(self.size_data) != 16:
            raise ValueError("size_data must have 16 entries")

    @staticmethod
    def from_size_data_tables(size_data, data_table):
        return HuffmanTable(size_data, data_table)


# ----------------------------------------------------------------------------- JPEGDecoder (decoder.rs:19-343)
class JPEGDecoder:
    """Builder with the reference's call sequence (mod.rs:388-415).

    `data` is the RAW entropy-coded segment (still byte-stuffed, from the first byte
    after the SOS header to the end of the file); the unstuffing loop of mod.rs:371-385
    runs on the GPU together with the decode.
    """

    def __init__(self, data, layout=LAYOUT_REF, restart_interval=0, device=0):
        self._data = np.frombuffer(bytes(data), np.uint8).copy() if not isinstance(data, np.ndarray) else data
        self._frame: Optional[FrameHeader] = None
        self._scan: Optional[ScanHeader] = None
        self._dims: Tuple[int, int] = (0, 0)
        self._ac = {}
        self._dc = {}
        self._qt = {}
        self.layout = layout
        self.restart_interval = restart_interval
        self.device = device

    @staticmethod
    def new(data, **kw):
        return JPEGDecoder(data, **kw)

    def frame_header(self, frame_header: FrameHeader):      # decoder.rs:83
        self._frame = frame_header
        return self

    def scan_header(self, scan_header: ScanHeader):         # decoder.rs:113
        self._scan = scan_header
        return self

    def dimensions(self, dimensions):                       # decoder.rs:66
        self._dims = (int(dimensions[0]), int(dimensions[1]))
        return self

    def huffman_ac_tables(self, id, table: HuffmanTable):   # decoder.rs:71
        self._ac[int(id)] = table

    def huffman_dc_tables(self, id, table: HuffmanTable):   # decoder.rs:75
        self._dc[int(id)] = table

    def quantization_table(self, id, table):                # decoder.rs:79 (64 entries, zigzag order)
        self._qt[int(id)] = [int(x) for x in table]

    def descriptor(self) -> _ffi.ImageDesc:
        """The POD the C ABI takes; component order = scan order (decoder.rs:141-150)."""
        if self._frame is None:
            raise JPEGPanic(_ffi.PANIC_NO_FRAME_HEADER)
        if self._scan is None:
            raise JPEGPanic(_ffi.NO_SCAN)
        d = _ffi.ImageDesc()
        d.width, d.height = self._dims
        d.layout = self.layout
        d.restart_interval = self.restart_interval
        comps = self._scan.scan_components
        if len(comps) > 4:
            raise JPEGPanic(_ffi.PANIC_COMPONENT_COUNT)
        d.ncomp = len(comps)
        for i, sc in enumerate(comps):
            fc = None
            for k in self._frame.frame_components:  # decoder.rs:86-95: later entries with the same id overwrite
                if k.component_id == sc.component_id:
                    fc = k
            if fc is None:
                raise JPEGPanic(_ffi.PANIC_ARITH)
            d.comp[i] = _ffi.Component(fc.component_id, fc.horizontal_sampling_factor, fc.vertical_sampling_factor,
                                       fc.quantization_selector, sc.dc_table_selector, sc.ac_table_selector)
        for tid, q in self._qt.items():
            if not 0 <= tid < 4 or len(q) != 64:
                raise JPEGPanic(_ffi.PANIC_INDEX_OOB)
            for k in range(64):
                d.qt[tid][k] = q[k]
            d.qt_present[tid] = 1
        for tabs, bits, vals, nvals, present in ((self._dc, d.dc_bits, d.dc_vals, d.dc_nvals, d.dc_present),
                                                 (self._ac, d.ac_bits, d.ac_vals, d.ac_nvals, d.ac_present)):
            for tid, t in tabs.items():
                if not 0 <= tid < 4:
                    raise JPEGPanic(_ffi.PANIC_INDEX_OOB)
                if len(t.data_table) > 256:
                    raise JPEGPanic(_ffi.ERR_BAD_HUFFMAN_TABLE)
                for k in range(16):
                    bits[tid][k] = t.size_data[k]
                for k, v in enumerate(t.data_table):
                    vals[tid][k] = v
                nvals[tid] = len(t.data_table)
                present[tid] = 1
        d.scan = self._data.ctypes.data
        d.scan_len = self._data.size
        return d

    def decode(self):
        """decoder.rs:162: returns (pixels as an (H*W, 3) uint8 array, bytes_read)."""
        d = self.descriptor()
        ctx = context(self.device)
        out = np.empty((d.height * d.width, 3), np.uint8)
        br = C.c_size_t(0)
        ctx._ck(_ffi.lib().jpgpu_decode(ctx.handle, C.byref(d), out.ctypes.data, C.byref(br)), "jpgpu_decode")
        return out, br.value


# ----------------------------------------------------------------------------- JPEGImage (mod.rs:59-87, 202-477)
def parse_descriptor(data, ext=EXT_NONE, layout=LAYOUT_REF):
    """jpgpu_parse: returns (status, ImageDesc, buffer that owns the bytes the descriptor points into)."""
    buf = data if isinstance(data, np.ndarray) else np.frombuffer(bytes(data), np.uint8).copy()
    d = _ffi.ImageDesc()
    st = _ffi.lib().jpgpu_parse(buf.ctypes.data, buf.size, ext, layout, C.byref(d))
    return st, d, buf


def parse_scans(data, ext=EXT_NONE, layout=LAYOUT_SPEC):
    """jpgpu_parse_scans: one descriptor per scan of the file (ext gets EXT_MULTISCAN).  A file of non-interleaved scans
    gives the consecutive part descriptors of one frame; the ordinary file gives one descriptor.
    Returns (status, list of ImageDesc, buffer that owns the bytes)."""
    buf = data if isinstance(data, np.ndarray) else np.frombuffer(bytes(data), np.uint8).copy()
    out = (_ffi.ImageDesc * 4)()
    n = C.c_size_t(0)
    st = _ffi.lib().jpgpu_parse_scans(buf.ctypes.data, buf.size, ext | _ffi.EXT_MULTISCAN, layout, out, 4, C.byref(n))
    descs = []
    for k in range(n.value):
        d = _ffi.ImageDesc()
        C.memmove(C.byref(d), C.byref(out[k]), C.sizeof(_ffi.ImageDesc))
        descs.append(d)
    return st, descs, buf


def plan_info(files=None, descs=None, ext=EXT_NONE, layout=LAYOUT_SPEC, copies=1):
    """What the planner decides for a batch (jpgpu_plan_info; host only, needs no GPU): dict with sub_bits,
    lookback_bits, seg_bits, write_parts, groups, interval_images, warp_jobs, device_bytes.  `copies` repeats the
    given images (descriptors pointing at the same bytes) to ask about a large batch without building one."""
    keep = []
    if descs is None:
        ds = []
        for f in files:
            st, d, buf = parse_descriptor(f, ext, layout)
            _check(st, "jpgpu_parse")
            ds.append(d)
            keep.append(buf)
        descs = ds
    descs = list(descs) * copies
    arr = (_ffi.ImageDesc * len(descs))(*descs)
    info = (C.c_uint64 * 8)()
    _check(_ffi.lib().jpgpu_plan_info(arr, len(descs), info), "jpgpu_plan_info")
    names = ["sub_bits", "lookback_bits", "seg_bits", "write_parts", "groups", "interval_images", "warp_jobs", "device_bytes"]
    return {k: int(info[i]) for i, k in enumerate(names)}


class JPEGImage:
    """Result of JPEGImage::parse (mod.rs:202): dimensions and decoded pixels."""

    def __init__(self):
        self._dimensions = (0, 0)
        self._image_data = None
        self.bytes_read = 0
        self.descriptor = None

    @staticmethod
    def parse(vec, ext=EXT_NONE, layout=LAYOUT_REF, device=0):
        """mod.rs:202-465. Raises JPEGPanic where the reference panics; like the reference it
        decodes the first scan and returns - unless ext holds EXT_MULTISCAN (a file of non-interleaved scans is then
        decoded whole, jpgpu_decode_file)."""
        if ext & EXT_MULTISCAN:
            buf = vec if isinstance(vec, np.ndarray) else np.frombuffer(bytes(vec), np.uint8).copy()
            _st, ds, _ = parse_scans(buf, ext, layout)
            _check(_st, "parse")
            d0 = ds[0]
            w, h = (d0.frame_width, d0.frame_height) if d0.frame_part else (d0.width, d0.height)
            img = JPEGImage()
            img._dimensions = (w, h)
            img.descriptor = d0
            img._buf = buf
            ctx = context(device)
            out = np.empty((h * w, 3), np.uint8)
            br, ww, hh = C.c_size_t(0), C.c_uint32(0), C.c_uint32(0)
            ctx._ck(_ffi.lib().jpgpu_decode_file(ctx.handle, buf.ctypes.data, buf.size, ext, layout, out.ctypes.data, out.size,
                                                 C.byref(ww), C.byref(hh), C.byref(br)), "jpgpu_decode_file")
            img._image_data = out
            img.bytes_read = br.value
            return img
        st, d, buf = parse_descriptor(vec, ext, layout)
        _check(st, "parse")
        img = JPEGImage()
        img._dimensions = (d.width, d.height)
        img.descriptor = d
        img._buf = buf
        ctx = context(device)
        out = np.empty((d.height * d.width, 3), np.uint8)
        br = C.c_size_t(0)
        ctx._ck(_ffi.lib().jpgpu_decode(ctx.handle, C.byref(d), out.ctypes.data, C.byref(br)), "jpgpu_decode")
        img._image_data = out
        img.bytes_read = br.value
        return img

    def width(self):           # mod.rs:467
        return self._dimensions[0]

    def height(self):          # mod.rs:471
        return self._dimensions[1]

    def image_data(self):      # mod.rs:475: row-major (r, g, b) triples
        return self._image_data

    def rgb(self):
        return self._image_data.reshape(self.height(), self.width(), 3)

    def write_ppm(self, path):
        """main.rs:34-39: ASCII PPM (P3), one `r g b` line per pixel."""
        with open(path, "w") as f:
            f.write(f"P3\n{self.width()} {self.height()}\n255\n")
            for r, g, b in self._image_data:
                f.write(f"{r} {g} {b}\n")


# ----------------------------------------------------------------------------- batches
class Batch:
    """Many independent images decoded with shared kernel launches (jpgpu_batch_*)."""

    def __init__(self, files=None, descs=None, ext=EXT_NONE, layout=LAYOUT_SPEC, device=0, keepalive=None, ctx=None):
        L = _ffi.lib()
        self.ctx = ctx if ctx is not None else context(device)  # pass an own Context to overlap batches on its stream
        self._keep = [keepalive]
        if descs is None:
            n = len(files)
            descs = (_ffi.ImageDesc * n)()
            self.parse_status = []
            for i, f in enumerate(files):
                st, d, buf = parse_descriptor(f, ext, layout)
                self.parse_status.append(st)
                descs[i] = d
                self._keep.append(buf)
        else:
            n = len(descs)
            self.parse_status = [0] * n
        self.n = n
        self.descs = descs
        self._h = C.c_void_p()
        self.ctx._ck(L.jpgpu_batch_create(self.ctx.handle, descs, n, C.byref(self._h)), "jpgpu_batch_create")

    def replan(self, files=None, descs=None, ext=EXT_NONE, layout=LAYOUT_SPEC, keepalive=None):
        """Plan another wave of images on this batch object (device arenas are reused and only grow)."""
        self._keep = [keepalive]
        if descs is None:
            n = len(files)
            descs = (_ffi.ImageDesc * n)()
            self.parse_status = []
            for i, f in enumerate(files):
                st, d, buf = parse_descriptor(f, ext, layout)
                self.parse_status.append(st)
                descs[i] = d
                self._keep.append(buf)
        else:
            n = len(descs)
            self.parse_status = [0] * n
        self.n = n
        self.descs = descs
        self.ctx._ck(_ffi.lib().jpgpu_batch_replan(self._h, descs, n), "jpgpu_batch_replan")
        return self

    def _call(self, name):
        self.ctx._ck(getattr(_ffi.lib(), "jpgpu_batch_" + name)(self._h), "jpgpu_batch_" + name)
        return self

    def upload(self):
        return self._call("upload")

    def entropy(self):
        return self._call("entropy")

    def idct(self):
        return self._call("idct")

    def decode(self):
        return self._call("decode")

    def set_output_format(self, fmt):
        """OUT_RGB_INTERLEAVED (default; the reference's Vec<(u8,u8,u8)>, arrays of shape (H, W, 3)) or
        OUT_RGB_PLANAR (three W x H planes R, G, B: arrays / tensors of shape (3, H, W)) or OUT_F32_PLANAR (the same
        planes as float32, sample * scale + bias, see set_normalisation) for the following idct / decode calls
        (SURVEY.md §8(f) row 2)."""
        self.ctx._ck(_ffi.lib().jpgpu_batch_set_output_format(self._h, int(fmt)), "jpgpu_batch_set_output_format")
        self.out_format = int(fmt)
        return self

    def set_normalisation(self, scale, bias):
        """OUT_F32_PLANAR: out = sample * scale[c] + bias[c] per channel (default 1/255, 0)."""
        sc = (C.c_float * 3)(*[float(x) for x in scale])
        bi = (C.c_float * 3)(*[float(x) for x in bias])
        self.ctx._ck(_ffi.lib().jpgpu_batch_set_normalisation(self._h, sc, bi), "jpgpu_batch_set_normalisation")
        return self

    def dtype(self):
        return np.float32 if getattr(self, "out_format", 0) == _ffi.OUT_F32_PLANAR else np.uint8

    def shape(self, i):
        if getattr(self, "out_format", _ffi.OUT_RGB_INTERLEAVED) != _ffi.OUT_RGB_INTERLEAVED:
            return (3, self.descs[i].height, self.descs[i].width)
        return (self.descs[i].height, self.descs[i].width, 3)

    def download(self, outs=None):
        """Device->host copy of every image (async on the context stream). Returns the list of arrays."""
        if outs is None:
            outs = [np.empty(self.shape(i), self.dtype()) for i in range(self.n)]
        ptrs = (C.c_void_p * self.n)(*[o.ctypes.data if hasattr(o, "ctypes") else int(o) for o in outs])
        self.ctx._ck(_ffi.lib().jpgpu_batch_download(self._h, ptrs), "jpgpu_batch_download")
        return outs

    def upload_from(self, buf):
        """All scans lie in the one host buffer `buf` (numpy uint8, ideally pinned): a single host->device copy."""
        self.ctx._ck(_ffi.lib().jpgpu_batch_upload_from(self._h, buf.ctypes.data, buf.size), "jpgpu_batch_upload_from")
        return self

    def download_contiguous(self, out):
        """One device->host copy of the whole output arena into `out` (numpy uint8 of output_bytes()); image i at rgb_offset(i)."""
        self.ctx._ck(_ffi.lib().jpgpu_batch_download_contiguous(self._h, out.ctypes.data, out.size), "jpgpu_batch_download_contiguous")
        return self

    def download_ptrs(self, ptrs):
        arr = (C.c_void_p * self.n)(*[int(p) for p in ptrs])
        self.ctx._ck(_ffi.lib().jpgpu_batch_download(self._h, arr), "jpgpu_batch_download")

    def results(self):
        """Synchronises. Returns (statuses, bytes_read) as int lists; parse failures take precedence."""
        st = (C.c_int32 * self.n)()
        br = (C.c_uint64 * self.n)()
        self.ctx._ck(_ffi.lib().jpgpu_batch_results(self._h, st, br), "jpgpu_batch_results")
        statuses = [self.parse_status[i] if self.parse_status[i] else st[i] for i in range(self.n)]
        return statuses, list(br)

    def set_device_scans(self, dev_base, offsets):
        """The raw scan bytes are already in device memory: image i at dev_base + offsets[i]."""
        offs = (C.c_uint64 * self.n)(*[int(o) for o in offsets])
        self.ctx._ck(_ffi.lib().jpgpu_batch_set_device_scans(self._h, C.c_void_p(int(dev_base)), offs), "set_device_scans")
        return self

    def set_device_output(self, dev_base, capacity):
        """Wave decoding: RGB of the next decode goes to caller-owned device memory (None = own arena)."""
        self.ctx._ck(_ffi.lib().jpgpu_batch_set_device_output(self._h, C.c_void_p(int(dev_base) if dev_base else 0), capacity),
                     "set_device_output")
        return self

    def output_bytes(self):
        """Bytes one wave of RGB output occupies (images at 256-byte aligned offsets; failed images take none)."""
        return int(_ffi.lib().jpgpu_batch_output_bytes(self._h))

    def rgb_offset(self, i):
        """(offset of image i inside the wave's output arena, its size in bytes; size 0 for a failed image)."""
        off, nb = C.c_size_t(0), C.c_size_t(0)
        _check(_ffi.lib().jpgpu_batch_rgb_offset(self._h, i, C.byref(off), C.byref(nb)), "jpgpu_batch_rgb_offset")
        return off.value, nb.value

    def device_tensor(self, i):
        """Zero-copy torch view of image i's RGB output in device memory: uint8, (H, W, 3) or, with
        OUT_RGB_PLANAR, (3, H, W)."""
        import torch
        ptr, nb = self.device_rgb(i)
        shp = tuple(self.shape(i))

        ts = "<f4" if self.dtype() == np.float32 else "|u1"

        class _Cai:
            __cuda_array_interface__ = {"shape": shp, "typestr": ts, "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(_Cai(), device=f"cuda:{self.ctx.device}")

    def device_rgb(self, i):
        nb = C.c_size_t(0)
        p = _ffi.lib().jpgpu_batch_device_rgb(self._h, i, C.byref(nb))
        return p, nb.value

    def coefficients(self, i):
        """Reference arrangement (decoder.rs:208-212): list of (nblocks, 64) int16 per component, zigzag, absolute DC."""
        d = self.descs[i]
        cap = ((d.width + 15) // 16 + 1) * ((d.height + 15) // 16 + 1) * 12 * 64
        out = np.zeros(cap, np.int16)
        nb = (C.c_uint32 * 4)()
        self.ctx._ck(_ffi.lib().jpgpu_batch_coefficients(self._h, i, out.ctypes.data, cap, nb), "coefficients")
        comps, off = [], 0
        for c in range(d.ncomp):
            comps.append(out[off:off + nb[c] * 64].reshape(-1, 64).copy())
            off += nb[c] * 64
        return comps

    def stats(self):
        s = (C.c_uint64 * 8)()
        _check(_ffi.lib().jpgpu_batch_stats(self._h, s))
        return {"scan_bytes": s[0], "coef_bytes": s[1], "rgb_bytes": s[2], "pixels": s[3], "blocks": s[4],
                "sequences": s[5], "subsequences": s[6], "device_bytes": s[7]}

    def profile(self):
        """Device time (ms) of every kernel of one decode, run one after the other (jpgpu_batch_profile)."""
        ms = (C.c_float * 8)()
        self.ctx._ck(_ffi.lib().jpgpu_batch_profile(self._h, ms), "jpgpu_batch_profile")
        names = ["prepass_count", "prepass_scan", "prepass_write", "sync", "verify_scan", "decode_write", "idct_colour"]
        return {n: float(ms[i]) for i, n in enumerate(names)}

    def launch_count(self):
        return int(_ffi.lib().jpgpu_batch_launch_count(self._h))

    def close(self):
        if self._h:
            _ffi.lib().jpgpu_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_files(files, pinned=True):
    """The files of a batch back to back (64-byte aligned) in ONE host buffer - pinned if torch with CUDA is there -
    so that the library moves them with one copy (jpgpu_batch_upload_from / jpgpu_pipeline_run).
    Returns (uint8 numpy view of the buffer, offsets, object that owns the memory)."""
    sizes = [len(f) for f in files]
    offs = np.zeros(len(files) + 1, np.int64)
    offs[1:] = np.cumsum([(s + 63) // 64 * 64 for s in sizes])
    owner = None
    if pinned:
        try:
            import torch
            owner = torch.empty(max(int(offs[-1]), 64), dtype=torch.uint8).pin_memory()
            buf = owner.numpy()
        except Exception:
            owner = None
    if owner is None:
        owner = buf = np.empty(max(int(offs[-1]), 64), np.uint8)
    for i, f in enumerate(files):
        buf[offs[i]:offs[i] + sizes[i]] = np.frombuffer(f, np.uint8)
    return buf, offs, owner


def parse_packed(buf, offs, sizes, ext=EXT_NONE, layout=LAYOUT_SPEC):
    """jpgpu_parse on every file of a packed buffer: (ImageDesc array pointing into buf, parse statuses)."""
    n = len(sizes)
    descs = (_ffi.ImageDesc * n)()
    statuses = []
    for i in range(n):
        st, d, _ = parse_descriptor(buf[offs[i]:offs[i] + sizes[i]], ext, layout)
        descs[i] = d
        statuses.append(st)
    return descs, statuses


class Pipeline:
    """Host files in, host pixels out through jpgpu_pipeline_* (chunks alternating between two stream sets; every
    transfer one copy).  `files` are packed into one pinned buffer; the output is one pinned buffer, image i at
    `offset(i)`."""

    def __init__(self, files=None, ext=EXT_NONE, layout=LAYOUT_SPEC, device=0, chunk=0, packed=None, descs=None):
        import torch
        L = _ffi.lib()
        if packed is None:
            self.buf, self.offs, self._owner = pack_files(files)
            self.descs, self.parse_status = parse_packed(self.buf, self.offs, [len(f) for f in files], ext, layout)
        else:   # (buf, owner, descs): already packed and parsed by the caller
            self.buf, self._owner = packed
            self.descs = descs
            self.parse_status = [0] * len(descs)
        self.n = len(self.descs)
        self.device = device
        self._h = C.c_void_p()
        _check(L.jpgpu_pipeline_create(device, self.descs, self.n, chunk, C.byref(self._h)), "jpgpu_pipeline_create")
        self.out_bytes = int(L.jpgpu_pipeline_output_bytes(self._h))
        self._out_owner = torch.empty(max(self.out_bytes, 256), dtype=torch.uint8).pin_memory()
        self.out = self._out_owner.numpy()

    def _ck(self, st, what):
        if st in (_ffi.ERR_CUDA, _ffi.ERR_OOM):
            what += " " + _ffi.lib().jpgpu_pipeline_last_error(self._h).decode()
        _check(st, what)

    def run(self):
        self._ck(_ffi.lib().jpgpu_pipeline_run(self._h, self.buf.ctypes.data, self.buf.size, self.out.ctypes.data, self.out.size),
                 "jpgpu_pipeline_run")
        return self

    def sync(self):
        self._ck(_ffi.lib().jpgpu_pipeline_sync(self._h), "jpgpu_pipeline_sync")
        return self

    def elapsed_ms(self):
        ms = C.c_float(0)
        self._ck(_ffi.lib().jpgpu_pipeline_elapsed_ms(self._h, C.byref(ms)), "jpgpu_pipeline_elapsed_ms")
        return float(ms.value)

    def offset(self, i):
        off, nb = C.c_size_t(0), C.c_size_t(0)
        _check(_ffi.lib().jpgpu_pipeline_image_offset(self._h, i, C.byref(off), C.byref(nb)))
        return off.value, nb.value

    def image(self, i):
        off, nb = self.offset(i)
        d = self.descs[i]
        return self.out[off:off + nb].reshape(d.height, d.width, 3)

    def results(self):
        st = (C.c_int32 * self.n)()
        br = (C.c_uint64 * self.n)()
        self._ck(_ffi.lib().jpgpu_pipeline_results(self._h, st, br), "jpgpu_pipeline_results")
        return [self.parse_status[i] if self.parse_status[i] else st[i] for i in range(self.n)], list(br)

    def launch_count(self):
        return int(_ffi.lib().jpgpu_pipeline_launch_count(self._h))

    def close(self):
        if self._h:
            _ffi.lib().jpgpu_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiDevice:
    """One process, several GPUs (jpgpu_multi_*): contiguous image ranges of about equal scan bytes, one per device,
    each with its own context, streams and worker thread; no communication between devices."""

    def __init__(self, devices):
        devs = (C.c_int * len(devices))(*devices)
        self._h = C.c_void_p()
        _check(_ffi.lib().jpgpu_multi_create(devs, len(devices), C.byref(self._h)), "jpgpu_multi_create")
        self.devices = list(devices)
        self.n = 0
        self._keep = None

    def plan(self, files=None, descs=None, ext=EXT_NONE, layout=LAYOUT_SPEC, keepalive=None):
        if descs is None:
            self._keep = []
            descs = (_ffi.ImageDesc * len(files))()
            self.parse_status = []
            for i, f in enumerate(files):
                st, d, buf = parse_descriptor(f, ext, layout)
                descs[i] = d
                self.parse_status.append(st)
                self._keep.append(buf)
        else:
            self._keep = keepalive
            self.parse_status = [0] * len(descs)
        self.descs = descs
        self.n = len(descs)
        _check(_ffi.lib().jpgpu_multi_plan(self._h, descs, self.n), "jpgpu_multi_plan")
        return self

    def ranges(self):
        out = []
        for k in range(len(self.devices)):
            dev, first, count = C.c_int(0), C.c_size_t(0), C.c_size_t(0)
            _check(_ffi.lib().jpgpu_multi_range(self._h, k, C.byref(dev), C.byref(first), C.byref(count)))
            out.append((dev.value, first.value, count.value))
        return out

    def upload(self):
        _check(_ffi.lib().jpgpu_multi_upload(self._h), "jpgpu_multi_upload")
        return self

    def decode(self):
        _check(_ffi.lib().jpgpu_multi_decode(self._h), "jpgpu_multi_decode")
        return self

    def sync(self):
        _check(_ffi.lib().jpgpu_multi_sync(self._h), "jpgpu_multi_sync")
        return self

    def download(self):
        outs = [np.empty((self.descs[i].height, self.descs[i].width, 3), np.uint8) for i in range(self.n)]
        ptrs = (C.c_void_p * self.n)(*[o.ctypes.data for o in outs])
        _check(_ffi.lib().jpgpu_multi_download(self._h, ptrs), "jpgpu_multi_download")
        return outs

    def results(self):
        st = (C.c_int32 * self.n)()
        br = (C.c_uint64 * self.n)()
        _check(_ffi.lib().jpgpu_multi_results(self._h, st, br), "jpgpu_multi_results")
        return [self.parse_status[i] if self.parse_status[i] else st[i] for i in range(self.n)], list(br)

    def coefficients(self, i):
        d = self.descs[i]
        cap = ((d.width + 15) // 16 + 1) * ((d.height + 15) // 16 + 1) * 12 * 64
        out = np.zeros(cap, np.int16)
        nb = (C.c_uint32 * 4)()
        _check(_ffi.lib().jpgpu_multi_coefficients(self._h, i, out.ctypes.data, cap, nb), "jpgpu_multi_coefficients")
        comps, off = [], 0
        for c in range(d.ncomp):
            comps.append(out[off:off + nb[c] * 64].reshape(-1, 64).copy())
            off += nb[c] * 64
        return comps

    def time_decode(self, steps):
        ms = (C.c_float * len(self.devices))()
        _check(_ffi.lib().jpgpu_multi_time_decode(self._h, steps, ms), "jpgpu_multi_time_decode")
        return [float(x) for x in ms]

    def launch_count(self):
        return int(_ffi.lib().jpgpu_multi_launch_count(self._h))

    def decode_batch(self, files, ext=EXT_NONE, layout=LAYOUT_SPEC):
        """jpgpu_multi_decode_batch with host output: (list of HxWx3 arrays, statuses, bytes_read)."""
        bufs, pst = [], []
        descs = (_ffi.ImageDesc * len(files))()
        for i, f in enumerate(files):
            st, d, buf = parse_descriptor(f, ext, layout)
            descs[i] = d
            pst.append(st)
            bufs.append(buf)
        n = len(files)
        outs = [np.zeros((max(1, descs[i].height), max(1, descs[i].width), 3), np.uint8) for i in range(n)]
        ptrs = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        st = (C.c_int32 * n)()
        br = (C.c_uint64 * n)()
        _check(_ffi.lib().jpgpu_multi_decode_batch(self._h, descs, n, ptrs, st, br, _ffi.MEMORY_HOST), "jpgpu_multi_decode_batch")
        return outs, [pst[i] if pst[i] else st[i] for i in range(n)], list(br)

    def close(self):
        if self._h:
            _ffi.lib().jpgpu_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def decode_scans(files, ext=EXT_NONE, layout=LAYOUT_SPEC, device=0, want_coefs=False):
    """Decode files that may hold one non-interleaved scan per component (jpgpu_parse_scans: a feature the reference
    lacks - it returns after the first scan, mod.rs:416-417).  Every scan becomes a descriptor of the one batch; a
    frame's pixels are the output of its first scan.  Returns (list of HxWx3 arrays or None, statuses, and - with
    want_coefs - per file the list of per-component (nblocks, 64) coefficient arrays in each scan's own block order)."""
    descs, owners, first, count, pstat = [], [], [], [], []
    for f in files:
        st, ds, buf = parse_scans(f, ext, layout)
        first.append(len(descs)); count.append(len(ds)); pstat.append(st)
        descs += ds; owners.append(buf)
    arr = (_ffi.ImageDesc * max(len(descs), 1))(*descs)
    b = Batch(descs=arr if descs else (_ffi.ImageDesc * 0)(), device=device, keepalive=owners)
    try:
        b.upload().decode()
        outs, ptrs = [], []
        for i, f in enumerate(files):
            d = descs[first[i]] if count[i] else None
            if d is None:
                outs.append(None)
                continue
            w, h = (d.frame_width, d.frame_height) if d.frame_part else (d.width, d.height)
            outs.append(np.zeros((h, w, 3), np.uint8))
        for k in range(len(descs)):
            owner = [i for i in range(len(files)) if first[i] == k and count[i]]
            ptrs.append(outs[owner[0]].ctypes.data if owner else 0)
        if descs:
            b.download_ptrs(ptrs)
        st, _ = b.results() if descs else ([], [])
        statuses = [pstat[i] if pstat[i] else st[first[i]] for i in range(len(files))]
        if not want_coefs:
            return outs, statuses
        coefs = []
        for i in range(len(files)):
            per = []
            for k in range(first[i], first[i] + count[i]):
                per += b.coefficients(k) if st[k] == 0 else [None]
            coefs.append(per)
        return outs, statuses, coefs
    finally:
        b.close()


def decode_batch(files, ext=EXT_NONE, layout=LAYOUT_SPEC, device=0):
    """Convenience: decode a list of JPEG byte strings; returns (list of HxWx3 arrays, statuses, bytes_read)."""
    b = Batch(files, ext=ext, layout=layout, device=device)
    try:
        b.upload().decode()
        outs = b.download()
        statuses, br = b.results()
        return outs, statuses, br
    finally:
        b.close()


def decode_waves(files, wave, ext=EXT_NONE, layout=LAYOUT_SPEC, device=0):
    """Decode a long list of JPEG byte strings in waves of `wave` images through ONE batch object (its bitstream and
    coefficient arenas are reused), keeping every RGB output resident in one device arena.
    Returns (list of zero-copy (H, W, 3) uint8 CUDA tensors, statuses, bytes_read)."""
    import torch
    descs_all, bufs, pstat = [], [], []
    for f in files:
        st, d, buf = parse_descriptor(f, ext, layout)
        descs_all.append(d); bufs.append(buf); pstat.append(st)
    sizes = [(d.width * d.height * 3 + 255) // 256 * 256 if st == 0 else 0 for d, st in zip(descs_all, pstat)]
    arena = torch.empty(sum(sizes) + 256, dtype=torch.uint8, device=f"cuda:{device}")
    base = (arena.data_ptr() + 255) // 256 * 256
    outs, statuses, bytes_read = [], [], []
    batch, off = None, 0
    for i0 in range(0, len(files), wave):
        ds = descs_all[i0:i0 + wave]
        arr = (_ffi.ImageDesc * len(ds))(*ds)
        if batch is None:
            batch = Batch(descs=arr, device=device, keepalive=bufs)
        else:
            batch.replan(descs=arr, keepalive=bufs)
        batch.parse_status = pstat[i0:i0 + wave]
        nbytes = batch.output_bytes()
        batch.set_device_output(base + off, max(nbytes, 256))
        batch.upload().decode()
        st, br = batch.results()
        statuses += st; bytes_read += br
        for i in range(len(ds)):
            outs.append(batch.device_tensor(i) if st[i] == 0 else None)
        off += (nbytes + 255) // 256 * 256
    if batch is not None:
        batch.set_device_output(None, 0)
        batch.close()
    for t in outs:
        if t is not None:
            t._jpgpu_arena = arena   # keep the arena alive as long as any view is
    return outs, statuses, bytes_read


def shard_range(n_items, rank, world_size):
    """Contiguous image range of `rank`: [rank*n/world, (rank+1)*n/world)."""
    return (rank * n_items) // world_size, ((rank + 1) * n_items) // world_size
