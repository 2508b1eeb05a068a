//! Binding of libjpgpu.so (include/jpgpu.h, ABI version 3).  Field order and types mirror the header;
//! every function names the reference interface it replaces.
#![allow(dead_code)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

/// One component in SCAN order (decoder.rs:39-52 after scan_header() reordering).
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct JpgpuComponent {
    pub id: u8,
    pub h: u8,
    pub v: u8,
    pub tq: u8,
    pub td: u8,
    pub ta: u8,
}

/// Everything mod.rs:388-413 hands to the JPEGDecoder builder (decoder.rs:55-152).
#[repr(C)]
pub struct JpgpuImageDesc {
    pub width: u32,
    pub height: u32,
    pub ncomp: u32,
    pub comp: [JpgpuComponent; 4],
    pub qt: [[u16; 64]; 4],
    pub qt_present: [u8; 4],
    pub dc_bits: [[u8; 16]; 4],
    pub dc_vals: [[u8; 256]; 4],
    pub dc_nvals: [u16; 4],
    pub dc_present: [u8; 4],
    pub ac_bits: [[u8; 16]; 4],
    pub ac_vals: [[u8; 256]; 4],
    pub ac_nvals: [u16; 4],
    pub ac_present: [u8; 4],
    pub restart_interval: u32,
    pub layout: u32,
    /// RAW (still byte-stuffed) bytes from the first byte after the SOS header to the end of the
    /// file: the range mod.rs:371-385 walks; the GPU removes the stuffing.
    pub scan: *const u8,
    pub scan_len: usize,
    /// Multi-scan files (jpgpu_parse_scans); all zero for the single scan the reference reads.
    pub frame_part: u32,
    pub frame_width: u32,
    pub frame_height: u32,
    pub frame_ncomp: u8,
    pub frame_comp: u8,
    pub frame_h: u8,
    pub frame_v: u8,
    pub frame_hmax: u8,
    pub frame_vmax: u8,
    pub frame_pad: [u8; 2],
}

pub const JPGPU_OK: c_int = 0;
pub const JPGPU_LAYOUT_REF: u32 = 0; // the reference's own placement, decoder.rs:259-312, 347-379
pub const JPGPU_LAYOUT_SPEC: u32 = 1; // T.81 A.2.3 geometry
pub const JPGPU_EXT_NONE: u32 = 0;
pub const JPGPU_MEMORY_HOST: u32 = 0;
pub const JPGPU_MEMORY_DEVICE: u32 = 1;

extern "C" {
    pub fn jpgpu_abi_version() -> c_int;
    pub fn jpgpu_create(device: c_int, out: *mut *mut c_void) -> c_int;
    pub fn jpgpu_destroy(ctx: *mut c_void);
    /// JPEGDecoder::decode(), decoder.rs:162 (+ the unstuffing loop mod.rs:371-385).
    pub fn jpgpu_decode(ctx: *mut c_void, desc: *const JpgpuImageDesc, rgb_out: *mut u8,
                        bytes_read: *mut usize) -> c_int;
    pub fn jpgpu_status_string(status: c_int) -> *const c_char;
    /// The reference's own panic text for statuses 1..15, NULL otherwise.
    pub fn jpgpu_panic_message(status: c_int) -> *const c_char;

    // Many files at once, one process, several GPUs (no counterpart in the reference).
    pub fn jpgpu_multi_create(devices: *const c_int, n_devices: c_int, out: *mut *mut c_void) -> c_int;
    pub fn jpgpu_multi_destroy(m: *mut c_void);
    pub fn jpgpu_multi_decode_batch(m: *mut c_void, descs: *const JpgpuImageDesc, n: usize,
                                    outs: *const *mut u8, statuses: *mut i32, bytes_read: *mut u64,
                                    memory_kind: u32) -> c_int;
}

/// Re-raises a non-zero status the way the reference would have failed: its own panic message
/// where it has one (mod.rs:258,427,446,449,457; huffman.rs:156,162,202; decoder.rs:148,224,330),
/// the library's description otherwise.
pub fn panic_with_status(status: c_int) -> ! {
    unsafe {
        let p = jpgpu_panic_message(status);
        let p = if p.is_null() { jpgpu_status_string(status) } else { p };
        panic!("{}", CStr::from_ptr(p).to_string_lossy());
    }
}

/// Process-wide context on device 0 (the reference is single-threaded and synchronous).
pub fn gpu_ctx() -> *mut c_void {
    use std::sync::{Once, ONCE_INIT};
    static INIT: Once = ONCE_INIT;
    static mut CTX: *mut c_void = 0 as *mut c_void;
    unsafe {
        INIT.call_once(|| {
            let st = jpgpu_create(0, &mut CTX);
            if st != JPGPU_OK {
                panic_with_status(st); // no B200: there is no CPU fallback
            }
        });
        CTX
    }
}
