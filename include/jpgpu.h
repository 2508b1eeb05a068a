/*
 * jpgpu.h — C ABI of the B200-native JPEG decode path (libjpgpu.so).
 *
 * Drop-in boundary: the reference (martinhath/jpeg-rust) has no FFI of its own;
 * the seam is the builder + call at src/jpeg/mod.rs:388-416:
 *
 *     JPEGDecoder::new(data)              decoder.rs:55
 *         .frame_header(..)               decoder.rs:83    per component id, H, V, Tq
 *         .scan_header(..)                decoder.rs:113   per component Td, Ta; scan order
 *         .dimensions((W, H))             decoder.rs:66
 *     .huffman_ac_tables / .huffman_dc_tables / .quantization_table   decoder.rs:71-81
 *     .decode() -> (Vec<(u8,u8,u8)>, usize)                           decoder.rs:162
 *
 * Everything those builder calls carry is one POD `jpgpu_image_desc`; `.decode()`
 * is `jpgpu_decode()` (one image, reference semantics) or the `jpgpu_batch_*`
 * family (many independent images, the path the benchmark measures).  Marker and
 * header parsing stays on the host (Rust in the reference, `jpgpu_parse()` here).
 * All pointers are plain host or device pointers; no torch/C++ types cross this
 * boundary.  There is no CPU fallback: every entry point that computes fails with
 * JPGPU_ERR_NO_DEVICE / JPGPU_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef JPGPU_H
#define JPGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JPGPU_ABI_VERSION 3

/* ------------------------------------------------------------------ statuses
 * 0 = success.  1..15 mirror the panics of the reference so that a host shim
 * can re-raise them (src/jpeg/mod.rs / decoder.rs / huffman.rs line numbers in
 * the comments).  >= 32 are errors of this library. */
enum {
    JPGPU_OK = 0,
    JPGPU_PANIC_UNHANDLED_MARKER = 1, /* mod.rs:457 */
    JPGPU_PANIC_DRI = 2,              /* mod.rs:427 "got to restart interval def" */
    JPGPU_PANIC_APP12_14 = 3,         /* mod.rs:446,449 */
    JPGPU_PANIC_DQT_PRECISION = 4,    /* mod.rs:258 */
    JPGPU_PANIC_SAMPLING_ASSERT = 5,  /* mod.rs:275-277 */
    JPGPU_PANIC_INDEX_OOB = 6,        /* slice index out of bounds anywhere in parse */
    JPGPU_PANIC_NO_FRAME_HEADER = 7,  /* mod.rs:388 */
    JPGPU_PANIC_MISSING_TABLE = 8,    /* decoder.rs:155,159,224 */
    JPGPU_PANIC_DC_LOOKUP = 9,        /* huffman.rs:156 */
    JPGPU_PANIC_AC_LOOKUP = 10,       /* huffman.rs:162 */
    JPGPU_PANIC_COMPONENT_COUNT = 11, /* decoder.rs:330 */
    JPGPU_PANIC_READ_BITS_ASSERT = 12,/* huffman.rs:202 */
    JPGPU_PANIC_SCAN_COMPONENT = 13,  /* decoder.rs:148 */
    JPGPU_NO_SCAN = 14,               /* parse() reached the end without SOS (image_data None) */
    JPGPU_PANIC_ARITH = 15,           /* debug-build arithmetic overflow, e.g. mod.rs:218 */

    JPGPU_ERR_INVALID_ARG = 32,
    JPGPU_ERR_NO_DEVICE = 33,         /* no CUDA device / not sm_100: there is no CPU fallback */
    JPGPU_ERR_CUDA = 34,              /* a CUDA runtime call failed; see jpgpu_last_error() */
    JPGPU_ERR_UNSUPPORTED = 35,       /* outside the decodable subset (e.g. H or V not in {1,2}) */
    JPGPU_ERR_BAD_HUFFMAN_TABLE = 36, /* BITS/HUFFVAL do not describe a prefix code */
    JPGPU_ERR_TRUNCATED = 37,         /* entropy data ended before all MCUs were decoded */
    JPGPU_ERR_BAD_CODE = 38,          /* bit pattern that is no code of the selected table
                                         (reference: huffman.rs:156/162 panic) */
    JPGPU_ERR_RESTART = 39,           /* RSTn markers missing / out of sequence */
    JPGPU_ERR_OOM = 40
};

/* Output geometry (SURVEY.md §8 "parity policy").
 * REF  = bug-compatible with the reference's placement (decoder.rs:259-312,
 *        347-379) and its MCU count (decoder.rs:191-192).
 * SPEC = what that code intends: T.81 A.2.3 MCU order, true MCU count, box
 *        replication of sub-sampled components, crop at right/bottom edge.
 * For gray, 4:4:4 with W%8==0 and 4:2:2 (H2V1) with W%16==0 both are identical. */
enum { JPGPU_LAYOUT_REF = 0, JPGPU_LAYOUT_SPEC = 1,
       /* SPEC geometry with libjpeg's "fancy" up-sampling of sub-sampled chroma (triangle filter: 3/4 of the nearer,
        * 1/4 of the farther sample, both directions for 4:2:0) instead of the reference's box replication.  A feature
        * the reference lacks: results differ from the reference by design; not the tuned path. */
       JPGPU_LAYOUT_SPEC_FANCY = 2 };

/* Sample arrangement of the output (the step after the path).  INTERLEAVED is the
 * reference's Vec<(u8,u8,u8)> (decoder.rs:162, 317-331): W*H triples, row-major.  PLANAR holds the same W*H*3
 * bytes as three W x H planes R, G, B (a CHW uint8 tensor).  Same values, same size, same per-image offsets. */
enum { JPGPU_OUT_RGB_INTERLEAVED = 0, JPGPU_OUT_RGB_PLANAR = 1,
       /* three W x H planes of float: the same 8-bit samples as sample * scale[c] + bias[c] (jpgpu_batch_set_normalisation;
        * default scale 1/255, bias 0: a CHW float32 tensor in [0, 1]).  Four bytes per sample: every output size and offset
        * (jpgpu_batch_output_bytes, jpgpu_batch_rgb_offset, jpgpu_batch_device_rgb, the download calls) is four times the
        * u8 one.  Converted inside the IDCT/colour kernel while a tile is copied out - no second pass over the pixels. */
       JPGPU_OUT_F32_PLANAR = 2 };

/* Parser extensions beyond the reference's accepted subset (bit flags). */
enum {
    JPGPU_EXT_NONE = 0,
    JPGPU_EXT_SKIP_APPN = 1, /* skip APPn/JPGn segments (incl. APP12/APP14) by their length */
    JPGPU_EXT_DRI = 2,       /* accept DRI; the scan is decoded per restart interval */
    JPGPU_EXT_MULTISCAN = 4  /* jpgpu_parse_scans only: go on after the first scan (non-interleaved component scans) */
};

/* One component, in SCAN order (decoder.rs:39-52, after scan_header() reordering). */
typedef struct jpgpu_component {
    uint8_t id; /* Ci */
    uint8_t h;  /* horizontal sampling factor, 1 or 2 (mod.rs:275) */
    uint8_t v;  /* vertical sampling factor, 1 or 2 (mod.rs:277) */
    uint8_t tq; /* quantisation table selector */
    uint8_t td; /* DC Huffman table selector */
    uint8_t ta; /* AC Huffman table selector */
} jpgpu_component;

/* Everything mod.rs:388-413 hands to the JPEGDecoder builder. */
typedef struct jpgpu_image_desc {
    uint32_t width, height;        /* SOF0 X, Y (mod.rs:295) */
    uint32_t ncomp;                /* components in the scan */
    jpgpu_component comp[4];
    uint16_t qt[4][64];            /* DQT entries in zigzag order, as in the file (mod.rs:238-256) */
    uint8_t qt_present[4];
    uint8_t dc_bits[4][16];        /* DHT BITS (mod.rs:313) */
    uint8_t dc_vals[4][256];       /* DHT HUFFVAL (mod.rs:321) */
    uint16_t dc_nvals[4];
    uint8_t dc_present[4];
    uint8_t ac_bits[4][16];
    uint8_t ac_vals[4][256];
    uint16_t ac_nvals[4];
    uint8_t ac_present[4];
    uint32_t restart_interval;     /* MCUs per interval; 0 = none (always 0 inside the reference's subset) */
    uint32_t layout;               /* JPGPU_LAYOUT_* */
    /* RAW (still byte-stuffed) entropy-coded bytes from the first byte after the
     * SOS header to the end of the file — the range mod.rs:371-385 walks.  The
     * GPU removes the stuffing and finds RSTn markers itself. */
    const uint8_t *scan;
    size_t scan_len;
    /* Multi-scan files (jpgpu_parse_scans, a feature the reference lacks: it returns after the first scan,
     * mod.rs:416-417).  Every non-interleaved scan is one descriptor - a one-component image of the component's own size
     * (width/height above = ceil(X*H/Hmax) x ceil(Y*V/Vmax), T.81 A.1.1) with the tables in force at its SOS - and the
     * descriptors of one frame are consecutive.  frame_part: 0 = an image of its own (everything jpgpu_parse fills in),
     * 1 = first scan of a frame (owns the frame's output), 2 = a further scan of the same frame. */
    uint32_t frame_part;
    uint32_t frame_width, frame_height; /* SOF0 X, Y */
    uint8_t frame_ncomp;                /* components of the frame (3) */
    uint8_t frame_comp;                 /* which one this scan carries, in SOF0 order (0 = Y, 1 = Cb, 2 = Cr) */
    uint8_t frame_h, frame_v;           /* its sampling factors */
    uint8_t frame_hmax, frame_vmax;     /* the largest of the frame */
    uint8_t frame_pad[2];
} jpgpu_image_desc;

typedef struct jpgpu_ctx jpgpu_ctx;     /* one per process and device */
typedef struct jpgpu_batch jpgpu_batch; /* a planned set of images with its device buffers */

/* ------------------------------------------------------------ host-only part */

/* Replaces JPEGImage::parse() up to the decode() call (mod.rs:202-414): walks
 * the markers, fills `out`, points out->scan into `file`.  Returns JPGPU_OK or
 * the status naming the panic the reference would raise.  `layout` is copied
 * into the descriptor.  Needs no GPU. */
int jpgpu_parse(const uint8_t *file, size_t len, uint32_t ext_flags, uint32_t layout, jpgpu_image_desc *out);

/* jpgpu_parse for files with several scans (ext_flags must hold JPGPU_EXT_MULTISCAN): walks the whole file and writes
 * one descriptor per scan into out[0 .. *n).  A single interleaved scan gives exactly what jpgpu_parse gives (*n = 1,
 * frame_part = 0).  Non-interleaved scans (one component each, all components of the frame present) give a frame of
 * consecutive part descriptors to hand to jpgpu_batch_create / jpgpu_multi_plan together; the frame's pixels are the
 * output of its first part.  Anything else - scans mixing interleaved and non-interleaved components, components coded
 * twice - is JPGPU_ERR_UNSUPPORTED.  Returns a status; *n is set in every case. */
int jpgpu_parse_scans(const uint8_t *file, size_t len, uint32_t ext_flags, uint32_t layout, jpgpu_image_desc *out,
                      size_t max_out, size_t *n);

/* Number of MCUs decode() reads and blocks per MCU for a descriptor (decoder.rs:164-192
 * for REF; true MCU count for SPEC).  Returns a status. */
int jpgpu_geometry(const jpgpu_image_desc *desc, uint32_t *mcus, uint32_t *blocks_per_mcu, uint32_t nblocks_per_comp[4]);

/* What the planner decides for these images (host only, needs no GPU): info[0] bits per subsequence (one decode thread
 * each), [1] look-back bits of the synchronisation pass, [2] checkpoint distance in bits, [3] write-pass units per
 * subsequence, [4] image groups of the pipelined decode, [5] images decoded from their restart-interval starts (no
 * synchronisation pass), [6] warp jobs (incl. padding), [7] device bytes of the arenas.  See DESIGN.md section 4. */
int jpgpu_plan_info(const jpgpu_image_desc *descs, size_t n, uint64_t info[8]);

const char *jpgpu_status_string(int status);
/* The literal part of the reference's own panic message for statuses 1..15 (e.g. "got to restart interval def",
 * mod.rs:427), NULL otherwise: what a strict drop-in shim hands to panic!() (rust/src/jpeg/ffi.rs). */
const char *jpgpu_panic_message(int status);
int jpgpu_abi_version(void);

/* ------------------------------------------------------------- device part */

/* Creates the context on CUDA device `device` (cudaSetDevice ordinal). */
int jpgpu_create(int device, jpgpu_ctx **out);
void jpgpu_destroy(jpgpu_ctx *ctx);
const char *jpgpu_last_error(const jpgpu_ctx *ctx);
/* All work of the context is enqueued on this cudaStream_t (default: a stream the
 * context owns).  Pass torch.cuda.current_stream().cuda_stream to time with torch events. */
int jpgpu_set_stream(jpgpu_ctx *ctx, void *cuda_stream);
int jpgpu_sync(jpgpu_ctx *ctx);

/* JPEGDecoder::decode() for one image (decoder.rs:162): writes W*H*3 interleaved
 * RGB bytes to the HOST buffer `rgb_out` and the reference's `bytes_read`. */
int jpgpu_decode(jpgpu_ctx *ctx, const jpgpu_image_desc *desc, uint8_t *rgb_out, size_t *bytes_read);

/* Convenience: parse + decode of a whole file (JPEGImage::parse, mod.rs:202). */
int jpgpu_decode_file(jpgpu_ctx *ctx, const uint8_t *file, size_t len, uint32_t ext_flags, uint32_t layout,
                      uint8_t *rgb_out, size_t rgb_cap, uint32_t *width, uint32_t *height, size_t *bytes_read);

/* Batch path (no counterpart in the reference; images are independent, so a batch
 * is simply many decode() calls sharing kernel launches).  Usage:
 *   create -> upload -> decode (= entropy + idct) -> download / device_rgb -> results.
 * A batch may be uploaded/decoded repeatedly (buffers are reused). */
int jpgpu_batch_create(jpgpu_ctx *ctx, const jpgpu_image_desc *descs, size_t n, jpgpu_batch **out);
void jpgpu_batch_destroy(jpgpu_batch *b);
/* Plans a different set of images on an existing batch object: device arenas are kept and only grow, so a long
 * job runs wave after wave through one batch (replan -> upload -> decode -> download).  Synchronises the stream. */
int jpgpu_batch_replan(jpgpu_batch *b, const jpgpu_image_desc *descs, size_t n);
/* Host->device copy of every image's scan bytes (async on the context stream).
 * desc.scan memory must stay valid until the stream has passed this point. */
int jpgpu_batch_upload(jpgpu_batch *b);
/* Alternative to upload: the raw scan bytes of all images are already in device
 * memory, image i at dev_base + offsets[i] (length = descs[i].scan_len). */
int jpgpu_batch_set_device_scans(jpgpu_batch *b, const void *dev_base, const uint64_t *offsets);
/* Wave decoding (16 384 x 1080p do not fit one GPU at once): direct the RGB output of the next
 * decode to caller-owned device memory instead of the batch's own arena — image i goes to dev_base + the same
 * 256-byte aligned offsets jpgpu_batch_device_rgb() reports relative to image 0.  One planned batch (coefficient
 * and bitstream arenas sized for one wave) then serves wave after wave: set_device_scans / upload, set_device_output,
 * decode.  NULL restores the batch's own arena (sized for the current plan).  `capacity` must be at least
 * jpgpu_batch_output_bytes(); it is remembered, and a later jpgpu_batch_replan() whose plan no longer fits it falls
 * back to the batch's own arena (check jpgpu_batch_device_rgb() or set the output again after a replan). */
int jpgpu_batch_set_device_output(jpgpu_batch *b, void *dev_base, size_t capacity);
/* Output arrangement (JPGPU_OUT_*) of the following idct / decode calls; kept across replans.  Default: interleaved. */
int jpgpu_batch_set_output_format(jpgpu_batch *b, uint32_t format);
/* JPGPU_OUT_F32_PLANAR: out = sample * scale[c] + bias[c] per channel c = R, G, B (e.g. 1/(255*std), -mean/std). */
int jpgpu_batch_set_normalisation(jpgpu_batch *b, const float scale[3], const float bias[3]);
int jpgpu_batch_entropy(jpgpu_batch *b); /* stage 1: unstuff/RST pre-pass, self-synchronising Huffman decode */
int jpgpu_batch_idct(jpgpu_batch *b);    /* stage 2+3: dequant, IDCT, upsample, YCbCr->RGB, interleaved store */
int jpgpu_batch_decode(jpgpu_batch *b);  /* entropy + idct */
/* Device->host copy of image i's RGB into outs[i] (host pointers, W*H*3 bytes each); async. */
int jpgpu_batch_download(jpgpu_batch *b, uint8_t *const *outs);
/* Device pointer / size of image i's interleaved RGB output. */
void *jpgpu_batch_device_rgb(jpgpu_batch *b, size_t i, size_t *nbytes);
/* Bytes the RGB outputs of the planned images occupy together (every image at a 256-byte aligned offset): the size a
 * caller-owned arena for jpgpu_batch_set_device_output() / jpgpu_batch_download_contiguous() needs. */
size_t jpgpu_batch_output_bytes(const jpgpu_batch *b);
/* Where image i's output lies inside that arena, and its size (0 for an image that failed to parse or plan; such an
 * image still has an offset - an empty slice - so one bad file never breaks the bookkeeping of a wave). */
int jpgpu_batch_rgb_offset(const jpgpu_batch *b, size_t i, size_t *offset, size_t *nbytes);
/* Synchronises the stream; per image: status (JPGPU_*) and the reference's bytes_read. Either may be NULL. */
int jpgpu_batch_results(jpgpu_batch *b, int32_t *statuses, uint64_t *bytes_read);
/* Debug export for the bit-exact gate: the decoded coefficients of image i in the
 * reference's own arrangement — blocks[component] after decoder.rs:208-212:
 * components in scan order, blocks in decode order, 64 x int16 in ZIGZAG order,
 * absolute DC.  `out` holds `cap` int16; nblocks[c] receives the block counts. Synchronises. */
int jpgpu_batch_coefficients(jpgpu_batch *b, size_t i, int16_t *out, size_t cap, uint32_t nblocks[4]);
/* Algorithmic byte counts of the last plan (for roofline arithmetic):
 * [0] raw scan bytes, [1] coefficient bytes (blocks*128), [2] RGB bytes, [3] pixels, [4] blocks. */
int jpgpu_batch_stats(jpgpu_batch *b, uint64_t stats[8]);
/* Measurement aid: runs the decode once, kernel after kernel on the context stream, with a CUDA event between
 * every two launches, and returns the device time of each in milliseconds:
 * [0] prepass_count [1] prepass_scan [2] prepass_write [3] sync [4] verify_scan [5] decode_write
 * [6] idct/colour (all sampling modes and the gather path together).  Synchronises. */
int jpgpu_batch_profile(jpgpu_batch *b, float ms[8]);
/* Number of kernel launches enqueued by this batch object so far. */
uint64_t jpgpu_batch_launch_count(const jpgpu_batch *b);

/* ------------------------------------------------------------------------------------------------------------------
 * Transfers as single copies, and the host-to-host pipeline (the step before and after the path). */

/* The cudaStream_t the context enqueues on. */
void *jpgpu_stream(const jpgpu_ctx *ctx);
/* Like jpgpu_batch_upload, for scans that all lie in ONE readable host buffer [host_base, host_base + host_bytes) -
 * the files of the batch back to back, ideally pinned: one host->device copy of the used span, then one kernel that
 * spreads it into the batch's raw arena. */
int jpgpu_batch_upload_from(jpgpu_batch *b, const void *host_base, size_t host_bytes);
/* Like jpgpu_batch_download, as ONE device->host copy of the whole output arena: image i lands at
 * host_base + jpgpu_batch_rgb_offset(i).  capacity >= jpgpu_batch_output_bytes(). */
int jpgpu_batch_download_contiguous(jpgpu_batch *b, void *host_base, size_t capacity);

/* Host buffers in, host buffers out, for many files at once: the images are cut into chunks of `chunk_images`
 * (0 = default) that alternate between two stream sets of `device`, so that the upload of one chunk, the kernels of
 * another and the download of a third overlap; every transfer is a single copy.  Plans and arenas are made at creation;
 * jpgpu_pipeline_run only enqueues.  The descriptors (tables, sizes, scan pointers) are fixed at creation, so a second
 * run decodes whatever those scan pointers hold then. */
typedef struct jpgpu_pipeline jpgpu_pipeline;
int jpgpu_pipeline_create(int device, const jpgpu_image_desc *descs, size_t n, size_t chunk_images, jpgpu_pipeline **out);
void jpgpu_pipeline_destroy(jpgpu_pipeline *p);
/* Size of the host output buffer, and where image i lies in it (nbytes = W*H*3, 0 for an image that failed to plan). */
size_t jpgpu_pipeline_output_bytes(const jpgpu_pipeline *p);
int jpgpu_pipeline_image_offset(const jpgpu_pipeline *p, size_t i, size_t *offset, size_t *nbytes);
/* Enqueue the whole job (asynchronous).  Every descs[i].scan must lie inside [host_in, host_in + host_in_bytes). */
int jpgpu_pipeline_run(jpgpu_pipeline *p, const void *host_in, size_t host_in_bytes, void *host_out, size_t host_out_bytes);
int jpgpu_pipeline_sync(jpgpu_pipeline *p);
/* Device time of the last run, first upload to last download (CUDA events).  Synchronises. */
int jpgpu_pipeline_elapsed_ms(jpgpu_pipeline *p, float *ms);
int jpgpu_pipeline_results(jpgpu_pipeline *p, int32_t *statuses, uint64_t *bytes_read);
uint64_t jpgpu_pipeline_launch_count(const jpgpu_pipeline *p);
const char *jpgpu_pipeline_last_error(const jpgpu_pipeline *p);
/* One synchronous call: create, run, results, destroy.  out_offsets[i] (may be NULL) receives image i's offset in host_out;
 * host_out_bytes must cover the sum of all W*H*3 rounded up to 256 per image and per chunk (n * 256 + sum is enough). */
int jpgpu_decode_batch_host(int device, const jpgpu_image_desc *descs, size_t n, const void *host_in, size_t host_in_bytes,
                            void *host_out, size_t host_out_bytes, size_t *out_offsets, int32_t *statuses, uint64_t *bytes_read);

/* ------------------------------------------------------------------------------------------------------------------
 * One process, several devices (SURVEY.md 8(b) "jpgpu_create(device_ids*, n_devices)", 8(e)).  The batch is cut into
 * contiguous image ranges of about equal scan bytes, one per device; every device has its own context, stream set,
 * batch object and host worker thread.  No collective, no peer traffic: the decode has no cross-image step. */
enum { JPGPU_MEMORY_HOST = 0, JPGPU_MEMORY_DEVICE = 1 };
typedef struct jpgpu_multi jpgpu_multi;
int jpgpu_multi_create(const int *devices, int n_devices, jpgpu_multi **out);
void jpgpu_multi_destroy(jpgpu_multi *m);
int jpgpu_multi_device_count(const jpgpu_multi *m);
/* Host only (needs no GPU): the partition jpgpu_multi_plan uses - n images into `parts` contiguous ranges of about equal
 * scan bytes; range k is first[k] .. first[k+1], `first` has parts + 1 entries. */
int jpgpu_partition(const jpgpu_image_desc *descs, size_t n, size_t parts, size_t *first);
int jpgpu_multi_plan(jpgpu_multi *m, const jpgpu_image_desc *descs, size_t n);
/* Range k: the CUDA device it runs on and its images [first, first + count). */
int jpgpu_multi_range(const jpgpu_multi *m, int k, int *device, size_t *first, size_t *count);
int jpgpu_multi_upload(jpgpu_multi *m);
int jpgpu_multi_decode(jpgpu_multi *m);       /* enqueues on every device; jpgpu_multi_sync waits */
int jpgpu_multi_sync(jpgpu_multi *m);
int jpgpu_multi_set_output_format(jpgpu_multi *m, uint32_t format);
int jpgpu_multi_download(jpgpu_multi *m, uint8_t *const *outs);
int jpgpu_multi_results(jpgpu_multi *m, int32_t *statuses, uint64_t *bytes_read);
/* Device pointer of image i's output and the CUDA device it lives on. */
void *jpgpu_multi_device_rgb(jpgpu_multi *m, size_t i, int *device, size_t *nbytes);
int jpgpu_multi_coefficients(jpgpu_multi *m, size_t i, int16_t *out, size_t cap, uint32_t nblocks[4]);
uint64_t jpgpu_multi_launch_count(const jpgpu_multi *m);
/* `steps` decodes on every device, all released together; ms[k] = device time of range k's steps (CUDA events). */
int jpgpu_multi_time_decode(jpgpu_multi *m, int steps, float *ms);
/* Plan + upload + decode (+ download to outs[i] for JPGPU_MEMORY_HOST) + results in one synchronous call. */
int jpgpu_multi_decode_batch(jpgpu_multi *m, const jpgpu_image_desc *descs, size_t n, uint8_t *const *outs,
                             int32_t *statuses, uint64_t *bytes_read, uint32_t memory_kind);

#ifdef __cplusplus
}
#endif
#endif /* JPGPU_H */
