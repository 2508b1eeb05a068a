/*
 * jpeg_oracle.h — CPU ORACLE for the jpeg-rust decode hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a statement-by-statement C restatement of the reference decoder
 * (martinhath/jpeg-rust): src/jpeg/mod.rs, src/jpeg/huffman.rs,
 * src/jpeg/decoder.rs and src/transform.rs.  It exists to CHECK the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product path (jpeg_rust_b200/, libjpgpu.so)
 * never links, imports or calls anything in this directory.
 *
 * Parity pin status: the reference ships NO golden vectors or tests (SURVEY.md
 * §4) and cannot be compiled here (no rustc/cargo).  The oracle is pinned
 * against the known-answer table of SURVEY.md §4 (coefficient-stream SHA-256,
 * bytes_read, block counts, spot pixels — tests/test_oracle_golden.py), against
 * exact scan consumption (decode ends 2 bytes before end of data on the
 * fixtures), against the in-repo encoder's ground-truth coefficients and
 * against libjpeg (PIL) within the expected gaps.  Against the Rust binary
 * itself parity is UNPINNED (stated in DESIGN.md).
 */
#ifndef JPEG_ORACLE_H
#define JPEG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Output geometry. REF = bug-compatible with decoder.rs:259-312,347-379.
 * SPEC = T.81 A.2.3 MCU order, true MCU count, box replication, crop. */
enum { ORACLE_LAYOUT_REF = 0, ORACLE_LAYOUT_SPEC = 1 };

/* Extensions beyond the reference's accepted subset (bit flags). */
enum {
    ORACLE_EXT_NONE = 0,
    ORACLE_EXT_SKIP_APPN = 1, /* skip APPn/COM/unknown length-carrying segments instead of panicking */
    ORACLE_EXT_DRI = 2        /* accept DRI and decode RSTn-separated intervals (SPEC layout only) */
};

/* How cos() is evaluated in the IDCT. CALL = cosf() per term exactly like
 * transform.rs:79-81 (used for CPU-baseline timing).  TABLE = the same 64
 * cosf() values looked up from a table (bit-identical results, ~10x faster;
 * used by the parity tests). */
enum { ORACLE_COS_CALL = 0, ORACLE_COS_TABLE = 1 };

/* Status = which reference panic (or none) the input would trigger. */
enum {
    ORACLE_OK = 0,
    ORACLE_PANIC_UNHANDLED_MARKER = 1, /* mod.rs:457 */
    ORACLE_PANIC_DRI = 2,              /* mod.rs:427 */
    ORACLE_PANIC_APP12_14 = 3,         /* mod.rs:446,449 */
    ORACLE_PANIC_DQT_PRECISION = 4,    /* mod.rs:258 */
    ORACLE_PANIC_SAMPLING_ASSERT = 5,  /* mod.rs:275-277 */
    ORACLE_PANIC_INDEX_OOB = 6,        /* any slice index out of bounds */
    ORACLE_PANIC_NO_FRAME_HEADER = 7,  /* mod.rs:388 unwrap */
    ORACLE_PANIC_MISSING_TABLE = 8,    /* decoder.rs:155,159,224 */
    ORACLE_PANIC_DC_LOOKUP = 9,        /* huffman.rs:156 */
    ORACLE_PANIC_AC_LOOKUP = 10,       /* huffman.rs:162 "ILLEGAL STATE!" */
    ORACLE_PANIC_COMPONENT_COUNT = 11, /* decoder.rs:330 */
    ORACLE_PANIC_READ_BITS_ASSERT = 12,/* huffman.rs:202 */
    ORACLE_PANIC_SCAN_COMPONENT = 13,  /* decoder.rs:148 unwrap */
    ORACLE_NO_SCAN = 14,               /* parse() returned Ok with image_data None (main.rs:38 unwrap) */
    ORACLE_PANIC_ARITH = 15,           /* debug-build arithmetic overflow (e.g. mod.rs:218 length-2) */
    ORACLE_ERR_UNSUPPORTED = 16        /* option combination the oracle does not implement */
};

typedef struct oracle_result {
    int status;          /* ORACLE_* */
    char msg[160];       /* panic message in the reference's words where it has one */
    int width, height;   /* SOF0 X, Y (mod.rs:295) */
    int ncomp;           /* components in the scan (decoder.rs:167) */
    int hs[4], vs[4];    /* sampling factors per component, scan order */
    int mcus_read;       /* MCUs decoded in phase 1 (decoder.rs:192 for REF; true count for SPEC) */
    size_t bytes_read;   /* decoder.rs:336-340 */
    size_t scan_len;     /* length of the unstuffed data vector (mod.rs:373-385) */
    uint8_t *rgb;        /* W*H*3, row-major, interleaved RGB (decoder.rs:317-331) */
    size_t rgb_len;
    /* blocks[component] after decoder.rs:208-212: 64 x i16 per block, ZIGZAG
     * order, absolute (predicted) DC, in decode order within the component. */
    int16_t *coefs[4];
    size_t nblocks[4];
    /* per-component W*H f32 planes after placement (decoder.rs:314), before colour */
    float *planes[4];
} oracle_result;

/* Whole-path entry: JPEGImage::parse(bytes) + JPEGDecoder::decode()
 * (mod.rs:202 → mod.rs:415 → decoder.rs:162). Never returns NULL. */
oracle_result *oracle_decode_file(const uint8_t *file, size_t len, int layout, int ext, int cos_mode);
void oracle_free(oracle_result *r);

/* Stage-level entry points for unit tests. */

/* transform.rs:55-87 */
void oracle_idct_8x8(const float in[64], float out[64], int cos_mode);
/* huffman.rs:37-58,80-98 : returns number of codes written (<=256) */
int oracle_build_codes(const uint8_t bits[16], const uint8_t *vals, int nvals,
                       uint8_t *out_len, uint16_t *out_code, uint8_t *out_val);
/* decoder.rs:382-402 */
void oracle_ycbcr_to_rgb(float y, float cb, float cr, uint8_t rgb[3]);
uint8_t oracle_f32_to_u8(float n);
/* huffman.rs:256-268 */
int16_t oracle_value_correction(uint16_t val, int len);
/* decoder.rs:404-407 */
const int *oracle_zigzag_indices(void);
/* mod.rs:371-385 (to end of buffer). out must hold len bytes. Returns unstuffed length,
 * or (size_t)-1 if the reference would index out of bounds (buffer ends in 0xFF). */
size_t oracle_unstuff(const uint8_t *in, size_t len, uint8_t *out);

/* Timing helper for the CPU baseline: decodes the same file `reps` times and
 * returns the best wall-clock seconds of one decode (parse + decode, like
 * main.rs:30-31 without file I/O and PPM output). */
double oracle_time_decode(const uint8_t *file, size_t len, int layout, int ext, int cos_mode, int reps);

#ifdef __cplusplus
}
#endif
#endif
