/*
 * jpeg_oracle.c — CPU ORACLE (test infrastructure, see jpeg_oracle.h).
 *
 * Restates, in the reference's own order of operations and in strict f32:
 *   src/jpeg/mod.rs      157-181 (bytes_to_marker), 202-465 (JPEGImage::parse)
 *   src/jpeg/huffman.rs  37-98 (table construction), 124-268 (HuffmanDecoder)
 *   src/jpeg/decoder.rs  83-152 (builders), 162-343 (decode), 347-402, 404-437
 *   src/transform.rs     55-87 (discrete_cosine_transform_inverse)
 * Rust panics are turned into status codes (longjmp to the entry point).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile). No
 * value-changing optimisation is allowed: every float expression below is
 * written with the association the Rust source has.
 */
#include "jpeg_oracle.h"

#include <math.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ context */

typedef struct ctx {
    jmp_buf jb;
    oracle_result *res;
    void **allocs;
    size_t nallocs, cap;
    int layout, ext, cos_mode;
} ctx;

static void panic_(ctx *c, int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c->res->msg, sizeof c->res->msg, fmt, ap);
    va_end(ap);
    c->res->status = code;
    longjmp(c->jb, 1);
}

/* scratch allocation, released when the entry point returns */
static void *xalloc(ctx *c, size_t n) {
    void *p = calloc(n ? n : 1, 1);
    if (!p) panic_(c, ORACLE_ERR_UNSUPPORTED, "out of memory");
    if (c->nallocs == c->cap) {
        c->cap = c->cap ? c->cap * 2 : 64;
        c->allocs = (void **)realloc(c->allocs, c->cap * sizeof(void *));
    }
    c->allocs[c->nallocs++] = p;
    return p;
}

static void release_scratch(ctx *c) {
    for (size_t k = 0; k < c->nallocs; k++) free(c->allocs[k]);
    free(c->allocs);
    c->allocs = NULL;
    c->nallocs = c->cap = 0;
}

/* Rust slice indexing: vec[i] panics when i >= len */
#define AT(vec, len, i) (((size_t)(i) < (size_t)(len)) ? (vec)[(i)] : (panic_(c, ORACLE_PANIC_INDEX_OOB, "index out of bounds: the len is %zu but the index is %zu", (size_t)(len), (size_t)(i)), (vec)[0]))
/* Rust range slicing vec[a..b] panics when b > len or a > b */
#define RANGE(len, a, b) do { if ((size_t)(b) > (size_t)(len) || (size_t)(a) > (size_t)(b)) panic_(c, ORACLE_PANIC_INDEX_OOB, "range end index %zu out of range for slice of length %zu", (size_t)(b), (size_t)(len)); } while (0)

/* ------------------------------------------------------- huffman.rs: tables */

typedef struct { uint8_t length; uint16_t code; uint8_t value; } hcode;      /* huffman.rs:13-21 */
typedef struct { hcode codes[256]; int n; int present; } htable;             /* huffman.rs:24-27 */

/* huffman.rs:80-98 make_code_table (T.81 Figure C.2). Returns count. */
static int make_code_table(const uint8_t *sizes, int nsizes, uint16_t *out) {
    int n = 0;
    uint16_t code = 0;
    if (nsizes == 0) return -1; /* sizes[0] would panic (index OOB) */
    unsigned current_size = sizes[0];
    for (int k = 0; k < nsizes; k++) {
        unsigned size = sizes[k];
        while (size > current_size) {
            code = (uint16_t)(code << 1);
            current_size += 1;
        }
        out[n++] = code;
        if (current_size > 16 || code == 0xffff) break;
        code = (uint16_t)(code + 1);
    }
    return n;
}

/* huffman.rs:37-58 from_size_data_tables */
static int build_table(const uint8_t size_data[16], const uint8_t *data_table, int ndata, htable *t) {
    uint8_t code_lengths[16 * 255 + 1];
    uint16_t code_table[16 * 255 + 1];
    int nl = 0;
    for (int i = 0; i < 16; i++)
        for (int k = 0; k < size_data[i]; k++) code_lengths[nl++] = (uint8_t)(i + 1);
    int nc = make_code_table(code_lengths, nl, code_table);
    if (nc < 0) return -1;
    /* zip of three iterators stops at the shortest (huffman.rs:45-55) */
    int n = ndata;
    if (nl < n) n = nl;
    if (nc < n) n = nc;
    if (n > 256) n = 256;
    for (int k = 0; k < n; k++) {
        t->codes[k].value = data_table[k];
        t->codes[k].length = code_lengths[k];
        t->codes[k].code = code_table[k];
    }
    t->n = n;
    t->present = 1;
    return n;
}

/* huffman.rs:60-76 codes_of_length: [a,b) of the contiguous run with that length */
static void codes_of_length(const htable *t, int len, int *a, int *b) {
    int i = 0;
    while (i < t->n && t->codes[i].length != len) i++;   /* skip_while */
    if (i == t->n) { *a = 0; *b = 0; return; }
    int j = i;
    while (j < t->n && t->codes[j].length == len) j++;   /* take_while */
    *a = i; *b = j;
}

int oracle_build_codes(const uint8_t bits[16], const uint8_t *vals, int nvals,
                       uint8_t *out_len, uint16_t *out_code, uint8_t *out_val) {
    htable t;
    memset(&t, 0, sizeof t);
    int n = build_table(bits, vals, nvals, &t);
    if (n < 0) return -1;
    for (int k = 0; k < n; k++) {
        out_len[k] = t.codes[k].length;
        out_code[k] = t.codes[k].code;
        out_val[k] = t.codes[k].value;
    }
    return n;
}

/* ------------------------------------------------ huffman.rs: HuffmanDecoder */

static const uint16_t BIT_MASKS[17] = {0x0, 0x8000, 0xC000, 0xE000, 0xF000, 0xF800, 0xFC00, 0xFE00, 0xFF00,
                                       0xFF80, 0xFFC0, 0xFFE0, 0xFFF0, 0xFFF8, 0xFFFC, 0xFFFE, 0xFFFF}; /* huffman.rs:5-6 */

typedef struct {
    const uint8_t *data; size_t len;
    size_t next_index;   /* huffman.rs:113 */
    size_t bits_read;    /* huffman.rs:115 */
    uint32_t current;    /* huffman.rs:120 */
} hdecoder;

/* huffman.rs:124-135 */
static void hd_new(ctx *c, hdecoder *d, const uint8_t *data, size_t len) {
    d->data = data; d->len = len;
    uint32_t b0 = AT(data, len, 0), b1 = AT(data, len, 1), b2 = AT(data, len, 2), b3 = AT(data, len, 3);
    d->current = (b0 << 24) | (b1 << 16) | (b2 << 8) | b3;
    d->next_index = 4;
    d->bits_read = 0;
}

/* huffman.rs:231-254 */
static void hd_shift_and_fix_current(hdecoder *d, size_t len) {
    if (len == 0) return;
    d->current <<= len;
    d->bits_read += len;
    while (d->bits_read >= 8) {
        d->bits_read -= 8;
        uint32_t next_num = (d->next_index >= d->len) ? 0xaa : d->data[d->next_index];
        d->current |= next_num << d->bits_read;
        d->next_index += 1;
    }
}

/* huffman.rs:198-208 */
static uint16_t hd_read_n_bits(ctx *c, hdecoder *d, size_t n) {
    if (n == 0) return 0;
    if (n > 16) panic_(c, ORACLE_PANIC_READ_BITS_ASSERT, "Should not read more than 16 bits at a time!");
    uint16_t mask = BIT_MASKS[n];
    uint16_t current_16 = (uint16_t)(d->current >> 16);
    uint16_t number = (uint16_t)((current_16 & mask) >> (16 - n));
    hd_shift_and_fix_current(d, n);
    return number;
}

/* huffman.rs:211-227: lengths 2..16, linear search of the codes of that length */
static int hd_next_code(hdecoder *d, const htable *t) {
    for (int len = 2; len < 17; len++) {
        uint16_t mask = BIT_MASKS[len];
        uint16_t current_16 = (uint16_t)(d->current >> 16);
        uint16_t bits = (uint16_t)((current_16 & mask) >> (16 - len));
        int a, b;
        codes_of_length(t, len, &a, &b);
        for (int k = a; k < b; k++) {
            if (t->codes[k].code == bits) {
                hd_shift_and_fix_current(d, (size_t)len);
                return t->codes[k].value;
            }
        }
    }
    return -1; /* None */
}

/* huffman.rs:256-268 (Table F.2 EXTEND); i16 arithmetic wraps instead of panicking */
int16_t oracle_value_correction(uint16_t val, int len) {
    if (len == 0) return 0;
    int16_t v = (int16_t)val;
    int16_t base = (int16_t)(1u << (len - 1));
    if (v < base) return (int16_t)(-2 * (int)base + 1 + (int)v);
    return v;
}

/* huffman.rs:146-195 next_block: DC difference + 63 AC, zigzag order */
static void hd_next_block(ctx *c, hdecoder *d, const htable *ac, const htable *dc, int16_t block[64]) {
    int num_bits = hd_next_code(d, dc);
    if (num_bits < 0) panic_(c, ORACLE_PANIC_DC_LOOKUP, "DC lookup fail");
    int16_t dc_coef = oracle_value_correction(hd_read_n_bits(c, d, (size_t)num_bits), num_bits);
    int blen = 0;
    block[blen++] = dc_coef;
    while (blen < 64) {
        int next_code = hd_next_code(d, ac);
        if (next_code < 0) panic_(c, ORACLE_PANIC_AC_LOOKUP, "ILLEGAL STATE!");
        if (next_code == 0x00) {            /* EOB: huffman.rs:164-169 */
            while (blen < 64) block[blen++] = 0;
            break;
        }
        if (next_code == 0xf0) {            /* ZRL: huffman.rs:170-175 */
            int to_push = 64 - blen < 16 ? 64 - blen : 16;
            for (int k = 0; k < to_push; k++) block[blen++] = 0;
            continue;
        }
        int prepending_zeroes = (next_code & 0xf0) >> 4;
        int nb = next_code & 0xf;
        uint16_t num = hd_read_n_bits(c, d, (size_t)nb);
        int16_t number = oracle_value_correction(num, nb);
        int zeroes_to_push = prepending_zeroes < 64 - blen - 1 ? prepending_zeroes : 64 - blen - 1;
        for (int k = 0; k < zeroes_to_push; k++) block[blen++] = 0;
        block[blen++] = number;
    }
}

/* ------------------------------------------------------------- transform.rs */

static float g_cos_table[8][8];
static int g_cos_table_ready = 0;

/* the exact expression of transform.rs:79-81: ((2f*xf + 1f) * uf * Pi / 16f).cos() */
static inline float ref_cos_term(int x, int u) {
    const float Pi = (float)3.14159265358979323846; /* transform.rs:16 */
    float xf = (float)x, uf = (float)u;
    return cosf((2.0f * xf + 1.0f) * uf * Pi / 16.0f);
}

static void ensure_cos_table(void) {
    if (g_cos_table_ready) return;
    for (int x = 0; x < 8; x++)
        for (int u = 0; u < 8; u++) g_cos_table[x][u] = ref_cos_term(x, u);
    g_cos_table_ready = 1;
}

/* transform.rs:55-87 discrete_cosine_transform_inverse for d = 8 */
void oracle_idct_8x8(const float in[64], float out[64], int cos_mode) {
    const float a0 = 1.0f / sqrtf(2.0f); /* transform.rs:56-62 alpha(0) */
    const int d = 8;
    if (cos_mode == ORACLE_COS_TABLE) ensure_cos_table();
    for (int y = 0; y < d; y++) {
        for (int x = 0; x < d; x++) {
            float sum = 0.0f;
            for (int v = 0; v < d; v++) {
                for (int u = 0; u < d; u++) {
                    float au = (u == 0) ? a0 : 1.0f;
                    float av = (v == 0) ? a0 : 1.0f;
                    float f_uv = in[v * d + u];
                    float cx, cy;
                    if (cos_mode == ORACLE_COS_TABLE) { cx = g_cos_table[x][u]; cy = g_cos_table[y][v]; }
                    else { cx = ref_cos_term(x, u); cy = ref_cos_term(y, v); }
                    /* transform.rs:78-81: alpha(u) * alpha(v) * f_uv * cos * cos, left-assoc */
                    float term = au * av;
                    term = term * f_uv;
                    term = term * cx;
                    term = term * cy;
                    sum = sum + term;
                }
            }
            out[y * d + x] = sum / 4.0f;
        }
    }
}

/* --------------------------------------------------------------- decoder.rs */

static const int ZIGZAG_INDICES[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27,
     20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58,
     59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63}; /* decoder.rs:404-407 */

const int *oracle_zigzag_indices(void) { return ZIGZAG_INDICES; }

/* decoder.rs:382-390 */
uint8_t oracle_f32_to_u8(float n) {
    if (n < 0.0f) return 0;
    if (n > 255.0f) return 255;
    return (uint8_t)n; /* `as u8`: truncation toward zero; NaN -> 0 in Rust */
}

/* decoder.rs:392-402 */
void oracle_ycbcr_to_rgb(float y, float cb, float cr, uint8_t rgb[3]) {
    const float c_red = 0.299f, c_green = 0.587f, c_blue = 0.114f;
    float kr = 2.0f - 2.0f * c_red;
    float kb = 2.0f - 2.0f * c_blue;
    float r = cr * kr;
    r = r + y;
    float b = cb * kb;
    b = b + y;
    float t1 = c_blue * b;
    float t2 = c_red * r;
    float g = y - t1;
    g = g - t2;
    g = g / c_green;
    rgb[0] = oracle_f32_to_u8(r + 128.0f);
    rgb[1] = oracle_f32_to_u8(g + 128.0f);
    rgb[2] = oracle_f32_to_u8(b + 128.0f);
}

/* decoder.rs:39-52 */
typedef struct {
    uint8_t component, dc_table_id, ac_table_id, quantization_id, h, v;
} compfields;

typedef struct {
    /* mod.rs:59-87 subset that reaches the decoder */
    htable ac[4], dc[4];
    uint16_t qt[4][64]; int qt_present[4];
    int have_frame;
    int frame_ncomp; uint8_t f_id[255], f_h[255], f_v[255], f_tq[255];
    int width, height;
    int restart_interval; /* EXT_DRI only */
} imgstate;

/* decoder.rs:259-288 get_indices closure */
static void get_indices(size_t x, size_t y, size_t max_x, size_t max_y, size_t x_factor, size_t y_factor,
                        size_t max_x_factor, size_t max_y_factor, size_t *ox, size_t *oy) {
    (void)max_y;
    if (max_y_factor > 1 && y_factor == 1) {
        if (max_x_factor > 1 && x_factor == 1) {
            int is_upper = (y & 1) == 0;
            if (is_upper) {
                int move_down = ((x / 2) & 1) == 1;
                if (move_down) { *ox = x / 2 - 1 + (x & 1); *oy = y + 1; return; }
                else { *ox = x / 2 + (x & 1); *oy = y; return; }
            } else {
                int move_up = y > 0 && ((x / 2) & 1) == 0;
                if (move_up) { *ox = max_x / 2 + x / 2 - 1 + (x & 1); *oy = y; return; }
                else { *ox = max_x / 2 + x / 2 + (x & 1); *oy = y - 1; return; }
            }
        } else {
            if ((y & 1) == 0) { *ox = x / 2; *oy = y + (x & 1); return; }
            else { *ox = x / 2 + max_x / 2; *oy = y - (x & 1); return; }
        }
    }
    *ox = x; *oy = y;
}

/* decoder.rs:347-379 fill_block_in_array */
static void fill_block_in_array(ctx *c, const float block[64], float *target, size_t target_len,
                                size_t x_scale, size_t y_scale, size_t x, size_t y, size_t stride) {
    for (size_t line_number = 0; line_number < 8; line_number++) {
        size_t start_x = x * 8 * x_scale;
        if (stride < start_x) continue;                 /* decoder.rs:363-365 (per line) */
        size_t start_i = y * 8 * y_scale * stride + line_number * stride + start_x;
        for (size_t ind = 0; ind < 8 * x_scale; ind++) { /* line = 8 samples, each repeated x_scale times */
            float n = block[line_number * 8 + ind / x_scale];
            size_t i = ind + start_i;
            for (size_t j = 0; j < y_scale; j++) {
                if (i + j * stride < target_len) {       /* decoder.rs:371 */
                    size_t idx = i + j * stride * 8;     /* decoder.rs:372 */
                    if (idx >= target_len)
                        panic_(c, ORACLE_PANIC_INDEX_OOB, "index out of bounds: the len is %zu but the index is %zu", target_len, idx);
                    target[idx] = n;
                }
            }
        }
    }
}

/* builder: decoder.rs:83-152 frame_header() then scan_header() */
static int build_component_fields(ctx *c, const imgstate *im, int ns, const uint8_t *s_id, const uint8_t *s_td,
                                  const uint8_t *s_ta, compfields *out) {
    compfields cf[512];
    int n = 0;
    for (int k = 0; k < im->frame_ncomp; k++) {          /* decoder.rs:84-109 */
        int found = -1;
        for (int m = 0; m < n; m++) if (cf[m].component == im->f_id[k]) { found = m; break; }
        if (found >= 0) {
            cf[found].h = im->f_h[k]; cf[found].v = im->f_v[k]; cf[found].quantization_id = im->f_tq[k];
        } else {
            cf[n].component = im->f_id[k]; cf[n].h = im->f_h[k]; cf[n].v = im->f_v[k];
            cf[n].quantization_id = im->f_tq[k]; cf[n].dc_table_id = 0xff; cf[n].ac_table_id = 0xff;
            n++;
        }
    }
    for (int k = 0; k < ns; k++) {                       /* decoder.rs:114-138 */
        int found = -1;
        for (int m = 0; m < n; m++) if (cf[m].component == s_id[k]) { found = m; break; }
        if (found >= 0) {
            cf[found].ac_table_id = s_ta[k]; cf[found].dc_table_id = s_td[k];
        } else {
            cf[n].component = s_id[k]; cf[n].h = 0xff; cf[n].v = 0xff; cf[n].quantization_id = 0xff;
            cf[n].dc_table_id = s_ta[k]; cf[n].ac_table_id = s_td[k]; /* decoder.rs:133-134 (swapped in the reference) */
            n++;
        }
    }
    for (int k = 0; k < ns; k++) {                       /* decoder.rs:141-150 */
        int found = -1;
        for (int m = 0; m < n; m++) if (cf[m].component == s_id[k]) { found = m; break; }
        if (found < 0) panic_(c, ORACLE_PANIC_SCAN_COMPONENT, "called `Option::unwrap()` on a `None` value");
        out[k] = cf[found];
    }
    return ns;
}

static const htable *get_ac(ctx *c, const imgstate *im, unsigned id) {   /* decoder.rs:154-156 */
    if (id >= 4) panic_(c, ORACLE_PANIC_INDEX_OOB, "index out of bounds: the len is 4 but the index is %u", id);
    if (!im->ac[id].present) panic_(c, ORACLE_PANIC_MISSING_TABLE, "called `Option::unwrap()` on a `None` value");
    return &im->ac[id];
}
static const htable *get_dc(ctx *c, const imgstate *im, unsigned id) {   /* decoder.rs:158-160 */
    if (id >= 4) panic_(c, ORACLE_PANIC_INDEX_OOB, "index out of bounds: the len is 4 but the index is %u", id);
    if (!im->dc[id].present) panic_(c, ORACLE_PANIC_MISSING_TABLE, "called `Option::unwrap()` on a `None` value");
    return &im->dc[id];
}

typedef struct { int16_t *v; size_t n, cap; } blockvec; /* Vec<Block>, 64 i16 each */

static void bv_push(ctx *c, blockvec *b, const int16_t blk[64]) {
    if (b->n == b->cap) {
        size_t ncap = b->cap ? b->cap * 2 : 1024;
        int16_t *nv = (int16_t *)realloc(b->v, ncap * 64 * sizeof(int16_t));
        if (!nv) panic_(c, ORACLE_ERR_UNSUPPORTED, "out of memory");
        b->v = nv; b->cap = ncap;
    }
    memcpy(b->v + b->n * 64, blk, 64 * sizeof(int16_t));
    b->n++;
}

/* dequantise (zigzag order) + zigzag_inverse + IDCT: decoder.rs:227-235, 425-437 */
static void dequant_dezigzag_idct(const int16_t zz[64], const uint16_t q[64], float out[64], int cos_mode) {
    float nat[64];
    for (int k = 0; k < 64; k++) {
        float n = (float)zz[k];            /* decoder.rs:204 `i as f32` (DC already summed in f32) */
        nat[ZIGZAG_INDICES[k]] = n * (float)q[k];
    }
    oracle_idct_8x8(nat, out, cos_mode);
}

static void colour_convert(ctx *c, oracle_result *r, int ncomp, float **planes, size_t npix) {
    r->rgb_len = npix * 3;
    r->rgb = (uint8_t *)malloc(r->rgb_len ? r->rgb_len : 1);
    if (ncomp == 1) {                                    /* decoder.rs:317-324 */
        for (size_t p = 0; p < npix; p++) {
            uint8_t u = oracle_f32_to_u8(planes[0][p] + 128.0f);
            r->rgb[3 * p] = u; r->rgb[3 * p + 1] = u; r->rgb[3 * p + 2] = u;
        }
    } else if (ncomp == 3) {                             /* decoder.rs:325-328 */
        for (size_t p = 0; p < npix; p++)
            oracle_ycbcr_to_rgb(planes[0][p], planes[1][p], planes[2][p], r->rgb + 3 * p);
    } else {
        panic_(c, ORACLE_PANIC_COMPONENT_COUNT, "asd"); /* decoder.rs:330 */
    }
}

/* JPEGDecoder::decode(), decoder.rs:162-343, REF layout (literal) */
static void decode_ref(ctx *c, const imgstate *im, const compfields *cf, int ncomp,
                       const uint8_t *data, size_t data_len) {
    oracle_result *r = c->res;
    size_t W = (size_t)im->width, H = (size_t)im->height;
    size_t num_blocks_x = (W + 7) / 8, num_blocks_y = (H + 7) / 8;
    size_t num_blocks = num_blocks_x * num_blocks_y;

    size_t maxh = 1, maxv = 1;                           /* decoder.rs:175-185 (unwrap_or(1)) */
    if (ncomp > 0) { maxh = 0; maxv = 0; }
    for (int k = 0; k < ncomp; k++) { if (cf[k].h > maxh) maxh = cf[k].h; if (cf[k].v > maxv) maxv = cf[k].v; }

    hdecoder hd;
    hd_new(c, &hd, data, data_len);                      /* decoder.rs:189 */

    size_t skip_factor = maxv * maxh;
    if (skip_factor == 0) panic_(c, ORACLE_PANIC_ARITH, "attempt to divide by zero");
    size_t num_read_blocks = (num_blocks + skip_factor - 1) / skip_factor; /* decoder.rs:191-192 */
    r->mcus_read = (int)num_read_blocks;

    blockvec blocks[4];
    memset(blocks, 0, sizeof blocks);
    float previous_dc[4] = {0, 0, 0, 0};
    if (ncomp > 4) panic_(c, ORACLE_ERR_UNSUPPORTED, "more than 4 scan components");

    /* Step 1: decoder.rs:195-215 */
    for (size_t m = 0; m < num_read_blocks; m++) {
        for (int ci = 0; ci < ncomp; ci++) {
            const htable *ac = get_ac(c, im, cf[ci].ac_table_id);
            const htable *dc = get_dc(c, im, cf[ci].dc_table_id);
            unsigned hv = (unsigned)cf[ci].h * (unsigned)cf[ci].v;
            if (hv > 255) panic_(c, ORACLE_PANIC_ARITH, "attempt to multiply with overflow");
            for (unsigned k = 0; k < hv; k++) {
                int16_t blk[64];
                hd_next_block(c, &hd, ac, dc, blk);
                float encoded = (float)blk[0];           /* decoder.rs:208-210, f32 */
                float dcv = encoded + previous_dc[ci];
                previous_dc[ci] = dcv;
                blk[0] = (int16_t)(dcv < -32768.0f ? -32768 : (dcv > 32767.0f ? 32767 : (int)dcv));
                bv_push(c, &blocks[ci], blk);
                r->coefs[ci] = blocks[ci].v; r->nblocks[ci] = blocks[ci].n; /* owned by the result */
            }
        }
    }

    /* Step 2: decoder.rs:220-315 */
    size_t num_pixels = W * H;
    for (int ci = 0; ci < ncomp; ci++) {
        unsigned qid = cf[ci].quantization_id;
        if (qid >= 4) panic_(c, ORACLE_PANIC_INDEX_OOB, "index out of bounds: the len is 4 but the index is %u", qid);
        if (!im->qt_present[qid]) panic_(c, ORACLE_PANIC_MISSING_TABLE, "Did not find quantization table for %u", qid);
        const uint16_t *q = im->qt[qid];

        float *component_blocks = (float *)xalloc(c, blocks[ci].n * 64 * sizeof(float));
        for (size_t b = 0; b < blocks[ci].n; b++)
            dequant_dezigzag_idct(blocks[ci].v + b * 64, q, component_blocks + b * 64, c->cos_mode);

        /* decoder.rs:239-251, f32 arithmetic */
        float x_i = ceilf((float)W * ((float)cf[ci].h / (float)maxh));
        float y_i = ceilf((float)H * ((float)cf[ci].v / (float)maxv));
        size_t x_factor = (size_t)ceilf((float)W / x_i);
        size_t y_factor = (size_t)ceilf((float)H / y_i);
        size_t stride = W;
        if (x_factor == 0 || y_factor == 0) panic_(c, ORACLE_PANIC_ARITH, "attempt to divide by zero");

        float *plane = (float *)malloc((num_pixels ? num_pixels : 1) * sizeof(float));
        for (size_t p = 0; p < num_pixels; p++) plane[p] = 0.0f;
        r->planes[ci] = plane;
        size_t block_i = 0;
        for (size_t y = 0; y < num_blocks_y / y_factor; y++) {
            for (size_t x = 0; x < num_blocks_x / x_factor; x++) {
                size_t xi, yi;
                get_indices(x, y, num_blocks_x, num_blocks_y, x_factor, y_factor, maxh, maxv, &xi, &yi);
                if (block_i >= blocks[ci].n)
                    panic_(c, ORACLE_PANIC_INDEX_OOB, "index out of bounds: the len is %zu but the index is %zu", blocks[ci].n, block_i);
                fill_block_in_array(c, component_blocks + block_i * 64, plane, num_pixels, x_factor, y_factor, xi, yi, stride);
                block_i += 1;
            }
        }
    }

    colour_convert(c, r, ncomp, r->planes, num_pixels);

    /* decoder.rs:336-340 */
    r->bytes_read = (hd.bits_read > 0 ? hd.next_index + 1 : hd.next_index) - 4;
}

/* SPEC layout: same entropy decoder, dequant, IDCT and colour as the reference;
 * geometry per T.81 A.1.1/A.2.3 (what decoder.rs:238-312 intends): true MCU
 * count, MCU-interleaved block order, box replication on both axes, crop at the
 * right/bottom edge.  With EXT_DRI the scan is split at RSTn markers and every
 * interval is unstuffed (mod.rs:371-385 logic) and decoded with fresh DC
 * predictors — the reference has no such mode (mod.rs:424-428 panics). */
static void decode_spec(ctx *c, const imgstate *im, const compfields *cf_in, int ncomp,
                        const uint8_t *raw, size_t raw_len) {
    oracle_result *r = c->res;
    size_t W = (size_t)im->width, H = (size_t)im->height;
    if (ncomp > 4 || ncomp < 1) panic_(c, ORACLE_PANIC_COMPONENT_COUNT, "asd");
    compfields cf[4];
    for (int k = 0; k < ncomp; k++) cf[k] = cf_in[k];
    if (ncomp == 1) { cf[0].h = 1; cf[0].v = 1; }        /* T.81 A.2.2: a one-component scan is non-interleaved */
    size_t maxh = 0, maxv = 0;
    for (int k = 0; k < ncomp; k++) { if (cf[k].h > maxh) maxh = cf[k].h; if (cf[k].v > maxv) maxv = cf[k].v; }
    for (int k = 0; k < ncomp; k++) {
        if (cf[k].h == 0 || cf[k].v == 0 || maxh % cf[k].h || maxv % cf[k].v)
            panic_(c, ORACLE_ERR_UNSUPPORTED, "SPEC layout needs sampling factors dividing the maximum");
        r->hs[k] = cf[k].h; r->vs[k] = cf[k].v;
    }
    size_t mcux = (W + 8 * maxh - 1) / (8 * maxh), mcuy = (H + 8 * maxv - 1) / (8 * maxv);
    size_t nmcu = mcux * mcuy;
    r->mcus_read = (int)nmcu;

    blockvec blocks[4];
    memset(blocks, 0, sizeof blocks);
    size_t ri = (c->ext & ORACLE_EXT_DRI) ? (size_t)im->restart_interval : 0;
    size_t pos = 0, mcu_done = 0, total_unstuffed = 0;
    uint8_t *un = (uint8_t *)xalloc(c, raw_len + 8);
    while (mcu_done < nmcu) {
        /* unstuff one interval: to the next RSTn (ri>0) or to the end of the buffer (mod.rs:371-385) */
        size_t n = 0, i = pos;
        int hit_rst = 0;
        while (i < raw_len) {
            if (ri && raw[i] == 0xff && i + 1 < raw_len && raw[i + 1] >= 0xd0 && raw[i + 1] <= 0xd7) { hit_rst = 1; break; }
            un[n++] = raw[i];
            if (raw[i] == 0xff && AT(raw, raw_len, i + 1) == 0x00) i += 1;
            i += 1;
        }
        size_t todo = ri ? (nmcu - mcu_done < ri ? nmcu - mcu_done : ri) : nmcu;
        hdecoder hd;
        if (ri && n < 4 && raw_len >= 4) {
            /* restart-interval extension (not in the reference): an interval may hold fewer than the four bytes
             * HuffmanDecoder::new reads (huffman.rs:127-128) - a few flat MCUs at low quality.  The window is
             * filled the way the refill fills it past the end of the data (huffman.rs:246: 0xaa). */
            hd.data = un; hd.len = n; hd.current = 0;
            for (size_t k = 0; k < 4; k++) hd.current = (hd.current << 8) | (k < n ? un[k] : 0xaau);
            hd.next_index = 4; hd.bits_read = 0;
        } else {
            hd_new(c, &hd, un, n);
        }
        float previous_dc[4] = {0, 0, 0, 0};
        for (size_t m = 0; m < todo; m++) {
            for (int ci = 0; ci < ncomp; ci++) {
                const htable *ac = get_ac(c, im, cf[ci].ac_table_id);
                const htable *dc = get_dc(c, im, cf[ci].dc_table_id);
                for (unsigned k = 0; k < (unsigned)cf[ci].h * cf[ci].v; k++) {
                    int16_t blk[64];
                    hd_next_block(c, &hd, ac, dc, blk);
                    float dcv = (float)blk[0] + previous_dc[ci];
                    previous_dc[ci] = dcv;
                    blk[0] = (int16_t)(dcv < -32768.0f ? -32768 : (dcv > 32767.0f ? 32767 : (int)dcv));
                    bv_push(c, &blocks[ci], blk);
                    r->coefs[ci] = blocks[ci].v; r->nblocks[ci] = blocks[ci].n;
                }
            }
        }
        mcu_done += todo;
        size_t used = (hd.bits_read > 0 ? hd.next_index + 1 : hd.next_index) - 4;
        r->bytes_read = total_unstuffed + used;
        total_unstuffed += n;
        if (!hit_rst) break;
        pos = i + 2; /* skip the RSTn marker */
    }
    if (mcu_done < nmcu) panic_(c, ORACLE_PANIC_INDEX_OOB, "scan ended after %zu of %zu MCUs", mcu_done, nmcu);
    size_t num_pixels = W * H;
    for (int ci = 0; ci < ncomp; ci++) {
        unsigned qid = cf[ci].quantization_id;
        if (qid >= 4) panic_(c, ORACLE_PANIC_INDEX_OOB, "index out of bounds: the len is 4 but the index is %u", qid);
        if (!im->qt_present[qid]) panic_(c, ORACLE_PANIC_MISSING_TABLE, "Did not find quantization table for %u", qid);
        const uint16_t *q = im->qt[qid];
        size_t hc = cf[ci].h, vc = cf[ci].v;
        size_t cbw = mcux * hc;                          /* component width in blocks (MCU padded) */
        size_t cw = cbw * 8, ch = mcuy * vc * 8;
        float *cplane = (float *)xalloc(c, cw * ch * sizeof(float));
        size_t b = 0;
        for (size_t my = 0; my < mcuy; my++)
            for (size_t mx = 0; mx < mcux; mx++)
                for (size_t by = 0; by < vc; by++)
                    for (size_t bx = 0; bx < hc; bx++, b++) {
                        float px[64];
                        dequant_dezigzag_idct(blocks[ci].v + b * 64, q, px, c->cos_mode);
                        size_t x0 = (mx * hc + bx) * 8, y0 = (my * vc + by) * 8;
                        for (int yy = 0; yy < 8; yy++)
                            for (int xx = 0; xx < 8; xx++) cplane[(y0 + yy) * cw + x0 + xx] = px[yy * 8 + xx];
                    }
        size_t xf = maxh / hc, yf = maxv / vc;           /* replication factors (decoder.rs:249-250 for factors in {1,2}) */
        float *plane = (float *)malloc((num_pixels ? num_pixels : 1) * sizeof(float));
        r->planes[ci] = plane;
        for (size_t y = 0; y < H; y++)
            for (size_t x = 0; x < W; x++) plane[y * W + x] = cplane[(y / yf) * cw + (x / xf)];
    }
    colour_convert(c, r, ncomp, r->planes, num_pixels);
}

/* ------------------------------------------------------------------ mod.rs */

enum { M_SOS, M_DHT, M_COM, M_DQT, M_SOF0, M_DRI, M_APP0, M_APP12, M_APP14, M_SOI, M_EOI, M_NONE, M_SKIP };

static inline uint16_t u8s_to_u16(ctx *c, const uint8_t *vec, size_t len, size_t i) { /* mod.rs:9-13 */
    uint16_t msb = AT(vec, len, i), lsb = AT(vec, len, i + 1);
    return (uint16_t)((msb << 8) + lsb);
}

size_t oracle_unstuff(const uint8_t *in, size_t len, uint8_t *out) { /* mod.rs:371-385 */
    size_t n = 0, i = 0;
    while (i < len) {
        out[n++] = in[i];
        if (in[i] == 0xff) {
            if (i + 1 >= len) return (size_t)-1; /* vec[i + 1] panics */
            if (in[i + 1] == 0x00) i += 1;
        }
        i += 1;
    }
    return n;
}

/* JPEGImage::parse, mod.rs:202-465 */
static void parse(ctx *c, const uint8_t *vec, size_t len) {
    oracle_result *r = c->res;
    imgstate *im = (imgstate *)xalloc(c, sizeof(imgstate));
    size_t i = 0;
    while (i < len) {
        /* bytes_to_marker(&vec[i..]), mod.rs:157-181 */
        int marker = M_NONE;
        uint8_t d0 = AT(vec, len, i);
        if (d0 == 0xff) {
            uint8_t n = AT(vec, len, i + 1);
            if (n == 0) n = AT(vec, len, i + 2);
            switch (n) {
                case 0xc0: marker = M_SOF0; break;
                case 0xc4: marker = M_DHT; break;
                case 0xd8: marker = M_SOI; break;
                case 0xd9: marker = M_EOI; break;
                case 0xda: marker = M_SOS; break;
                case 0xdb: marker = M_DQT; break;
                case 0xdd: marker = M_DRI; break;
                case 0xe0: marker = M_APP0; break;
                case 0xec: marker = M_APP12; break;
                case 0xee: marker = M_APP14; break;
                case 0xfe: marker = M_COM; break;
                default:
                    /* extension: any other APPn / JPGn / DNL-free segment that carries a length */
                    if ((c->ext & ORACLE_EXT_SKIP_APPN) && ((n >= 0xe1 && n <= 0xef) || (n >= 0xf0 && n <= 0xfd)))
                        marker = M_SKIP;
                    break;
            }
        }
        if (marker == M_NONE)                            /* mod.rs:456-462 */
            panic_(c, ORACLE_PANIC_UNHANDLED_MARKER, "Unhandled byte marker: %02x %02x (i=%zu/%zu)",
                   d0, AT(vec, len, i + 1), i, len);
        if (marker == M_EOI || marker == M_SOI) { i += 2; continue; } /* mod.rs:208-214 */

        uint16_t seglen = u8s_to_u16(c, vec, len, i + 2);
        if (seglen < 2) panic_(c, ORACLE_PANIC_ARITH, "attempt to subtract with overflow"); /* mod.rs:218 */
        size_t data_length = (size_t)(seglen - 2);
        i += 4;

        switch (marker) {
        case M_COM:                                      /* mod.rs:222-227 */
            RANGE(len, i, i + data_length);
            break;
        case M_DQT: {                                    /* mod.rs:228-261 */
            size_t index = i;
            while (index < i + data_length) {
                uint8_t pq_tq = AT(vec, len, index);
                unsigned precision = (pq_tq & 0xf0) >> 4, identifier = pq_tq & 0x0f;
                if (precision == 0) {
                    RANGE(len, index + 1, index + 65);
                    if (identifier >= 4) panic_(c, ORACLE_PANIC_INDEX_OOB, "index out of bounds: the len is 4 but the index is %u", identifier);
                    for (int k = 0; k < 64; k++) im->qt[identifier][k] = vec[index + 1 + k];
                    im->qt_present[identifier] = 1;
                    index += 65;
                } else if (precision == 1) {
                    RANGE(len, index + 1, index + 129);
                    if (identifier >= 4) panic_(c, ORACLE_PANIC_INDEX_OOB, "index out of bounds: the len is 4 but the index is %u", identifier);
                    for (int k = 0; k < 64; k++)
                        im->qt[identifier][k] = (uint16_t)((vec[index + 1 + 2 * k] << 8) | vec[index + 2 + 2 * k]);
                    im->qt_present[identifier] = 1;
                    index += 129;
                } else {
                    panic_(c, ORACLE_PANIC_DQT_PRECISION, "Unknown precision of quantization table: %u", precision);
                }
            }
            break;
        }
        case M_SOF0: {                                   /* mod.rs:262-298 */
            (void)AT(vec, len, i);                       /* sample_precision */
            uint16_t num_lines = u8s_to_u16(c, vec, len, i + 1);
            uint16_t samples_per_line = u8s_to_u16(c, vec, len, i + 3);
            uint8_t image_components = AT(vec, len, i + 5);
            size_t index = i + 6;
            im->frame_ncomp = 0;
            for (unsigned comp = 0; comp < image_components; comp++) {
                uint8_t id = AT(vec, len, index);
                uint8_t hv = AT(vec, len, index + 1);
                uint8_t h = (hv & 0xf0) >> 4, v = hv & 0x0f;
                if (!(h > 0 && h < 3)) panic_(c, ORACLE_PANIC_SAMPLING_ASSERT, "assertion failed: horizontal_sampling_factor > 0 && horizontal_sampling_factor < 3");
                if (!(v > 0 && v < 3)) panic_(c, ORACLE_PANIC_SAMPLING_ASSERT, "assertion failed: vertical_sampling_factor > 0 && vertical_sampling_factor < 3");
                uint8_t tq = AT(vec, len, index + 2);
                im->f_id[im->frame_ncomp] = id; im->f_h[im->frame_ncomp] = h;
                im->f_v[im->frame_ncomp] = v; im->f_tq[im->frame_ncomp] = tq;
                im->frame_ncomp++;
                index += 3;
            }
            im->width = samples_per_line; im->height = num_lines; /* mod.rs:295 */
            im->have_frame = 1;
            r->width = im->width; r->height = im->height;
            break;
        }
        case M_DHT: {                                    /* mod.rs:299-336 */
            size_t huffman_index = i, segment_end = i + data_length;
            while (huffman_index < segment_end) {
                uint8_t tc_th = AT(vec, len, huffman_index);
                unsigned table_class = (tc_th & 0xf0) >> 4, table_dest_id = tc_th & 0x0f;
                huffman_index += 1;
                RANGE(len, huffman_index, huffman_index + 16);
                const uint8_t *size_area = vec + huffman_index;
                huffman_index += 16;
                size_t number_of_codes = 0;
                for (int k = 0; k < 16; k++) number_of_codes += size_area[k];
                RANGE(len, huffman_index, huffman_index + number_of_codes);
                const uint8_t *data_area = vec + huffman_index;
                huffman_index += number_of_codes;
                htable t;
                memset(&t, 0, sizeof t);
                if (build_table(size_area, data_area, (int)number_of_codes, &t) < 0)
                    panic_(c, ORACLE_PANIC_INDEX_OOB, "index out of bounds: the len is 0 but the index is 0"); /* huffman.rs:85 sizes[0] */
                if (table_dest_id >= 4) panic_(c, ORACLE_PANIC_INDEX_OOB, "index out of bounds: the len is 4 but the index is %u", table_dest_id);
                if (table_class == 0) im->dc[table_dest_id] = t; else im->ac[table_dest_id] = t;
            }
            break;
        }
        case M_SOS: {                                    /* mod.rs:337-423 */
            uint8_t num_components = AT(vec, len, i);
            uint8_t s_id[255], s_td[255], s_ta[255];
            for (unsigned comp = 0; comp < num_components; comp++) {
                s_id[comp] = AT(vec, len, i + 1);
                uint8_t tdta = AT(vec, len, i + 2);
                s_td[comp] = (tdta & 0xf0) >> 4; s_ta[comp] = tdta & 0x0f;
                i += 2;
            }
            (void)AT(vec, len, i + 1); (void)AT(vec, len, i + 2); (void)AT(vec, len, i + 3); /* Ss, Se, Ah/Al */
            i += 4;

            /* mod.rs:371-385: unstuff from here to the end of the file */
            uint8_t *encoded = (uint8_t *)xalloc(c, len - (i < len ? i : len) + 8);
            size_t enc_len = 0;
            if (i < len) {
                enc_len = oracle_unstuff(vec + i, len - i, encoded);
                if (enc_len == (size_t)-1)
                    panic_(c, ORACLE_PANIC_INDEX_OOB, "index out of bounds: the len is %zu but the index is %zu", len, len);
            }
            r->scan_len = enc_len;

            if (!im->have_frame) panic_(c, ORACLE_PANIC_NO_FRAME_HEADER, "called `Option::unwrap()` on a `None` value"); /* mod.rs:388 */
            compfields cf[255];
            int ncomp = build_component_fields(c, im, num_components, s_id, s_td, s_ta, cf);
            r->ncomp = ncomp;
            for (int k = 0; k < ncomp && k < 4; k++) { r->hs[k] = cf[k].h; r->vs[k] = cf[k].v; }

            if (c->layout == ORACLE_LAYOUT_REF) {
                if (im->restart_interval)
                    panic_(c, ORACLE_ERR_UNSUPPORTED, "restart intervals need the SPEC layout");
                decode_ref(c, im, cf, ncomp, encoded, enc_len);       /* mod.rs:415 */
            } else {
                decode_spec(c, im, cf, ncomp, vec + (i < len ? i : len), len - (i < len ? i : len));
            }
            return;                                      /* mod.rs:416-417: first scan only */
        }
        case M_DRI:                                      /* mod.rs:424-428 */
            if (c->ext & ORACLE_EXT_DRI) {
                im->restart_interval = u8s_to_u16(c, vec, len, i);
            } else {
                panic_(c, ORACLE_PANIC_DRI, "got to restart interval def");
            }
            break;
        case M_APP0:                                     /* mod.rs:429-444: fixed absolute offsets */
            RANGE(len, i, i + 6);
            (void)AT(vec, len, 7); (void)AT(vec, len, 8); (void)AT(vec, len, 13);
            RANGE(len, 10, 12); RANGE(len, 12, 14);
            (void)AT(vec, len, 14); (void)AT(vec, len, 15);
            break;
        case M_APP12:                                    /* mod.rs:445-447 */
            if (!(c->ext & ORACLE_EXT_SKIP_APPN)) panic_(c, ORACLE_PANIC_APP12_14, "got ApplicationSegment12");
            break;
        case M_APP14:                                    /* mod.rs:448-450 */
            if (!(c->ext & ORACLE_EXT_SKIP_APPN)) panic_(c, ORACLE_PANIC_APP12_14, "got ApplicationSegment14");
            break;
        default: break;                                  /* M_SKIP (extension) */
        }
        i += data_length;                                /* mod.rs:455 */
    }
    r->status = ORACLE_NO_SCAN;                          /* mod.rs:464 Ok(image) with image_data None */
    snprintf(r->msg, sizeof r->msg, "no scan: image_data is None");
}

/* ------------------------------------------------------------ entry points */

oracle_result *oracle_decode_file(const uint8_t *file, size_t len, int layout, int ext, int cos_mode) {
    oracle_result *r = (oracle_result *)calloc(1, sizeof *r);
    ctx c;
    memset(&c, 0, sizeof c);
    c.res = r; c.layout = layout; c.ext = ext; c.cos_mode = cos_mode;
    if (cos_mode == ORACLE_COS_TABLE) ensure_cos_table();
    if (setjmp(c.jb) == 0) {
        if ((ext & ORACLE_EXT_DRI) && layout != ORACLE_LAYOUT_SPEC)
            panic_(&c, ORACLE_ERR_UNSUPPORTED, "EXT_DRI needs the SPEC layout");
        parse(&c, file, len);
    }
    release_scratch(&c);
    return r;
}

void oracle_free(oracle_result *r) {
    if (!r) return;
    free(r->rgb);
    for (int k = 0; k < 4; k++) { free(r->coefs[k]); free(r->planes[k]); }
    free(r);
}

double oracle_time_decode(const uint8_t *file, size_t len, int layout, int ext, int cos_mode, int reps) {
    double best = 1e300;
    for (int k = 0; k < reps; k++) {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        oracle_result *r = oracle_decode_file(file, len, layout, ext, cos_mode);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        int ok = r->status == ORACLE_OK;
        oracle_free(r);
        if (!ok) return -1.0;
        double s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        if (s < best) best = s;
    }
    return best;
}
