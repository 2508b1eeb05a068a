// jpgpu_core.h — data layout and the per-thread decode core shared by the CUDA
// kernels (jpgpu_kernels.cu) and, for host-side unit tests of the algorithm, by a
// CPU simulation harness under tests/ (never by the product path).
//
// Reference semantics restated here (file:line in martinhath/jpeg-rust):
//   huffman.rs:146-195  next_block       -> decode_span(): DC size/EXTEND, EOB, ZRL, (run,size)
//   huffman.rs:211-227  next_code        -> huff_lookup(): canonical prefix decode
//   huffman.rs:256-268  value_correction -> extend()
//   decoder.rs:195-215  MCU loop + DC prediction -> g/c bookkeeping + dc predictors
// The parallel formulation (self-synchronising subsequences) is new; see DESIGN.md.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define JPGPU_HD __host__ __device__ __forceinline__
#else
#define JPGPU_HD inline
#endif

namespace jpgpu {

// ------------------------------------------------------------------ constants
constexpr int kLutBits = 10;               // first-level Huffman LUT index width
constexpr int kLutSize = 1 << kLutBits;
constexpr int kMaxBlocksPerMcu = 12;       // 3 components x (H,V in {1,2})
constexpr int kMaxLutSlots = 6;            // distinct (DC, AC) tables one image can reference
#ifndef JPGPU_SEQ_THREADS
#define JPGPU_SEQ_THREADS 256
#endif
// Bits per subsequence (one decode thread each) are a per-batch plan parameter
// (BatchDev::sub_bits): 1024, 2048 or 4096 — larger means fewer re-decodes until
// synchronisation, smaller means more threads for small batches.
constexpr int kMinSubseqBits = 1024;
constexpr int kMaxSubseqBits = 4096;
constexpr int kSeqThreads = JPGPU_SEQ_THREADS;  // subsequences per sequence (= CTA size of the sync/write kernels)
constexpr int kStreamPadWords = 8;         // zero words readable past every image's stream

// status bits accumulated per image on the device (mapped to JPGPU_* by the host)
enum : uint32_t {
    kStBadCode = 1u,       // a bit pattern matched no code (true decode path only)
    kStDone = 2u,          // all MCUs were decoded
    kStRestart = 4u,       // RST markers out of sequence / wrong count
    kStDcSize = 8u         // DC size category > 16 (reference: huffman.rs:202 assert)
};

// One Huffman table in device format.
// Packed decode entry: bits 0-7 symbol, 8-12 code length, 13-18 total bits (code +
// value bits), 19-25 zigzag advance (1 for DC; run+1; 16 for ZRL; 64 for EOB), 26 DC
// size category > 16 (huffman.rs:202 assert).  With `advance`, next_block's three cases
// (huffman.rs:164-189) collapse into nz = min(z + advance, 64).
struct HuffLut {
    uint32_t fast[kLutSize];  // entry for codes of len <= kLutBits, 0 otherwise
    int32_t maxcode[18];      // largest code of length l (right aligned), -1 if none; l = 1..16
    int32_t valoff[18];       // index of first symbol of length l minus its smallest code
    uint8_t vals[256];        // HUFFVAL
    uint32_t is_dc;
    uint32_t pad[3];
};

JPGPU_HD uint32_t make_entry(uint32_t sym, uint32_t len, bool is_dc) {
    uint32_t size, adv, big = 0;
    if (is_dc) { size = sym; adv = 1; if (size > 16) { size = 16; big = 1; } }
    else if (sym == 0x00) { size = 0; adv = 64; }
    else if (sym == 0xf0) { size = 0; adv = 16; }
    else { size = sym & 15; adv = (sym >> 4) + 1; }
    return sym | (len << 8) | ((len + size) << 13) | (adv << 19) | (big << 26);
}
constexpr uint32_t kBadEntry = (16u << 8) | (16u << 13) | (64u << 19);  // unknown code: 16 bits, ends the block

// Per-image plan, written by the host, read by every kernel.
struct ImgDev {
    uint32_t width, height;
    uint32_t raw_len;         // stuffed scan bytes
    uint32_t stream_cap_words;
    uint64_t raw_off;         // byte offset into the raw arena (16-byte aligned)
    uint64_t stream_off;      // word offset into the compacted stream arena (32-word aligned)
    uint64_t coef_off;        // int16 offset into the coefficient arena
    uint64_t rgb_off;         // byte offset into the RGB arena (256-byte aligned)
    uint32_t seg_off;         // offset into the segment table (entries: start bit of each restart interval)
    uint32_t nseg_cap;        // expected number of intervals (1 without DRI)
    uint32_t sub_off;         // offset into the subsequence-info array
    uint32_t nsub_cap;        // capacity in subsequences (from raw_len)
    uint32_t seq_first;       // index of this image's first sequence in the global sequence list
    uint32_t nseq;
    uint32_t mcux, mcuy;      // MCU grid (SPEC geometry)
    uint32_t units;           // MCUs to decode (decoder.rs:192 for REF, mcux*mcuy for SPEC)
    uint32_t restart_interval;
    uint32_t seg_units;       // restart_interval * blocks_per_mcu * 64 (coefficient positions per interval)
    uint32_t total_coefs;     // units * blocks_per_mcu * 64
    uint8_t ncomp, blocks_per_mcu, hmax, vmax;
    uint8_t h[4], v[4];
    uint8_t blk_comp[kMaxBlocksPerMcu];   // component of block c inside an MCU
    uint8_t blk_dc_slot[kMaxBlocksPerMcu];
    uint8_t blk_ac_slot[kMaxBlocksPerMcu];
    uint8_t nslots, kind, layout, pad0;   // kind: colour kernel variant (see ImgKind)
    uint32_t slot_lut[kMaxLutSlots];      // index into the global HuffLut array
    uint32_t qt_off[4];                   // per component: offset (in floats) of its 64 pre-scaled multipliers
    uint32_t tiles_x, tiles_y;            // colour-kernel tiles (128 px x 8*vmax rows)
    // gather path only (kind == kKindGeneric): REF placement / generic sampling
    uint64_t map_off;                     // u32 offset of this shape's placement map (ncomp planes of map_plane entries)
    uint64_t smp_off;                     // float offset of this image's per-block IDCT samples
    uint32_t map_plane;                   // entries per component plane (W*H rounded up to 4)
    uint32_t pad1;
};
constexpr uint32_t kMapNone = 0xffffffffu;  // placement-map entry of a pixel no block was ever written to

enum ImgKind : uint8_t { kKindGray = 0, kKind444 = 1, kKind422 = 2, kKind420 = 3, kKind440 = 4, kKindGeneric = 5 };

// Written by the pre-pass for each image.
struct ImgDyn {
    uint32_t stream_bits;   // length of the compacted (unstuffed, marker-free) stream
    uint32_t nseg;          // restart intervals found (RST markers + 1)
    uint32_t status;        // kSt* bits
    uint32_t bits_consumed; // bit position after the last decoded MCU -> bytes_read
};

// Synchronisation record of one subsequence.
struct SubInfo {
    uint32_t p;     // bit position of the first symbol starting at or after the subsequence end
    uint32_t czf;   // bits 0-5 z, 6-9 c (block in MCU), 16 crossed, 17 bad
    int32_t n;      // coefficient positions advanced (relative) — or absolute position if crossed;
                    // after the scan kernel: absolute position at the end of the subsequence
    int32_t dc[3];  // sum of DC differences per component (relative/absolute like n)
    uint32_t pad[2];
};
constexpr uint32_t kCzMask = 0x3ffu;
constexpr uint32_t kCrossed = 1u << 16;

// ------------------------------------------------------------ zigzag mappings
// decoder.rs:404-407: ZIGZAG_INDICES[k] = natural (row-major v*8+u) index of zigzag position k.
#ifdef __CUDACC__
#define JPGPU_CONST_TABLE __device__ __constant__
#else
#define JPGPU_CONST_TABLE static const
#endif

static const uint8_t kZigzagNaturalHost[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// Coefficients live in HBM in "column-major natural" order: position u*8+v holds
// F[v][u] (u = horizontal frequency), so that the 16 bytes one lane loads are one
// column — the input of the vertical IDCT pass.  The mapping from zigzag index is
// folded into the entropy decoder's store address.
JPGPU_HD int zigzag_to_colmajor(int k, const uint8_t* zz_nat) {
    int nat = zz_nat[k];
    return ((nat & 7) << 3) | (nat >> 3);
}

// ------------------------------------------------------------------ bit reader
// The compacted stream is stored as 32-bit words holding 4 stream bytes each,
// first byte in the most significant position, so a word IS the next 32 bits.
struct BitReader {
    const uint32_t* w;
    uint32_t widx;   // next word to load
    uint32_t avail;  // valid bits at the top of buf
    uint64_t buf;

    JPGPU_HD void seek(uint32_t p) {
        widx = p >> 5;
        uint32_t off = p & 31;
        uint64_t a = w[widx], b = w[widx + 1];
        buf = ((a << 32) | b) << off;
        avail = 64 - off;
        widx += 2;
    }
    JPGPU_HD uint32_t pos() const { return widx * 32 - avail; }
    JPGPU_HD uint32_t peek() const { return (uint32_t)(buf >> 32); }
    JPGPU_HD void skip(uint32_t n) { buf <<= n; avail -= n; }  // n <= 32
    JPGPU_HD void refill() {
        if (avail <= 32) {
            buf |= (uint64_t)w[widx++] << (32 - avail);
            avail += 32;
        }
    }
};

// huffman.rs:211-227 next_code for codes longer than kLutBits: canonical prefix decode
// (T.81 F.2.2.3). Returns a packed entry, 0 if no code matches.
JPGPU_HD uint32_t huff_slow(const HuffLut& t, uint32_t peek32) {
    uint32_t code16 = peek32 >> 16;
#pragma unroll 1
    for (int l = kLutBits + 1; l <= 16; l++) {
        int32_t code = (int32_t)(code16 >> (16 - l));
        if (code <= t.maxcode[l]) return make_entry(t.vals[(t.valoff[l] + code) & 255], (uint32_t)l, t.is_dc != 0);
    }
    return 0;
}

// huffman.rs:256-268 value_correction (T.81 F.2.2.1 EXTEND)
JPGPU_HD int32_t extend(uint32_t v, uint32_t size) {
    return (size && v < (1u << (size - 1))) ? (int32_t)v - (int32_t)(1u << size) + 1 : (int32_t)v;
}

// ------------------------------------------------------------ decoder state
struct DecCtx {               // per-image constants of the entropy decoder
    const uint32_t* words;    // compacted stream of this image
    const uint32_t* seg;      // seg[k] = first bit of restart interval k; seg[nseg] = stream_bits
    uint32_t nseg;
    uint32_t stream_bits;
    uint32_t seg_units;
    int32_t nblk;             // blocks per MCU
    const HuffLut* luts;      // slot array
    const uint8_t* blk_comp;
    const uint8_t* blk_dc_slot;
    const uint8_t* blk_ac_slot;
};

struct DecState {
    BitReader br;
    int32_t g;        // coefficient position: (block index << 6) | zigzag index; relative or absolute
    int32_t c;        // block index within the MCU
    uint32_t seg;     // current restart interval
    uint32_t seg_end; // its end bit
    int32_t dc0, dc1, dc2;  // DC sums / predictors per component
    uint32_t flags;   // kCrossed | kSt* bits
};

// Largest k with seg[k] <= p.
JPGPU_HD uint32_t find_segment(const DecCtx& cx, uint32_t p) {
    uint32_t lo = 0, hi = cx.nseg;  // invariant: seg[lo] <= p < seg[hi] (seg[nseg] = stream_bits, p < stream_bits)
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (cx.seg[mid] <= p) lo = mid; else hi = mid;
    }
    return lo;
}

// Establish a decoder at bit p with block/zigzag state (c, z) and predictors.
// A decoder standing exactly on the first bit of a restart interval is in a
// known absolute state whatever it was told (predictors 0, MCU boundary).
JPGPU_HD void init_state(const DecCtx& cx, DecState& st, uint32_t p, int32_t g, int32_t c, int32_t d0, int32_t d1,
                         int32_t d2) {
    st.flags = 0;
    if (p >= cx.stream_bits) {  // nothing to decode
        st.seg = cx.nseg ? cx.nseg - 1 : 0;
        st.seg_end = cx.stream_bits;
        st.br.w = cx.words; st.br.widx = (p >> 5) + 2; st.br.avail = 64 - (p & 31); st.br.buf = 0;
        st.g = g; st.c = c; st.dc0 = d0; st.dc1 = d1; st.dc2 = d2;
        return;
    }
    uint32_t k = find_segment(cx, p);
    st.seg = k;
    st.seg_end = cx.seg[k + 1];
    st.br.w = cx.words;
    st.br.seek(p);
    if (cx.seg[k] == p) {
        st.g = (int32_t)(k * cx.seg_units);
        st.c = 0;
        st.dc0 = st.dc1 = st.dc2 = 0;
        st.flags = kCrossed;
    } else {
        st.g = g; st.c = c; st.dc0 = d0; st.dc1 = d1; st.dc2 = d2;
    }
}

// Decode symbols that START before end_bit (and, when WRITE, while g < g_limit).
// WRITE = false: synchronisation pass, only state is tracked.
// WRITE = true : coefficients are stored to coefs[(g & ~63) + store_pos[z]] (buffer pre-zeroed).
template <bool WRITE>
JPGPU_HD void decode_span(const DecCtx& cx, DecState& st, uint32_t end_bit, int32_t g_limit, int16_t* coefs,
                          const uint8_t* store_pos) {
    BitReader br = st.br;
    int32_t g = st.g, c = st.c;
    int32_t dc0 = st.dc0, dc1 = st.dc1, dc2 = st.dc2;
    uint32_t seg_end = st.seg_end, flags = st.flags;
    uint32_t p = br.pos();
    if (end_bit > cx.stream_bits) end_bit = cx.stream_bits;
    const HuffLut* tdc = cx.luts + cx.blk_dc_slot[c];
    const HuffLut* tac = cx.luts + cx.blk_ac_slot[c];
    int32_t comp = cx.blk_comp[c];
#pragma unroll 1
    while (p < end_bit) {
        if (WRITE && g >= g_limit) break;
        br.refill();
        if (p + 8 > seg_end) {  // fewer than 8 bits left in this restart interval (or already past it)
            bool cross = p >= seg_end;
            if (!cross) {
                uint32_t rem = seg_end - p;  // 1..7 pad bits must all be 1 (T.81 F.1.2.3)
                cross = (br.peek() >> (32 - rem)) == ((1u << rem) - 1u);
            }
            if (cross) {
                if (st.seg + 1 >= cx.nseg) {  // end of the entropy-coded data
                    br.seek(seg_end);
                    p = seg_end;
                    break;
                }
                st.seg += 1;
                p = seg_end;
                seg_end = cx.seg[st.seg + 1];
                br.seek(p);
                g = (int32_t)(st.seg * cx.seg_units);
                c = 0;
                tdc = cx.luts + cx.blk_dc_slot[0]; tac = cx.luts + cx.blk_ac_slot[0]; comp = cx.blk_comp[0];
                dc0 = dc1 = dc2 = 0;
                flags |= kCrossed;
                continue;
            }
        }
        const uint32_t peek = br.peek();
        const int32_t z = g & 63;
        const HuffLut* t = z ? tac : tdc;
        uint32_t e = t->fast[peek >> (32 - kLutBits)];
        if (e == 0) {
            e = huff_slow(*t, peek);
            if (e == 0) { flags |= kStBadCode; e = kBadEntry; }  // huffman.rs:156/162 panic
        }
        const uint32_t len = (e >> 8) & 31u, tb = (e >> 13) & 63u;
        const int32_t adv = (int32_t)((e >> 19) & 127u);
        if (WRITE || z == 0) {
            const uint32_t size = tb - len;
            const uint32_t v = size ? ((peek << len) >> (32 - size)) : 0u;
            const int32_t val = extend(v, size);
            if (z == 0) {  // DC difference -> predictor (decoder.rs:208-210)
                if (e & (1u << 26)) flags |= kStDcSize;
                int32_t pred;
                if (comp == 0) { dc0 += val; pred = dc0; } else if (comp == 1) { dc1 += val; pred = dc1; } else { dc2 += val; pred = dc2; }
                if (WRITE && pred != 0) coefs[(g & ~63) + store_pos[0]] = (int16_t)pred;
            } else if (WRITE && val != 0) {  // huffman.rs:183-189: min(run, 64 - len - 1) zeros, then the value
                const int32_t pos = z + adv - 1 < 63 ? z + adv - 1 : 63;
                coefs[(g & ~63) + store_pos[pos]] = (int16_t)val;
            }
        }
        br.skip(tb);
        p += tb;
        if (z + adv >= 64) {  // block complete (EOB, ZRL/run past the end, or coefficient 63)
            g = (g | 63) + 1;
            c += 1;
            if (c == cx.nblk) c = 0;
            tdc = cx.luts + cx.blk_dc_slot[c]; tac = cx.luts + cx.blk_ac_slot[c]; comp = cx.blk_comp[c];
        } else {
            g += adv;
        }
    }
    st.br = br; st.g = g; st.c = c; st.dc0 = dc0; st.dc1 = dc1; st.dc2 = dc2;
    st.seg_end = seg_end; st.flags = flags;
}

JPGPU_HD uint32_t pack_czf(const DecState& st) {
    return (uint32_t)(st.g & 63) | ((uint32_t)st.c << 6) | (st.flags & kCrossed) | ((st.flags & kStBadCode) ? (1u << 17) : 0u);
}

// Prepare the per-subsequence accumulators before decoding the next subsequence
// with a carried state: positions become relative to "now".
JPGPU_HD void begin_subsequence(DecState& st, int32_t& g_base) {
    st.flags &= ~kCrossed;
    st.dc0 = st.dc1 = st.dc2 = 0;
    g_base = st.g;
}

// Summarise the subsequence just decoded.
JPGPU_HD void summarise(const DecState& st, int32_t g_base, SubInfo& out) {
    out.p = st.br.pos();
    out.czf = pack_czf(st);
    out.n = (st.flags & kCrossed) ? st.g : st.g - g_base;
    out.dc[0] = st.dc0; out.dc[1] = st.dc1; out.dc[2] = st.dc2;
}

// ------------------------------------------------------------------- IDCT
// Scaled 8-point inverse DCT (Arai-Agui-Nakajima factorisation, float): inputs are
// coefficients pre-multiplied by aan[k] (and by 1/8 over the two passes); 5 multiplies.
// Replaces the O(N^4) direct form of transform.rs:55-87; results differ from it only
// by float rounding (<= 1 LSB after truncation, measured in the parity tests).
JPGPU_HD void idct8(float& x0, float& x1, float& x2, float& x3, float& x4, float& x5, float& x6, float& x7) {
    // even part
    float t10 = x0 + x4, t11 = x0 - x4;
    float t13 = x2 + x6;
    float t12 = (x2 - x6) * 1.414213562f - t13;
    float e0 = t10 + t13, e3 = t10 - t13, e1 = t11 + t12, e2 = t11 - t12;
    // odd part
    float z13 = x5 + x3, z10 = x5 - x3, z11 = x1 + x7, z12 = x1 - x7;
    float o7 = z11 + z13;
    float t11o = (z11 - z13) * 1.414213562f;
    float z5 = (z10 + z12) * 1.847759065f;
    float t10o = z5 - z12 * 1.082392200f;
    float t12o = z5 - z10 * 2.613125930f;
    float o6 = t12o - o7;
    float o5 = t11o - o6;
    float o4 = t10o - o5;
    x0 = e0 + o7; x7 = e0 - o7;
    x1 = e1 + o6; x6 = e1 - o6;
    x2 = e2 + o5; x5 = e2 - o5;
    x3 = e3 + o4; x4 = e3 - o4;
}

static const double kAanScale[8] = {1.0, 1.387039845, 1.306562965, 1.175875602, 1.0, 0.785694958, 0.541196100, 0.275899379};

}  // namespace jpgpu
