// jpgenc.cpp — in-repo baseline-sequential JPEG encoder + deterministic synthetic
// image generator (host only, g++).  It produces the offline inputs the north-star
// asks for: files restricted to the feature subset the reference decoder accepts
// (SOI, APP0, DQT, SOF0, DHT, SOS, EOI; 8-bit; H,V in {1,2}; one interleaved scan;
// Annex-K Huffman tables, which contain no 1-bit code — reference huffman.rs:61,212),
// plus, optionally, DRI/RSTn for the restart-interval extension corpus, image-specific
// ("optimised") Huffman tables built by T.81 K.2, and 16-bit quantisation tables.
//
// It also returns the quantised coefficients it entropy-coded (zigzag order,
// absolute DC, per component in decode order), which is decoder-independent ground
// truth for the bit-exact coefficient gate.
//
// The forward DCT here is the textbook separable DCT-II (T.81 A.3.3); it is not a
// port of reference src/transform.rs:18-53 (dead code there).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

// ---------------------------------------------------------------- Annex K tables
const uint8_t kQLum[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                           14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                           18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                           49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kQChr[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                           24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                           99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                           99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

const uint8_t kDcLumBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChrBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcLumVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t kAcChrBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcChrVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
    0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

const int kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                         41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                         30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct EncTable {
    uint16_t code[256];
    uint8_t size[256];
};

// T.81 Annex C: canonical code assignment
void make_enc_table(const uint8_t bits[16], const uint8_t* vals, EncTable* t) {
    std::memset(t, 0, sizeof *t);
    uint32_t code = 0;
    int k = 0;
    for (int len = 1; len <= 16; len++) {
        for (int i = 0; i < bits[len - 1]; i++, k++) {
            t->code[vals[k]] = (uint16_t)code;
            t->size[vals[k]] = (uint8_t)len;
            code++;
        }
        code <<= 1;
    }
}

struct BitWriter {
    std::vector<uint8_t>& out;
    uint64_t acc = 0;
    int nbits = 0;
    explicit BitWriter(std::vector<uint8_t>& o) : out(o) {}
    inline void emit_byte(uint8_t b) {
        out.push_back(b);
        if (b == 0xff) out.push_back(0x00);  // byte stuffing (T.81 F.1.2.3)
    }
    inline void put(uint32_t v, int n) {
        if (!n) return;
        acc = (acc << n) | (v & ((1u << n) - 1));
        nbits += n;
        while (nbits >= 8) {
            emit_byte((uint8_t)(acc >> (nbits - 8)));
            nbits -= 8;
        }
    }
    inline void flush_ones() {  // pad the final byte with 1-bits
        if (nbits) put((1u << (8 - nbits)) - 1, 8 - nbits);
        acc = 0;
        nbits = 0;
    }
};

inline int bit_size(int v) {
    v = v < 0 ? -v : v;
    int n = 0;
    while (v) { n++; v >>= 1; }
    return n;
}

void put_marker_seg(std::vector<uint8_t>& o, uint8_t m, const std::vector<uint8_t>& payload) {
    o.push_back(0xff);
    o.push_back(m);
    size_t len = payload.size() + 2;
    o.push_back((uint8_t)(len >> 8));
    o.push_back((uint8_t)len);
    o.insert(o.end(), payload.begin(), payload.end());
}

void scaled_qtable(const uint8_t* base, int quality, uint16_t out[64], bool wide = false) {
    quality = std::max(1, std::min(100, quality));
    int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;  // libjpeg's jpeg_quality_scaling
    for (int i = 0; i < 64; i++) {
        int v = (base[i] * scale + 50) / 100;
        out[i] = (uint16_t)std::max(1, std::min(wide ? 32767 : 255, v));  // 8-bit tables clamp at 255 (baseline)
    }
}

// DCT-II basis: C[u][x] = a(u)/2 * cos((2x+1)u*pi/16)
struct DctBasis {
    double c[8][8];
    DctBasis() {
        for (int u = 0; u < 8; u++)
            for (int x = 0; x < 8; x++)
                c[u][x] = (u == 0 ? std::sqrt(0.5) : 1.0) * 0.5 * std::cos((2 * x + 1) * u * M_PI / 16.0);
    }
};
const DctBasis kBasis;

void fdct8x8(const double in[64], double out[64]) {
    double tmp[64];
    for (int y = 0; y < 8; y++)
        for (int u = 0; u < 8; u++) {
            double s = 0;
            for (int x = 0; x < 8; x++) s += kBasis.c[u][x] * in[y * 8 + x];
            tmp[y * 8 + u] = s;
        }
    for (int v = 0; v < 8; v++)
        for (int u = 0; u < 8; u++) {
            double s = 0;
            for (int y = 0; y < 8; y++) s += kBasis.c[v][y] * tmp[y * 8 + u];
            out[v * 8 + u] = s;
        }
}

// SplitMix64
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double gauss() {  // Box-Muller
        double u1 = uniform(), u2 = uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * M_PI * u2);
    }
};

// T.81 K.2 (as in libjpeg's jpeg_gen_optimal_table): code lengths from symbol frequencies, limited to 16 bits,
// with one reserved code point so that no code is all ones.  vals gets the symbols in order of code length.
void gen_optimal_table(const long freq_in[256], uint8_t bits_out[16], std::vector<uint8_t>& vals) {
    long freq[257];
    int codesize[257], others[257];
    for (int i = 0; i < 256; i++) freq[i] = freq_in[i];
    freq[256] = 1;
    for (int i = 0; i < 257; i++) { codesize[i] = 0; others[i] = -1; }
    for (;;) {
        int c1 = -1, c2 = -1;
        long v = 1000000000L;
        for (int i = 0; i <= 256; i++) if (freq[i] && freq[i] <= v) { v = freq[i]; c1 = i; }
        v = 1000000000L;
        for (int i = 0; i <= 256; i++) if (freq[i] && freq[i] <= v && i != c1) { v = freq[i]; c2 = i; }
        if (c2 < 0) break;
        freq[c1] += freq[c2];
        freq[c2] = 0;
        codesize[c1]++;
        while (others[c1] >= 0) { c1 = others[c1]; codesize[c1]++; }
        others[c1] = c2;
        codesize[c2]++;
        while (others[c2] >= 0) { c2 = others[c2]; codesize[c2]++; }
    }
    int bits[33] = {0};
    for (int i = 0; i <= 256; i++) if (codesize[i]) bits[std::min(codesize[i], 32)]++;
    for (int i = 32; i > 16; i--)
        while (bits[i] > 0) {
            int j = i - 2;
            while (bits[j] == 0) j--;
            bits[i] -= 2; bits[i - 1]++; bits[j + 1] += 2; bits[j]--;
        }
    int i = 16;
    while (bits[i] == 0) i--;
    bits[i]--;  // the reserved code point
    for (int l = 1; l <= 16; l++) bits_out[l - 1] = (uint8_t)bits[l];
    vals.clear();
    for (int l = 1; l <= 32; l++)
        for (int sym = 0; sym < 256; sym++) if (codesize[sym] == l) vals.push_back((uint8_t)sym);
}

}  // namespace

extern "C" {

// Deterministic synthetic image: per channel a sum of three
// low-frequency 2-D sinusoids with seeded frequency/phase plus Gaussian noise.
void jpgenc_synth_rgb(uint32_t seed, int width, int height, double noise_sigma, uint8_t* rgb) {
    Rng rng(0x5EED000000000000ull ^ ((uint64_t)seed * 0x9E3779B97F4A7C15ull));
    struct Wave { double fx, fy, ph, amp; } w[3][3];
    double base[3];
    for (int c = 0; c < 3; c++) {
        base[c] = 96.0 + 64.0 * rng.uniform();
        for (int k = 0; k < 3; k++) {
            w[c][k].fx = (0.5 + 5.5 * rng.uniform()) * 2.0 * M_PI / std::max(width, 1);
            w[c][k].fy = (0.5 + 5.5 * rng.uniform()) * 2.0 * M_PI / std::max(height, 1);
            w[c][k].ph = 2.0 * M_PI * rng.uniform();
            w[c][k].amp = 12.0 + 28.0 * rng.uniform();
        }
    }
    std::vector<double> sx(3 * 3 * (size_t)width), cx(3 * 3 * (size_t)width);
    for (int c = 0; c < 3; c++)
        for (int k = 0; k < 3; k++)
            for (int x = 0; x < width; x++) {
                sx[(c * 3 + k) * (size_t)width + x] = std::sin(w[c][k].fx * x + w[c][k].ph);
                cx[(c * 3 + k) * (size_t)width + x] = std::cos(w[c][k].fx * x + w[c][k].ph);
            }
    for (int y = 0; y < height; y++) {
        double sy[3][3], cy[3][3];
        for (int c = 0; c < 3; c++)
            for (int k = 0; k < 3; k++) { sy[c][k] = std::sin(w[c][k].fy * y); cy[c][k] = std::cos(w[c][k].fy * y); }
        for (int x = 0; x < width; x++) {
            for (int c = 0; c < 3; c++) {
                double v = base[c];
                for (int k = 0; k < 3; k++)  // sin(a+b) = sin a cos b + cos a sin b
                    v += w[c][k].amp * (sx[(c * 3 + k) * (size_t)width + x] * cy[c][k] +
                                        cx[(c * 3 + k) * (size_t)width + x] * sy[c][k]);
                // cheap approx-Gaussian noise: sum of 4 uniforms (Irwin-Hall), variance 4/12
                if (noise_sigma > 0) {
                    uint64_t r = rng.next();
                    double u = ((r & 0xffff) + ((r >> 16) & 0xffff) + ((r >> 32) & 0xffff) + ((r >> 48) & 0xffff)) *
                                   (1.0 / 65536.0) - 2.0;
                    v += noise_sigma * u * 1.7320508075688772;  // /sqrt(1/3)
                }
                int iv = (int)std::lround(v);
                rgb[((size_t)y * width + x) * 3 + c] = (uint8_t)std::max(0, std::min(255, iv));
            }
        }
    }
}

size_t jpgenc_max_size(int width, int height) {
    return (size_t)width * height * 3 + ((size_t)width * height) / 2 + 65536;
}

// Encode an interleaved RGB image (or its luma when gray != 0).
//   hy, vy: luma sampling factors in {1,2}; chroma is always (1,1).
//   restart_interval: 0 = no DRI (inside the reference's subset); >0 = MCUs per interval.
//   coef_dump (optional): ncomp consecutive arrays, component c holding nblocks[c]*64 int16
//   (zigzag order, absolute DC) in decode order; coef_cap in int16 units.
// Returns the number of bytes written, or 0 if out_cap / coef_cap is too small.
// flags: 1 = image-specific Huffman tables (T.81 K.2; may contain 1-bit codes, which the reference cannot decode),
//        2 = 16-bit quantisation tables (Pq = 1) with unclamped libjpeg scaling,
//        4 = NON-INTERLEAVED: one scan per component (T.81 A.2.2: blocks in raster order over the component's own
//            ceil(Xc/8) x ceil(Yc/8) grid, no MCU padding), Annex-K tables (flag 1 is ignored), restart intervals count
//            blocks.  The reference returns after the first scan (mod.rs:416-417); coef_dump then holds every component
//            in that raster order.
size_t jpgenc_encode_ex(const uint8_t* rgb, int width, int height, int gray, int hy, int vy, int quality,
                        int restart_interval, int flags, uint8_t* out, size_t out_cap, int16_t* coef_dump,
                        size_t coef_cap, size_t* nblocks /*[3]*/) {
    const int ncomp = gray ? 1 : 3;
    if (gray) { hy = 1; vy = 1; }
    const int H[3] = {hy, 1, 1}, V[3] = {vy, 1, 1};
    const int mcuw = 8 * hy, mcuh = 8 * vy;
    const int mcux = (width + mcuw - 1) / mcuw, mcuy = (height + mcuh - 1) / mcuh;
    const int pw = mcux * mcuw, ph = mcuy * mcuh;  // padded luma size

    // colour conversion (JFIF BT.601 full range), edge replication into the MCU padding
    std::vector<float> Y((size_t)pw * ph), Cb, Cr;
    if (!gray) { Cb.resize((size_t)pw * ph); Cr.resize((size_t)pw * ph); }
    for (int y = 0; y < ph; y++) {
        int sy = std::min(y, height - 1);
        for (int x = 0; x < pw; x++) {
            int sx = std::min(x, width - 1);
            const uint8_t* p = rgb + ((size_t)sy * width + sx) * 3;
            double r = p[0], g = p[1], b = p[2];
            Y[(size_t)y * pw + x] = (float)(0.299 * r + 0.587 * g + 0.114 * b);
            if (!gray) {
                Cb[(size_t)y * pw + x] = (float)(-0.168735892 * r - 0.331264108 * g + 0.5 * b + 128.0);
                Cr[(size_t)y * pw + x] = (float)(0.5 * r - 0.418687589 * g - 0.081312411 * b + 128.0);
            }
        }
    }
    // box down-sampling of chroma
    const int cw = pw / hy, chh = ph / vy;
    std::vector<float> cbs, crs;
    if (!gray) {
        cbs.resize((size_t)cw * chh);
        crs.resize((size_t)cw * chh);
        for (int y = 0; y < chh; y++)
            for (int x = 0; x < cw; x++) {
                double a = 0, b = 0;
                for (int j = 0; j < vy; j++)
                    for (int i = 0; i < hy; i++) {
                        a += Cb[(size_t)(y * vy + j) * pw + x * hy + i];
                        b += Cr[(size_t)(y * vy + j) * pw + x * hy + i];
                    }
                cbs[(size_t)y * cw + x] = (float)(a / (hy * vy));
                crs[(size_t)y * cw + x] = (float)(b / (hy * vy));
            }
    }

    const bool planar_scans = (flags & 4) != 0 && !gray;
    const bool optimise = (flags & 1) != 0 && !planar_scans, dqt16 = (flags & 2) != 0;
    uint16_t qt[2][64];
    scaled_qtable(kQLum, quality, qt[0], dqt16);
    scaled_qtable(kQChr, quality, qt[1], dqt16);

    // ---- pass 1: quantised coefficients of every block, in decode (MCU) order
    const float* planes[3] = {Y.data(), cbs.data(), crs.data()};
    const int pstride[3] = {pw, cw, cw};
    const int bpm = hy * vy + (gray ? 0 : 2);
    std::vector<int16_t> blocks((size_t)mcux * mcuy * bpm * 64);
    {
        size_t bi = 0;
        for (int my = 0; my < mcuy; my++)
            for (int mx = 0; mx < mcux; mx++)
                for (int c = 0; c < ncomp; c++) {
                    const uint16_t* q = qt[c == 0 ? 0 : 1];
                    for (int by = 0; by < V[c]; by++)
                        for (int bx = 0; bx < H[c]; bx++, bi++) {
                            double in[64], coef[64];
                            int x0 = (mx * H[c] + bx) * 8, y0 = (my * V[c] + by) * 8;
                            for (int yy = 0; yy < 8; yy++)
                                for (int xx = 0; xx < 8; xx++)
                                    in[yy * 8 + xx] = (double)planes[c][(size_t)(y0 + yy) * pstride[c] + x0 + xx] - 128.0;
                            fdct8x8(in, coef);
                            int16_t* zz = blocks.data() + bi * 64;
                            for (int k = 0; k < 64; k++) {
                                int nat = kZigzag[k];
                                int v = (int)std::lround(coef[nat] / (double)q[nat]);
                                zz[k] = (int16_t)std::max(-1023, std::min(1023, v));  // baseline ranges
                            }
                        }
                }
    }

    // ---- Huffman tables: Annex K, or built from this image's symbol statistics (T.81 K.2)
    uint8_t tbits[4][16];                // [DC lum, AC lum, DC chr, AC chr]
    std::vector<uint8_t> tvals[4];
    std::memcpy(tbits[0], kDcLumBits, 16); tvals[0].assign(kDcVals, kDcVals + 12);
    std::memcpy(tbits[1], kAcLumBits, 16); tvals[1].assign(kAcLumVals, kAcLumVals + 162);
    std::memcpy(tbits[2], kDcChrBits, 16); tvals[2].assign(kDcVals, kDcVals + 12);
    std::memcpy(tbits[3], kAcChrBits, 16); tvals[3].assign(kAcChrVals, kAcChrVals + 162);
    if (optimise) {
        long freq[4][256];
        std::memset(freq, 0, sizeof freq);
        int pred[3] = {0, 0, 0}, in_interval = 0;
        size_t bi = 0;
        for (int m = 0; m < mcux * mcuy; m++) {
            if (restart_interval > 0 && in_interval == restart_interval) { in_interval = 0; pred[0] = pred[1] = pred[2] = 0; }
            for (int c = 0; c < ncomp; c++)
                for (int k2 = 0; k2 < H[c] * V[c]; k2++, bi++) {
                    const int16_t* zz = blocks.data() + bi * 64;
                    long* fd = freq[c == 0 ? 0 : 2];
                    long* fa = freq[c == 0 ? 1 : 3];
                    fd[bit_size(zz[0] - pred[c])]++;
                    pred[c] = zz[0];
                    int run = 0;
                    for (int k = 1; k < 64; k++) {
                        if (zz[k] == 0) { run++; continue; }
                        while (run > 15) { fa[0xf0]++; run -= 16; }
                        fa[(run << 4) | bit_size(zz[k])]++;
                        run = 0;
                    }
                    if (run > 0) fa[0x00]++;
                }
            in_interval++;
        }
        for (int t = 0; t < (gray ? 2 : 4); t++) gen_optimal_table(freq[t], tbits[t], tvals[t]);
    }
    EncTable enc[4];
    for (int t = 0; t < 4; t++) make_enc_table(tbits[t], tvals[t].data(), &enc[t]);

    std::vector<uint8_t> o;
    o.reserve((size_t)width * height / 2 + 4096);
    o.push_back(0xff); o.push_back(0xd8);  // SOI
    put_marker_seg(o, 0xe0, {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0});  // APP0 JFIF 1.01
    for (int t = 0; t < (gray ? 1 : 2); t++) {  // DQT, zigzag order, 8- or 16-bit entries
        std::vector<uint8_t> p;
        p.push_back((uint8_t)((dqt16 ? 0x10 : 0x00) | t));
        for (int k = 0; k < 64; k++) {
            const uint16_t v = qt[t][kZigzag[k]];
            if (dqt16) p.push_back((uint8_t)(v >> 8));
            p.push_back((uint8_t)v);
        }
        put_marker_seg(o, 0xdb, p);
    }
    {  // SOF0
        std::vector<uint8_t> p = {8, (uint8_t)(height >> 8), (uint8_t)height, (uint8_t)(width >> 8), (uint8_t)width,
                                  (uint8_t)ncomp};
        for (int c = 0; c < ncomp; c++) {
            p.push_back((uint8_t)(c + 1));
            p.push_back((uint8_t)((H[c] << 4) | V[c]));
            p.push_back((uint8_t)(c == 0 ? 0 : 1));
        }
        put_marker_seg(o, 0xc0, p);
    }
    for (int t = 0; t < (gray ? 2 : 4); t++) {  // DHT: Tc = t & 1, Th = t >> 1
        std::vector<uint8_t> p;
        p.push_back((uint8_t)(((t & 1) << 4) | (t >> 1)));
        p.insert(p.end(), tbits[t], tbits[t] + 16);
        p.insert(p.end(), tvals[t].begin(), tvals[t].end());
        put_marker_seg(o, 0xc4, p);
    }
    if (restart_interval > 0)
        put_marker_seg(o, 0xdd, {(uint8_t)(restart_interval >> 8), (uint8_t)restart_interval});
    if (planar_scans) {
        // one scan per component, blocks in raster order over the component's own block grid
        size_t dump_at = 0;
        for (int c = 0; c < ncomp; c++) {
            const int wc = (width * H[c] + hy - 1) / hy, hc = (height * V[c] + vy - 1) / vy;
            const int nbx = (wc + 7) / 8, nby = (hc + 7) / 8;
            if (nblocks) nblocks[c] = (size_t)nbx * nby;
            if (coef_dump && coef_cap < dump_at + (size_t)nbx * nby * 64) return 0;
            put_marker_seg(o, 0xda, {1, (uint8_t)(c + 1), (uint8_t)(c == 0 ? 0x00 : 0x11), 0, 63, 0});
            BitWriter bw(o);
            const EncTable& dct = enc[c == 0 ? 0 : 2];
            const EncTable& act = enc[c == 0 ? 1 : 3];
            int first = 0;
            for (int k = 0; k < c; k++) first += H[k] * V[k];
            int pred = 0, rst_count = 0, in_interval = 0;
            for (int by = 0; by < nby; by++)
                for (int bx = 0; bx < nbx; bx++) {
                    if (restart_interval > 0 && in_interval == restart_interval) {
                        bw.flush_ones();
                        o.push_back(0xff);
                        o.push_back((uint8_t)(0xd0 + (rst_count & 7)));
                        rst_count++;
                        in_interval = 0;
                        pred = 0;
                    }
                    const size_t mcu = (size_t)(by / V[c]) * mcux + bx / H[c];
                    const int16_t* zz = blocks.data() + (mcu * bpm + first + (by % V[c]) * H[c] + bx % H[c]) * 64;
                    if (coef_dump) { std::memcpy(coef_dump + dump_at, zz, 64 * sizeof(int16_t)); dump_at += 64; }
                    int diff = zz[0] - pred;
                    pred = zz[0];
                    int sz0 = bit_size(diff);
                    bw.put(dct.code[sz0], dct.size[sz0]);
                    if (sz0) bw.put((uint32_t)(diff < 0 ? diff - 1 : diff), sz0);
                    int run = 0;
                    for (int k = 1; k < 64; k++) {
                        if (zz[k] == 0) { run++; continue; }
                        while (run > 15) { bw.put(act.code[0xf0], act.size[0xf0]); run -= 16; }
                        int sz = bit_size(zz[k]);
                        int sym = (run << 4) | sz;
                        bw.put(act.code[sym], act.size[sym]);
                        bw.put((uint32_t)(zz[k] < 0 ? zz[k] - 1 : zz[k]), sz);
                        run = 0;
                    }
                    if (run > 0) bw.put(act.code[0x00], act.size[0x00]);
                    in_interval++;
                }
            bw.flush_ones();
        }
        o.push_back(0xff); o.push_back(0xd9);  // EOI
        if (o.size() > out_cap) return 0;
        std::memcpy(out, o.data(), o.size());
        return o.size();
    }
    {  // SOS
        std::vector<uint8_t> p = {(uint8_t)ncomp};
        for (int c = 0; c < ncomp; c++) {
            p.push_back((uint8_t)(c + 1));
            p.push_back((uint8_t)(c == 0 ? 0x00 : 0x11));
        }
        p.push_back(0); p.push_back(63); p.push_back(0);
        put_marker_seg(o, 0xda, p);
    }

    // coefficient dump bookkeeping
    size_t nb[3] = {(size_t)mcux * mcuy * hy * vy, (size_t)mcux * mcuy, (size_t)mcux * mcuy};
    size_t dump_off[3] = {0, nb[0] * 64, (nb[0] + nb[1]) * 64};
    size_t dump_need = 0;
    for (int c = 0; c < ncomp; c++) dump_need += nb[c] * 64;
    if (coef_dump && coef_cap < dump_need) return 0;
    size_t dump_n[3] = {0, 0, 0};
    if (nblocks) for (int c = 0; c < 3; c++) nblocks[c] = c < ncomp ? nb[c] : 0;

    // ---- pass 2: entropy coding
    BitWriter bw(o);
    int pred[3] = {0, 0, 0};
    int rst_count = 0, mcus_in_interval = 0;
    size_t bi = 0;
    for (int m = 0; m < mcux * mcuy; m++) {
        if (restart_interval > 0 && mcus_in_interval == restart_interval) {
            bw.flush_ones();
            o.push_back(0xff);
            o.push_back((uint8_t)(0xd0 + (rst_count & 7)));
            rst_count++;
            mcus_in_interval = 0;
            pred[0] = pred[1] = pred[2] = 0;
        }
        for (int c = 0; c < ncomp; c++) {
            const EncTable& dct = enc[c == 0 ? 0 : 2];
            const EncTable& act = enc[c == 0 ? 1 : 3];
            for (int k2 = 0; k2 < H[c] * V[c]; k2++, bi++) {
                const int16_t* zz = blocks.data() + bi * 64;
                if (coef_dump) {
                    std::memcpy(coef_dump + dump_off[c] + dump_n[c] * 64, zz, 64 * sizeof(int16_t));
                    dump_n[c]++;
                }
                // DC
                int diff = zz[0] - pred[c];
                pred[c] = zz[0];
                int sz0 = bit_size(diff);
                bw.put(dct.code[sz0], dct.size[sz0]);
                if (sz0) bw.put((uint32_t)(diff < 0 ? diff - 1 : diff), sz0);
                // AC
                int run = 0;
                for (int k = 1; k < 64; k++) {
                    if (zz[k] == 0) { run++; continue; }
                    while (run > 15) { bw.put(act.code[0xf0], act.size[0xf0]); run -= 16; }
                    int sz = bit_size(zz[k]);
                    int sym = (run << 4) | sz;
                    bw.put(act.code[sym], act.size[sym]);
                    bw.put((uint32_t)(zz[k] < 0 ? zz[k] - 1 : zz[k]), sz);
                    run = 0;
                }
                if (run > 0) bw.put(act.code[0x00], act.size[0x00]);
            }
        }
        mcus_in_interval++;
    }
    bw.flush_ones();
    o.push_back(0xff); o.push_back(0xd9);  // EOI
    if (o.size() > out_cap) return 0;
    std::memcpy(out, o.data(), o.size());
    return o.size();
}

size_t jpgenc_encode(const uint8_t* rgb, int width, int height, int gray, int hy, int vy, int quality,
                     int restart_interval, uint8_t* out, size_t out_cap, int16_t* coef_dump, size_t coef_cap,
                     size_t* nblocks /*[3]*/) {
    return jpgenc_encode_ex(rgb, width, height, gray, hy, vy, quality, restart_interval, 0, out, out_cap, coef_dump,
                            coef_cap, nblocks);
}

// synth + encode in one call (used by the batch generators; thread-safe)
size_t jpgenc_synth_encode_ex(uint32_t seed, int width, int height, double noise_sigma, int gray, int hy, int vy,
                              int quality, int restart_interval, int flags, uint8_t* out, size_t out_cap,
                              int16_t* coef_dump, size_t coef_cap, size_t* nblocks) {
    std::vector<uint8_t> rgb((size_t)width * height * 3);
    jpgenc_synth_rgb(seed, width, height, noise_sigma, rgb.data());
    return jpgenc_encode_ex(rgb.data(), width, height, gray, hy, vy, quality, restart_interval, flags, out, out_cap,
                            coef_dump, coef_cap, nblocks);
}

size_t jpgenc_synth_encode(uint32_t seed, int width, int height, double noise_sigma, int gray, int hy, int vy,
                           int quality, int restart_interval, uint8_t* out, size_t out_cap, int16_t* coef_dump,
                           size_t coef_cap, size_t* nblocks) {
    return jpgenc_synth_encode_ex(seed, width, height, noise_sigma, gray, hy, vy, quality, restart_interval, 0, out, out_cap,
                                  coef_dump, coef_cap, nblocks);
}

}  // extern "C"
