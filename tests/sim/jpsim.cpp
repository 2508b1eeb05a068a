// jpsim.cpp — CPU SIMULATION of the CUDA kernels' algorithm.  TEST INFRASTRUCTURE ONLY.
//
// It executes the same per-thread code (jpgpu_core.h: decode_span, init_state, idct8,
// ...) and the same host planner (jpgpu_host.cpp) as the product, with the kernels'
// thread/round/barrier structure replayed serially.  It lets the not-gpu test suite
// check the self-synchronising decode, the scan, the coefficient layout and the fused
// IDCT/colour arithmetic against the oracle on a box without a GPU.  It is NOT a
// fallback: nothing in jpeg_rust_b200/ or libjpgpu.so can reach it.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../jpeg_rust_b200/csrc/jpgpu_host.h"

using namespace jpgpu;

namespace {

struct SimBatch {
    HostPlan plan;
    std::vector<uint8_t> raw;
    std::vector<uint32_t> stream;
    std::vector<uint32_t> segtab;
    std::vector<SubInfo> subs;
    std::vector<ImgDyn> dyn;
    std::vector<int16_t> coefs;
    std::vector<uint8_t> rgb;
    uint8_t store_pos[64];
    // diagnostics
    uint32_t max_rounds = 0;       // most intra-sequence rounds any CTA needed
    uint32_t inter_iters = 0;      // most inter-sequence iterations any image needed
    uint64_t inter_walk = 0;       // total subsequences decoded by inter-sequence walkers
    uint64_t intra_decodes = 0;    // total subsequence decodes in the intra pass
};

// mirrors prepass_kernel's per-byte rule (the warp/CTA scan itself is GPU plumbing)
void sim_prepass(SimBatch& sb, size_t img) {
    const ImgDev& im = sb.plan.imgs[img];
    const uint8_t* in = sb.raw.data() + im.raw_off;
    const uint32_t n = im.raw_len;
    uint32_t* out = sb.stream.data() + im.stream_off;
    uint32_t* seg = sb.segtab.data() + im.seg_off;
    const bool dri = im.restart_interval != 0;
    uint32_t emitted = 0, rst_total = 0, status = 0;
    std::vector<uint8_t> bytes;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t cur = in[i], p = i ? in[i - 1] : 0, nxt = i + 1 < n ? in[i + 1] : 0;
        const bool has_next = i + 1 < n;
        bool drop = (cur == 0 && p == 0xff);
        bool r1 = false;
        if (dri) {
            r1 = cur == 0xff && has_next && (nxt & 0xf8) == 0xd0;
            const bool r2 = p == 0xff && (cur & 0xf8) == 0xd0;
            const bool fill = cur == 0xff && has_next && nxt == 0xff;
            drop = drop || r1 || r2 || fill;
        }
        if (r1) {
            if (rst_total + 1 < im.nseg_cap) seg[rst_total + 1] = emitted * 8;
            if ((nxt & 7) != (rst_total & 7)) status |= kStRestart;
            rst_total++;
        }
        if (!drop) { bytes.push_back((uint8_t)cur); emitted++; }
    }
    for (uint32_t k = 0; k < emitted; k++) {
        if ((k & 3) == 0) out[k >> 2] = 0;
        out[k >> 2] |= (uint32_t)bytes[k] << (24 - 8 * (k & 3));
    }
    uint32_t wbase = (emitted + 3) >> 2;
    for (int i = 0; i < kStreamPadWords; i++) out[wbase + i] = 0;
    uint32_t nseg = rst_total + 1;
    if (nseg != im.nseg_cap && dri) status |= kStRestart;
    if (nseg > im.nseg_cap) nseg = im.nseg_cap;
    seg[0] = 0;
    seg[nseg] = emitted * 8;
    sb.dyn[img] = ImgDyn{emitted * 8, nseg, status, 0};
}

DecCtx make_ctx(const SimBatch& sb, size_t img, const HuffLut* slots) {
    const ImgDev& im = sb.plan.imgs[img];
    DecCtx cx;
    cx.words = sb.stream.data() + im.stream_off;
    cx.seg = sb.segtab.data() + im.seg_off;
    cx.nseg = sb.dyn[img].nseg;
    cx.stream_bits = sb.dyn[img].stream_bits;
    cx.seg_units = im.seg_units;
    cx.nblk = im.blocks_per_mcu;
    cx.luts = slots;
    cx.blk_comp = im.blk_comp;
    cx.blk_dc_slot = im.blk_dc_slot;
    cx.blk_ac_slot = im.blk_ac_slot;
    return cx;
}

void load_slots(const SimBatch& sb, size_t img, std::vector<HuffLut>& slots) {
    const ImgDev& im = sb.plan.imgs[img];
    slots.resize(kMaxLutSlots);
    for (int s = 0; s < im.nslots; s++) slots[s] = sb.plan.luts[im.slot_lut[s]];
}

// mirrors sync_intra_kernel: one "CTA" per sequence, barrier-separated rounds
void sim_sync_intra(SimBatch& sb, const SeqDesc& sd) {
    std::vector<HuffLut> slots;
    load_slots(sb, sd.img, slots);
    const ImgDev& im = sb.plan.imgs[sd.img];
    const ImgDyn dyn = sb.dyn[sd.img];
    const uint32_t S = sb.plan.sub_bits;
    const uint32_t nsub = (dyn.stream_bits + S - 1) / S;
    if (sd.first_sub >= nsub) return;
    const DecCtx cx = make_ctx(sb, sd.img, slots.data());
    std::vector<DecState> st(kSeqThreads);
    std::vector<int32_t> g_base(kSeqThreads, 0);
    std::vector<char> active(kSeqThreads, 0);
    std::vector<SubInfo> s_info(kSeqThreads);
    for (uint32_t tid = 0; tid < (uint32_t)kSeqThreads; tid++) {
        const uint32_t j = sd.first_sub + tid;
        active[tid] = j < nsub;
        if (!active[tid]) continue;
        init_state(cx, st[tid], j * S, 0, 0, 0, 0, 0);
        g_base[tid] = st[tid].g;
        decode_span<false>(cx, st[tid], (j + 1) * S, 0, nullptr, nullptr);
        SubInfo mine;
        summarise(st[tid], g_base[tid], mine);
        mine.pad[0] = mine.pad[1] = 0;
        s_info[tid] = mine;
        sb.intra_decodes++;
    }
    uint32_t rounds = 1;
    for (uint32_t r = 1; r < (uint32_t)kSeqThreads; r++) {
        bool any = false;
        // within a round every thread reads/writes only s_info[tid + r]: order is irrelevant
        for (uint32_t tid = 0; tid < (uint32_t)kSeqThreads; tid++) {
            const uint32_t j = sd.first_sub + tid, tgt = tid + r;
            if (active[tid] && (tgt >= (uint32_t)kSeqThreads || j + r >= nsub)) active[tid] = 0;
            if (!active[tid]) continue;
            begin_subsequence(st[tid], g_base[tid]);
            decode_span<false>(cx, st[tid], (j + r + 1) * S, 0, nullptr, nullptr);
            SubInfo mine;
            summarise(st[tid], g_base[tid], mine);
            mine.pad[0] = mine.pad[1] = 0;
            const SubInfo old = s_info[tgt];
            const bool same = old.p == mine.p && ((old.czf ^ mine.czf) & kCzMask) == 0u;
            s_info[tgt] = mine;
            if (same) active[tid] = 0;
            sb.intra_decodes++;
            any = any || active[tid];
        }
        rounds++;
        if (!any) break;
    }
    sb.max_rounds = std::max(sb.max_rounds, rounds);
    for (uint32_t tid = 0; tid < (uint32_t)kSeqThreads; tid++) {
        const uint32_t j = sd.first_sub + tid;
        if (j < nsub) sb.subs[im.sub_off + j] = s_info[tid];
    }
}

// mirrors sync_inter_scan_kernel
void sim_sync_inter_scan(SimBatch& sb, size_t img) {
    std::vector<HuffLut> slots;
    load_slots(sb, img, slots);
    const ImgDev& im = sb.plan.imgs[img];
    const ImgDyn dyn = sb.dyn[img];
    const uint32_t S = sb.plan.sub_bits;
    const uint32_t nsub = (dyn.stream_bits + S - 1) / S;
    const uint32_t nseq = (nsub + kSeqThreads - 1) / kSeqThreads;
    const DecCtx cx = make_ctx(sb, img, slots.data());
    SubInfo* subs = sb.subs.data() + im.sub_off;
    std::vector<uint32_t> need_a(nseq + 1, 0), need_b(nseq + 1, 0);
    for (uint32_t q = 1; q < nseq; q++) need_a[q] = 1;
    uint32_t iters = 0;
    for (uint32_t iter = 0; iter < nseq; iter++) {
        bool any = false;
        iters++;
        for (uint32_t q0 = 1; q0 < nseq; q0 += 64) {
            // snapshot phase
            std::vector<SubInfo> start(64);
            std::vector<char> work(64, 0);
            for (uint32_t tid = 0; tid < 64; tid++) {
                const uint32_t q = q0 + tid;
                work[tid] = q < nseq && need_a[q];
                if (work[tid]) start[tid] = subs[q * kSeqThreads - 1];
            }
            // walk phase (threads of one pass are independent except through need_b)
            for (uint32_t tid = 0; tid < 64; tid++) {
                if (!work[tid]) continue;
                const uint32_t q = q0 + tid;
                need_a[q] = 0;
                DecState st;
                init_state(cx, st, start[tid].p, (int32_t)(start[tid].czf & 63u), (int32_t)((start[tid].czf >> 6) & 15u), 0, 0, 0);
                int32_t g_base = st.g;
                bool first = true;
                for (uint32_t t = 0; t < (uint32_t)kSeqThreads; t++) {
                    const uint32_t jj = q * kSeqThreads + t;
                    if (jj >= nsub) break;
                    if (!first) begin_subsequence(st, g_base);
                    else { g_base = st.g; st.dc0 = st.dc1 = st.dc2 = 0; }
                    first = false;
                    decode_span<false>(cx, st, (jj + 1) * S, 0, nullptr, nullptr);
                    sb.inter_walk++;
                    SubInfo mine;
                    summarise(st, g_base, mine);
                    mine.pad[0] = mine.pad[1] = 0;
                    const SubInfo old = subs[jj];
                    const bool same = old.p == mine.p && ((old.czf ^ mine.czf) & kCzMask) == 0u;
                    subs[jj] = mine;
                    if (same) break;
                    if (t == (uint32_t)kSeqThreads - 1 && q + 1 < nseq) { need_b[q + 1] = 1; any = true; }
                }
            }
        }
        if (!any) break;
        for (uint32_t q = 0; q < nseq; q++) { need_a[q] = need_b[q]; need_b[q] = 0; }
    }
    sb.inter_iters = std::max(sb.inter_iters, iters);
    // scan (serial form of the chunked scan)
    int32_t run[4] = {0, 0, 0, 0};
    for (uint32_t jj = 0; jj < nsub; jj++) {
        SubInfo& s = subs[jj];
        if (s.czf & kCrossed) { run[0] = s.n; run[1] = s.dc[0]; run[2] = s.dc[1]; run[3] = s.dc[2]; }
        else { run[0] += s.n; run[1] += s.dc[0]; run[2] += s.dc[1]; run[3] += s.dc[2]; }
        s.n = run[0]; s.dc[0] = run[1]; s.dc[1] = run[2]; s.dc[2] = run[3];
    }
}

// mirrors decode_write_kernel
void sim_decode_write(SimBatch& sb, const SeqDesc& sd) {
    std::vector<HuffLut> slots;
    load_slots(sb, sd.img, slots);
    const ImgDev& im = sb.plan.imgs[sd.img];
    const ImgDyn dyn = sb.dyn[sd.img];
    const uint32_t S = sb.plan.sub_bits;
    const uint32_t nsub = (dyn.stream_bits + S - 1) / S;
    if (sd.first_sub >= nsub) return;
    const DecCtx cx = make_ctx(sb, sd.img, slots.data());
    for (uint32_t tid = 0; tid < (uint32_t)kSeqThreads; tid++) {
        const uint32_t j = sd.first_sub + tid;
        if (j >= nsub) continue;
        DecState st;
        if (j == 0) init_state(cx, st, 0u, 0, 0, 0, 0, 0);
        else {
            const SubInfo prev = sb.subs[im.sub_off + j - 1];
            init_state(cx, st, prev.p, prev.n, (int32_t)((prev.czf >> 6) & 15u), prev.dc[0], prev.dc[1], prev.dc[2]);
        }
        const int32_t total = (int32_t)im.total_coefs;
        const int32_t g_start = st.g;
        st.flags &= ~kCrossed;
        decode_span<true>(cx, st, (j + 1) * S, total, sb.coefs.data() + im.coef_off, sb.store_pos);
        uint32_t bits = st.flags & (kStBadCode | kStDcSize);
        if (g_start < total && st.g >= total) { sb.dyn[sd.img].bits_consumed = st.br.pos(); bits |= kStDone; }
        sb.dyn[sd.img].status |= bits;
    }
}

inline uint8_t sat_u8_trunc(float x) {  // cvt.rzi.sat.u8.f32
    if (!(x > 0.0f)) return 0;
    if (x >= 255.0f) return 255;
    return (uint8_t)x;
}

// the arithmetic of block_idct (vertical pass on columns, horizontal pass on rows)
void sim_block_idct(const int16_t* src, bool valid, const float* qt, float dc_bias, float out[64]) {
    float tmp[64];
    for (int t = 0; t < 8; t++) {
        float f[8];
        for (int v = 0; v < 8; v++) f[v] = valid ? (float)src[t * 8 + v] * qt[t * 8 + v] : 0.0f;
        if (t == 0) f[0] += dc_bias;
        idct8(f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7]);
        for (int y = 0; y < 8; y++) tmp[y * 8 + t] = f[y];
    }
    for (int t = 0; t < 8; t++) {
        float* r = tmp + t * 8;
        idct8(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
        for (int x = 0; x < 8; x++) out[t * 8 + x] = r[x];
    }
}

// mirrors idct_colour_kernel (any HY, VY in {1,2}; gray)
void sim_idct_colour(SimBatch& sb, size_t img) {
    const ImgDev& im = sb.plan.imgs[img];
    const bool gray = im.kind == kKindGray;
    const int HY = gray ? 1 : im.h[0], VY = gray ? 1 : im.v[0];
    const int NY = HY * VY, NB = gray ? 1 : NY + 2;
    const int16_t* coefs = sb.coefs.data() + im.coef_off;
    uint8_t* rgb = sb.rgb.data() + im.rgb_off;
    const float* qt = sb.plan.qt.data();
    const uint32_t W = im.width, H = im.height;
    for (uint32_t my = 0; my < im.mcuy; my++)
        for (uint32_t mx = 0; mx < im.mcux; mx++) {
            const uint32_t mcu = my * im.mcux + mx;
            const bool valid = mcu < im.units;
            float cb[64], cr[64];
            if (!gray) {
                sim_block_idct(coefs + ((size_t)mcu * NB + NY) * 64, valid, qt + im.qt_off[1], 0.0f, cb);
                sim_block_idct(coefs + ((size_t)mcu * NB + NY + 1) * 64, valid, qt + im.qt_off[2], 0.0f, cr);
            }
            for (int sub = 0; sub < NY; sub++) {
                const int by = sub / HY, bx = sub % HY;
                float y[64];
                sim_block_idct(coefs + ((size_t)mcu * NB + sub) * 64, valid, qt + im.qt_off[0], 128.0f, y);
                for (int t = 0; t < 8; t++)
                    for (int x = 0; x < 8; x++) {
                        const uint32_t px = (mx * HY + bx) * 8 + x, py = (my * VY + by) * 8 + t;
                        if (px >= W || py >= H) continue;
                        uint8_t* o = rgb + ((size_t)py * W + px) * 3;
                        if (gray) {
                            o[0] = o[1] = o[2] = sat_u8_trunc(y[t * 8 + x]);
                        } else {
                            const int crow = (by * 8 + t) / VY, ccol = (bx * 8 + x) / HY;
                            const float cbv = cb[crow * 8 + ccol], crv = cr[crow * 8 + ccol], yy = y[t * 8 + x];
                            o[0] = sat_u8_trunc(fmaf(crv, 1.402f, yy));
                            o[1] = sat_u8_trunc(fmaf(cbv, -0.34413629f, fmaf(crv, -0.71413629f, yy)));
                            o[2] = sat_u8_trunc(fmaf(cbv, 1.772f, yy));
                        }
                    }
            }
        }
}

// mirrors block_idct_kernel + gather_colour_kernel (REF placement / generic sampling)
void sim_gather(SimBatch& sb, size_t img) {
    const ImgDev& im = sb.plan.imgs[img];
    const uint32_t nblk = im.units * im.blocks_per_mcu;
    const int16_t* coefs = sb.coefs.data() + im.coef_off;
    const float* qt = sb.plan.qt.data();
    std::vector<float> smp((size_t)nblk * 64);
    for (uint32_t blk = 0; blk < nblk; blk++) {
        const int comp = im.blk_comp[blk % im.blocks_per_mcu];
        sim_block_idct(coefs + (size_t)blk * 64, true, qt + im.qt_off[comp], comp == 0 ? 128.0f : 0.0f, smp.data() + (size_t)blk * 64);
    }
    const uint32_t* map = sb.plan.gmap.data() + im.map_off;
    uint8_t* rgb = sb.rgb.data() + im.rgb_off;
    const size_t npix = (size_t)im.width * im.height;
    for (size_t p = 0; p < npix; p++) {
        float v[3];
        for (int c = 0; c < 3; c++) {
            const float none = c == 0 ? 128.0f : 0.0f;
            const uint32_t m = c < im.ncomp ? map[(size_t)c * im.map_plane + p] : kMapNone;
            v[c] = m == kMapNone ? none : smp[m];
        }
        uint8_t* o = rgb + p * 3;
        if (im.ncomp == 1) {
            o[0] = o[1] = o[2] = sat_u8_trunc(v[0]);
        } else {
            o[0] = sat_u8_trunc(fmaf(v[2], 1.402f, v[0]));
            o[1] = sat_u8_trunc(fmaf(v[1], -0.34413629f, fmaf(v[2], -0.71413629f, v[0])));
            o[2] = sat_u8_trunc(fmaf(v[1], 1.772f, v[0]));
        }
    }
}

}  // namespace

extern "C" {

// Runs the whole simulated pipeline on a batch of descriptors.
//   rgb_out[i]   : W*H*3 bytes (may be NULL)
//   coef_out[i]  : reference-order coefficients (may be NULL), coef_cap[i] int16 each
//   diag[4]      : max intra rounds, max inter iterations, inter walk decodes, intra decodes
int jpsim_decode_batch(const jpgpu_image_desc* descs, size_t n, uint8_t* const* rgb_out, int16_t* const* coef_out,
                       const size_t* coef_cap, uint32_t* nblocks /* n x 4 */, int32_t* statuses, uint64_t* bytes_read,
                       uint64_t* diag, uint32_t sub_bits) {
    SimBatch sb;
    int st = build_plan(descs, n, sb.plan, sub_bits);
    if (st != JPGPU_OK) return st;
    HostPlan& p = sb.plan;
    sb.raw.assign(p.raw_bytes + 64, 0);
    sb.stream.assign(p.stream_words + 64, 0);
    sb.segtab.assign(p.seg_entries + 8, 0);
    sb.subs.assign(p.sub_entries + 1, SubInfo());
    sb.dyn.assign(n + 1, ImgDyn());
    sb.coefs.assign(p.coef_elems + 64, 0);
    sb.rgb.assign(p.rgb_bytes + 256, 0);
    for (int k = 0; k < 64; k++) sb.store_pos[k] = (uint8_t)zigzag_to_colmajor(k, kZigzagNaturalHost);
    for (size_t i = 0; i < n; i++)
        if (p.status[i] == JPGPU_OK) memcpy(sb.raw.data() + p.imgs[i].raw_off, descs[i].scan, p.imgs[i].raw_len);
    for (size_t i = 0; i < n; i++) sim_prepass(sb, i);
    for (const SeqDesc& sd : p.seqs) sim_sync_intra(sb, sd);
    for (size_t i = 0; i < n; i++) sim_sync_inter_scan(sb, i);
    for (const SeqDesc& sd : p.seqs) sim_decode_write(sb, sd);
    for (size_t i = 0; i < n; i++)
        if (p.status[i] == JPGPU_OK) { if (p.imgs[i].kind == kKindGeneric) sim_gather(sb, i); else sim_idct_colour(sb, i); }
    for (size_t i = 0; i < n; i++) {
        int32_t s = p.status[i];
        uint64_t br = 0;
        if (s == JPGPU_OK) {
            const uint32_t f = sb.dyn[i].status;
            if (f & kStDcSize) s = JPGPU_PANIC_READ_BITS_ASSERT;
            else if (f & kStBadCode) s = JPGPU_ERR_BAD_CODE;
            else if (!(f & kStDone)) s = (f & kStRestart) ? JPGPU_ERR_RESTART : JPGPU_ERR_TRUNCATED;
            br = ((uint64_t)sb.dyn[i].bits_consumed + 7) / 8;
            const ImgDev& im = p.imgs[i];
            if (rgb_out && rgb_out[i]) memcpy(rgb_out[i], sb.rgb.data() + im.rgb_off, (size_t)im.width * im.height * 3);
            if (coef_out && coef_out[i] && coef_cap[i] >= im.total_coefs)
                export_reference_order(im, sb.coefs.data() + im.coef_off, coef_out[i], nblocks + 4 * i);
        }
        if (statuses) statuses[i] = s;
        if (bytes_read) bytes_read[i] = br;
    }
    if (diag) { diag[0] = sb.max_rounds; diag[1] = sb.inter_iters; diag[2] = sb.inter_walk; diag[3] = sb.intra_decodes; }
    return JPGPU_OK;
}

// idct8-based 8x8 IDCT of natural-order dequantised coefficients, for unit tests of the math
void jpsim_idct_8x8(const float in_natural[64], float out[64]) {
    uint16_t ones[64];
    for (int k = 0; k < 64; k++) ones[k] = 1;
    float mult[64];
    build_qt_multipliers(ones, mult);  // column-major AAN multipliers for q = 1
    float tmp[64];
    for (int u = 0; u < 8; u++) {
        float f[8];
        for (int v = 0; v < 8; v++) f[v] = in_natural[v * 8 + u] * mult[u * 8 + v];
        idct8(f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7]);
        for (int y = 0; y < 8; y++) tmp[y * 8 + u] = f[y];
    }
    for (int t = 0; t < 8; t++) {
        float* r = tmp + t * 8;
        idct8(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
        for (int x = 0; x < 8; x++) out[t * 8 + x] = r[x];
    }
}

}  // extern "C"
