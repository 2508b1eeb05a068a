// jpgpu_api.cu — device part of the C ABI (include/jpgpu.h): context, batch
// planning on the device, stage launches, transfers.  There is no CPU fallback:
// without a usable sm_100 device every entry point returns JPGPU_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "jpgpu_host.h"
#include "jpgpu_kernels.cuh"

using namespace jpgpu;

// No C++ exception may cross the C ABI (std::bad_alloc from a planner vector under a memory limit would otherwise
// terminate the caller's process): every int-returning entry point is a function-try-block ending in this.
#define JPGPU_CATCH_ALL                                              \
    catch (const std::bad_alloc&) { return JPGPU_ERR_OOM; }          \
    catch (...) { return JPGPU_ERR_INVALID_ARG; }

struct jpgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // auxiliary streams on which jpgpu_batch_decode runs every other group of a batch (fork/join around the
    // caller's stream, so the call keeps its stream semantics)
    static constexpr int kAux = 3;
    cudaStream_t aux[kAux] = {nullptr, nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[kAux] = {nullptr, nullptr, nullptr};
    jpgpu_batch* single = nullptr;   // batch object jpgpu_decode() reuses from call to call (arenas only grow)
    std::string err;
};

struct jpgpu_batch {
    jpgpu_ctx* ctx = nullptr;
    size_t n = 0;
    HostPlan plan;
    std::vector<const uint8_t*> host_scan;
    std::vector<size_t> host_scan_len;
    std::vector<uint64_t> scan_offs;   // jpgpu_batch_upload_from(): where each scan lies in the staging arena
    std::vector<uint8_t> plan_blob;    // host image of the plan's device tables (one copy per plan)
    BatchDev dev;
    struct Arena { void* p = nullptr; size_t cap = 0; };
    enum { kPlanBlob, kSamples, kRaw, kDyn, kStream, kSegtab, kSubs, kSegs, kChunks, kCoefs,
           kRgb, kScanOffs, kStage, kNumArenas };
    Arena arena[kNumArenas];   // device allocations, grown on demand by jpgpu_batch_replan()
    uint64_t launches = 0;
    size_t coef_bytes = 0;
    uint8_t* own_rgb = nullptr;
    size_t ext_rgb_cap = 0;    // capacity of the caller's output arena while jpgpu_batch_set_device_output() is in force
    bool decoded = false;
    // jpgpu_batch_decode() as one CUDA graph launch: the launch parameters of a decode only change with the plan, the
    // output arena and the output format (dev_gen counts those changes), so the second decode of an unchanged batch
    // captures the whole chain - all groups, their forks and joins over the auxiliary streams - and later ones replay it.
    uint64_t dev_gen = 1, graph_gen = 0, seen_gen = 0, graph_launches = 0;
    cudaGraphExec_t graph_exec = nullptr;
    bool graph_ok = true;
};

namespace {

int fail(jpgpu_ctx* c, cudaError_t e, const char* what) {
    if (c) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
        c->err = buf;
    }
    return e == cudaErrorMemoryAllocation ? JPGPU_ERR_OOM : JPGPU_ERR_CUDA;
}

#define CK(call)                                              \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) return fail(ctx, e_, #call);   \
    } while (0)

// Makes arena `which` at least `count` elements of T large (contents undefined after growth).
template <typename T>
int dev_ensure(jpgpu_batch* b, int which, T** out, size_t count) {
    jpgpu_ctx* ctx = b->ctx;
    jpgpu_batch::Arena& a = b->arena[which];
    const size_t need = std::max<size_t>(count * sizeof(T), 256);
    if (a.cap < need) {
        if (a.p) {
            CK(cudaStreamSynchronize(ctx->stream));
            CK(cudaFree(a.p));
            a.p = nullptr;
            a.cap = 0;
        }
        CK(cudaMalloc(&a.p, need));
        a.cap = need;
    }
    *out = reinterpret_cast<T*>(a.p);
    return JPGPU_OK;
}

// bytes per output sample: 1 (u8 formats) or 4 (JPGPU_OUT_F32_PLANAR); every size and offset of the output scales with it
size_t sample_bytes(const jpgpu_batch* b) { return b->dev.out_planar == 2u ? 4u : 1u; }

}  // namespace

extern "C" int jpgpu_create(int device, jpgpu_ctx** out) try {
    if (!out) return JPGPU_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return JPGPU_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return JPGPU_ERR_NO_DEVICE;
    if (prop.major != 10) return JPGPU_ERR_NO_DEVICE;  // kernels are built for sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return JPGPU_ERR_NO_DEVICE;
    jpgpu_ctx* c = new jpgpu_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return JPGPU_ERR_CUDA; }
    c->own_stream = true;
    if (init_constants() != cudaSuccess) { cudaStreamDestroy(c->stream); delete c; return JPGPU_ERR_CUDA; }
    *out = c;
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" void jpgpu_destroy(jpgpu_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->single) { jpgpu_batch_destroy(c->single); c->single = nullptr; }
    for (int i = 0; i < jpgpu_ctx::kAux; i++) {
        if (c->aux[i]) { cudaStreamSynchronize(c->aux[i]); cudaStreamDestroy(c->aux[i]); }
        if (c->join[i]) cudaEventDestroy(c->join[i]);
    }
    if (c->fork) cudaEventDestroy(c->fork);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" const char* jpgpu_last_error(const jpgpu_ctx* c) { return c ? c->err.c_str() : "null context"; }

extern "C" int jpgpu_set_stream(jpgpu_ctx* c, void* s) try {
    if (!c) return JPGPU_ERR_INVALID_ARG;
    cudaSetDevice(c->device);
    if (c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    c->stream = reinterpret_cast<cudaStream_t>(s);
    c->own_stream = false;
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" void* jpgpu_stream(const jpgpu_ctx* c) { return c ? reinterpret_cast<void*>(c->stream) : nullptr; }

extern "C" int jpgpu_sync(jpgpu_ctx* ctx) try {
    if (!ctx) return JPGPU_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" void jpgpu_batch_destroy(jpgpu_batch* b);

extern "C" void jpgpu_batch_destroy(jpgpu_batch* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    if (b->graph_exec) cudaGraphExecDestroy(b->graph_exec);
    for (auto& a : b->arena) if (a.p) cudaFree(a.p);
    delete b;
}

extern "C" int jpgpu_batch_replan(jpgpu_batch* b, const jpgpu_image_desc* descs, size_t n) try {
    if (!b || (!descs && n)) return JPGPU_ERR_INVALID_ARG;
    jpgpu_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    // the previous plan's host vectors may still be the source of an asynchronous upload
    CK(cudaStreamSynchronize(ctx->stream));
    b->n = n;
    int st = build_plan(descs, n, b->plan);
    if (st != JPGPU_OK) return st;
    b->host_scan.resize(n);
    b->host_scan_len.resize(n);
    for (size_t i = 0; i < n; i++) { b->host_scan[i] = descs[i].scan; b->host_scan_len[i] = descs[i].scan_len; }
    HostPlan& p = b->plan;
    const bool external_rgb = b->dev.rgb && b->dev.rgb != b->own_rgb;
    uint8_t* const ext_rgb = b->dev.rgb;
    const BatchDev prev = b->dev;   // the output format outlives a replan
    memset(&b->dev, 0, sizeof b->dev);
    BatchDev& d = b->dev;
    d.out_planar = prev.out_planar;
    for (int k = 0; k < 3; k++) { d.out_scale[k] = prev.out_scale[k]; d.out_bias[k] = prev.out_bias[k]; }
    d.n_images = (uint32_t)n;
    d.n_seqs = (uint32_t)p.seqs.size();
    d.nsync = p.nsync;
    d.sub_bits = p.sub_bits;
    d.lw = p.lw;
    d.lookback_bits = p.lookback_bits;
    d.max_slots = p.max_slots;
    d.seg_bits = p.seg_bits;
    d.wp_shift = p.wp_shift;
#define TRY(x) do { st = (x); if (st != JPGPU_OK) return st; } while (0)
    {
        // every table of the plan goes to the device as ONE copy: the sections are laid out in a host blob (256-byte
        // aligned), copied once, and the device pointers are the blob's base plus the section offsets.  (Eleven separate
        // copies from pageable memory were ~60 us of a single-image call.)
        std::vector<uint8_t>& blob = b->plan_blob;
        blob.clear();
        auto add = [&blob](const void* src, size_t bytes) {
            const size_t off = (blob.size() + 255) & ~(size_t)255;
            blob.resize(off + bytes);
            if (bytes) memcpy(blob.data() + off, src, bytes);
            return off;
        };
        const size_t o_imgs = add(p.imgs.data(), p.imgs.size() * sizeof(ImgDev));
        const size_t o_seqs = add(p.seqs.data(), p.seqs.size() * sizeof(SeqDesc));
        const size_t o_luts = add(p.luts.data(), p.luts.size() * sizeof(HuffLut));
        const size_t o_mlut = add(p.mluts.data(), p.mluts.size() * sizeof(uint32_t));
        const size_t o_moff = add(p.mlut_off.data(), p.mlut_off.size() * sizeof(uint32_t));
        const size_t o_qt = add(p.qt.data(), p.qt.size() * sizeof(float));
        size_t o_kind[kNumKinds];
        for (int k = 0; k < kNumKinds; k++) o_kind[k] = add(p.kind_imgs[k].data(), p.kind_imgs[k].size() * sizeof(uint32_t));
        const size_t o_gmap = add(p.gmap.data(), p.gmap.size() * sizeof(uint32_t));
        const size_t o_frames = add(p.frames.data(), p.frames.size() * sizeof(FrameDev));
        uint8_t* base = nullptr;
        TRY(dev_ensure(b, jpgpu_batch::kPlanBlob, &base, blob.size() + 256));
        if (!blob.empty()) CK(cudaMemcpyAsync(base, blob.data(), blob.size(), cudaMemcpyHostToDevice, ctx->stream));
        d.imgs = reinterpret_cast<const ImgDev*>(base + o_imgs);
        d.seqs = reinterpret_cast<const SeqDesc*>(base + o_seqs);
        d.luts = reinterpret_cast<const HuffLut*>(base + o_luts);
        d.mlut = reinterpret_cast<const uint32_t*>(base + o_mlut);
        d.mlut_off = reinterpret_cast<const uint32_t*>(base + o_moff);
        d.qt = reinterpret_cast<const float*>(base + o_qt);
        for (int k = 0; k < kNumKinds; k++) d.kind_imgs[k] = reinterpret_cast<const uint32_t*>(base + o_kind[k]);
        d.gmap = reinterpret_cast<const uint32_t*>(base + o_gmap);
        d.frames = reinterpret_cast<const FrameDev*>(base + o_frames);
        d.n_frames = (uint32_t)p.frames.size();
        d.frame_max_quads = p.frame_max_quads;
    }
    d.max_mlut_words = p.max_mlut_words;
    {
        const char* e = getenv("JPGPU_SYNC_MULTI");   // experiments: 0 = single-symbol synchronisation pass
        d.sync_multi = e ? (atoi(e) ? 1u : 0u) : 1u;
        // The repair walks are latency-bound, one CTA per image: with the multi-symbol tables in shared memory three CTAs
        // fit an SM instead of seven, which pays while every image's CTA is resident at once (256 x 1080p: 0.40 -> 0.26 ms)
        // and costs beyond (1024 images: 0.59 -> 0.64 ms).
        const char* v = getenv("JPGPU_VERIFY_MULTI");
        d.verify_multi = d.sync_multi && (v ? atoi(v) != 0 : n <= 3 * 148) ? 1u : 0u;
    }
    for (int k = 0; k < kNumKinds; k++) {
        d.kind_count[k] = (uint32_t)p.kind_imgs[k].size();
        d.kind_max_tiles[k] = p.kind_max_tiles[k];
    }
    d.gather_max_blocks = p.gather_max_blocks;
    d.gather_max_quads = p.gather_max_quads;
    if (p.sample_floats) TRY(dev_ensure(b, jpgpu_batch::kSamples, &d.samples, p.sample_floats));
    uint8_t* raw = nullptr;
    TRY(dev_ensure(b, jpgpu_batch::kRaw, &raw, p.raw_bytes + 64));
    d.raw = raw;
    TRY(dev_ensure(b, jpgpu_batch::kDyn, &d.dyn, n + 1));
    // slack: the decoders' window prefetch may read one warp group past an image's last word (see fast_advance)
    const size_t stream_alloc_words = p.stream_words + 64 + ((size_t)32 << p.lw);
    TRY(dev_ensure(b, jpgpu_batch::kStream, &d.stream, stream_alloc_words));
    TRY(dev_ensure(b, jpgpu_batch::kSegtab, &d.segtab, p.seg_entries + 8));
    TRY(dev_ensure(b, jpgpu_batch::kSubs, &d.subs, p.sub_entries + 1));
    TRY(dev_ensure(b, jpgpu_batch::kSegs, &d.segs, p.sub_entries * (p.sub_bits / p.seg_bits) + 1));
    TRY(dev_ensure(b, jpgpu_batch::kChunks, &d.chunk_counts, p.chunk_entries + 1));
    d.max_chunks = p.max_chunks;
    TRY(dev_ensure(b, jpgpu_batch::kCoefs, &d.coefs, p.coef_elems + 64));
    if (external_rgb && b->ext_rgb_cap >= p.rgb_bytes * sample_bytes(b)) {
        d.rgb = ext_rgb;   // stays where jpgpu_batch_set_device_output() pointed it: the new plan fits the caller's arena
    } else {
        // no external arena, or one too small for the new plan: the batch's own arena (grown to the plan)
        b->ext_rgb_cap = 0;
        TRY(dev_ensure(b, jpgpu_batch::kRgb, &d.rgb, p.rgb_bytes * sample_bytes(b) + 256));
    }
    b->own_rgb = static_cast<uint8_t*>(b->arena[jpgpu_batch::kRgb].p);
#undef TRY
    b->coef_bytes = p.coef_elems * sizeof(int16_t);
    b->dev_gen++;
    // defined contents for everything a speculative decoder may read
    CK(cudaMemsetAsync(raw, 0, p.raw_bytes + 64, ctx->stream));
    CK(cudaMemsetAsync(d.stream, 0, stream_alloc_words * 4, ctx->stream));
    CK(cudaMemsetAsync(d.dyn, 0, (n + 1) * sizeof(ImgDyn), ctx->stream));
    // (the coefficient arena is not cleared: the write pass stores every block of a complete scan, and the blocks a
    // damaged scan never reaches are zero-filled from ImgDyn::coef_end on by zero_tail_kernel)
    CK(cudaMemsetAsync(d.segtab, 0, (p.seg_entries + 8) * 4, ctx->stream));
    // (a caller-owned arena kept across the replan still holds the previous wave's pixels at the previous plan's offsets:
    // its gaps are cleared when jpgpu_batch_set_device_output() points the new plan at its slice)
    if (d.rgb == b->own_rgb) { launch_zero_output_pads(d, ctx->stream); b->launches += 1; }
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" int jpgpu_batch_create(jpgpu_ctx* ctx, const jpgpu_image_desc* descs, size_t n, jpgpu_batch** out) try {
    if (!ctx || !out || (!descs && n)) return JPGPU_ERR_INVALID_ARG;
    *out = nullptr;
    jpgpu_batch* b = new jpgpu_batch();
    b->ctx = ctx;
    memset(&b->dev, 0, sizeof b->dev);
    const int st = jpgpu_batch_replan(b, descs, n);
    if (st != JPGPU_OK) { jpgpu_batch_destroy(b); return st; }
    *out = b;
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" int jpgpu_batch_upload(jpgpu_batch* b) try {
    if (!b) return JPGPU_ERR_INVALID_ARG;
    jpgpu_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    uint8_t* raw = const_cast<uint8_t*>(b->dev.raw);
    for (size_t i = 0; i < b->n; i++) {
        if (b->plan.status[i] != JPGPU_OK) continue;
        CK(cudaMemcpyAsync(raw + b->plan.imgs[i].raw_off, b->host_scan[i], b->plan.imgs[i].raw_len,
                           cudaMemcpyHostToDevice, ctx->stream));
    }
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" int jpgpu_batch_set_device_scans(jpgpu_batch* b, const void* dev_base, const uint64_t* offsets) try {
    if (!b || !dev_base || !offsets) return JPGPU_ERR_INVALID_ARG;
    jpgpu_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    if (!b->n) return JPGPU_OK;
    // one launch for the whole batch (a copy per image would serialise ~4 us each on the stream)
    uint64_t* dev_offs = nullptr;
    int st = dev_ensure(b, jpgpu_batch::kScanOffs, &dev_offs, b->n);
    if (st != JPGPU_OK) return st;
    CK(cudaMemcpyAsync(dev_offs, offsets, b->n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    launch_gather_scans(b->dev, dev_base, dev_offs, ctx->stream);
    CK(cudaGetLastError());
    b->launches += 1;
    return JPGPU_OK;
} JPGPU_CATCH_ALL

// One host->device copy for the whole batch: all scans lie in one readable host range, so the used
// span goes over PCIe as ONE transfer into a staging arena and one kernel spreads it into the raw arena (per-image
// copies cost ~2 us of launch each and 1024 of them keep the copy engine far from its rate).
extern "C" int jpgpu_batch_upload_from(jpgpu_batch* b, const void* host_base, size_t host_bytes) try {
    if (!b || !host_base) return JPGPU_ERR_INVALID_ARG;
    jpgpu_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    if (!b->n) return JPGPU_OK;
    const uint8_t* base = static_cast<const uint8_t*>(host_base);
    size_t lo = SIZE_MAX, hi = 0;
    for (size_t i = 0; i < b->n; i++) {
        if (b->plan.status[i] != JPGPU_OK) continue;
        if (b->host_scan[i] < base) return JPGPU_ERR_INVALID_ARG;
        const size_t o = (size_t)(b->host_scan[i] - base);
        if (o > host_bytes || b->host_scan_len[i] > host_bytes - o) return JPGPU_ERR_INVALID_ARG;
        lo = std::min(lo, o);
        hi = std::max(hi, o + b->host_scan_len[i]);
    }
    if (lo >= hi) return JPGPU_OK;   // no decodable image
    lo &= ~(size_t)15;                // keeps every scan's alignment relative to the staging arena
    uint8_t* stage = nullptr;
    int st = dev_ensure(b, jpgpu_batch::kStage, &stage, hi - lo + 64);
    if (st != JPGPU_OK) return st;
    uint64_t* dev_offs = nullptr;
    st = dev_ensure(b, jpgpu_batch::kScanOffs, &dev_offs, b->n);
    if (st != JPGPU_OK) return st;
    b->scan_offs.resize(b->n);
    for (size_t i = 0; i < b->n; i++)
        b->scan_offs[i] = b->plan.status[i] == JPGPU_OK ? (uint64_t)(b->host_scan[i] - base) - lo : 0;
    CK(cudaMemcpyAsync(stage, base + lo, hi - lo, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dev_offs, b->scan_offs.data(), b->n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    launch_gather_scans(b->dev, stage, dev_offs, ctx->stream);
    CK(cudaGetLastError());
    b->launches += 1;
    return JPGPU_OK;
} JPGPU_CATCH_ALL

// One device->host copy of the whole output arena: image i lands at host_base + jpgpu_batch_rgb_offset(i), the
// layout the device arena has (256-byte aligned slices).
extern "C" int jpgpu_batch_download_contiguous(jpgpu_batch* b, void* host_base, size_t capacity) try {
    if (!b || !host_base) return JPGPU_ERR_INVALID_ARG;
    const size_t total = b->plan.rgb_bytes * sample_bytes(b);
    if (capacity < total) return JPGPU_ERR_INVALID_ARG;
    jpgpu_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    if (total) CK(cudaMemcpyAsync(host_base, b->dev.rgb, total, cudaMemcpyDeviceToHost, ctx->stream));
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" int jpgpu_batch_set_device_output(jpgpu_batch* b, void* dev_base, size_t capacity) try {
    if (!b) return JPGPU_ERR_INVALID_ARG;
    if (!dev_base) {
        jpgpu_ctx* ctx = b->ctx;
        CK(cudaSetDevice(ctx->device));
        // the own arena may never have been allocated, or only for an earlier, smaller plan (replans under an external
        // output skip it): size it for the current plan
        int st = dev_ensure(b, jpgpu_batch::kRgb, &b->own_rgb, b->plan.rgb_bytes * sample_bytes(b) + 256);
        if (st != JPGPU_OK) return st;
        b->dev.rgb = b->own_rgb;
        b->ext_rgb_cap = 0;
        b->dev_gen++;
        launch_zero_output_pads(b->dev, ctx->stream);
        return JPGPU_OK;
    }
    if (capacity < b->plan.rgb_bytes * sample_bytes(b) || (reinterpret_cast<uintptr_t>(dev_base) & 255u)) return JPGPU_ERR_INVALID_ARG;
    b->dev.rgb = static_cast<uint8_t*>(dev_base);
    b->ext_rgb_cap = capacity;
    b->dev_gen++;
    {
        jpgpu_ctx* ctx = b->ctx;
        CK(cudaSetDevice(ctx->device));
        launch_zero_output_pads(b->dev, ctx->stream);
    }
    return JPGPU_OK;
} JPGPU_CATCH_ALL

namespace {
// kernels one entropy chain (pre-pass .. write pass) launches for the images / jobs `d` covers
uint64_t entropy_launches(const BatchDev& d) {
    const uint64_t slices = (d.n_images + 65534u) / 65535u;   // kernels indexed by image in grid.y go in slices
    uint64_t n = 0;
    if (d.n_images && d.max_chunks) n += 2 * slices + 1;      // prepass count / scan / write
    if (d.n_seqs && d.nsync) n += 1;                          // sync
    if (d.n_images && d.nsync) n += 1;                        // verify + scan
    if (d.n_seqs) n += 1 + (d.n_images ? 1 : 0);              // write pass, zero tail
    return n;
}
}  // namespace

extern "C" int jpgpu_batch_entropy(jpgpu_batch* b) try {
    if (!b) return JPGPU_ERR_INVALID_ARG;
    jpgpu_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    launch_prepass(b->dev, s);
    CK(launch_sync(b->dev, s));
    CK(launch_verify_scan(b->dev, s));
    CK(launch_decode_write(b->dev, s));
    b->launches += entropy_launches(b->dev);
    CK(cudaGetLastError());
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" int jpgpu_batch_idct(jpgpu_batch* b) try {
    if (!b) return JPGPU_ERR_INVALID_ARG;
    jpgpu_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    b->launches += (uint64_t)launch_idct_colour(b->dev, ctx->stream);
    CK(cudaGetLastError());
    b->decoded = true;
    return JPGPU_OK;
} JPGPU_CATCH_ALL

namespace {

BatchDev group_dev(const jpgpu_batch* b, const GroupPlan& g) {
    BatchDev d = b->dev;
    d.img0 = g.img0; d.n_images = g.nimg;
    d.job0 = g.job0; d.n_seqs = g.njobs;
    d.nsync = g.nsync;
    d.max_chunks = g.max_chunks;
    for (int k = 0; k < kNumKinds; k++) {
        d.kind_imgs[k] = b->dev.kind_imgs[k] + g.kind_lo[k];
        d.kind_count[k] = g.kind_hi[k] - g.kind_lo[k];
        d.kind_max_tiles[k] = g.kind_max_tiles[k];
    }
    d.gather_max_blocks = g.gather_max_blocks;
    d.gather_max_quads = g.gather_max_quads;
    d.frames = b->dev.frames + g.frame_lo;
    d.n_frames = g.frame_hi - g.frame_lo;
    d.frame_max_quads = g.frame_max_quads;
    return d;
}

int ensure_aux(jpgpu_ctx* ctx) {
    for (int i = 0; i < jpgpu_ctx::kAux; i++) {
        if (!ctx->aux[i]) CK(cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking));
        if (!ctx->join[i]) CK(cudaEventCreateWithFlags(&ctx->join[i], cudaEventDisableTiming));
    }
    if (!ctx->fork) CK(cudaEventCreateWithFlags(&ctx->fork, cudaEventDisableTiming));
    return JPGPU_OK;
}

}  // namespace

// entropy + idct.  A large batch is cut into groups of images (HostPlan::groups) whose kernel chains alternate
// between three auxiliary streams: the low-occupancy ends of one group's kernels (repair walks, last waves) overlap
// the next group's work.  Forked from and joined back into the context stream.
static int enqueue_decode(jpgpu_batch* b);

extern "C" int jpgpu_batch_decode(jpgpu_batch* b) try {
    if (!b) return JPGPU_ERR_INVALID_ARG;
    jpgpu_ctx* ctx = b->ctx;
    const char* genv = getenv("JPGPU_GRAPH");   // experiments / tests: 0 = always kernel by kernel
    const bool graphs_on = !genv || atoi(genv) != 0;
    if (!graphs_on || !b->graph_ok || !ctx->stream || b->seen_gen != b->dev_gen) {
        // first decode with these launch parameters (or graphs unavailable): enqueue kernel by kernel.  A batch decoded
        // once - jpgpu_decode, a wave of a long job - never pays for a capture.
        b->seen_gen = b->dev_gen;
        return enqueue_decode(b);
    }
    CK(cudaSetDevice(ctx->device));
    if (!b->graph_exec || b->graph_gen != b->dev_gen) {
        if (b->graph_exec) { cudaGraphExecDestroy(b->graph_exec); b->graph_exec = nullptr; }
        cudaGraph_t graph = nullptr;
        const uint64_t before = b->launches;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            b->graph_ok = false;
            return enqueue_decode(b);
        }
        const int st = enqueue_decode(b);
        const cudaError_t ec = cudaStreamEndCapture(ctx->stream, &graph);
        b->graph_launches = b->launches - before;
        b->launches = before;
        if (st != JPGPU_OK || ec != cudaSuccess || !graph || cudaGraphInstantiate(&b->graph_exec, graph, 0) != cudaSuccess) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            b->graph_exec = nullptr;
            b->graph_ok = false;          // this batch goes on kernel by kernel
            return st != JPGPU_OK ? st : enqueue_decode(b);
        }
        cudaGraphDestroy(graph);
        b->graph_gen = b->dev_gen;
    }
    CK(cudaGraphLaunch(b->graph_exec, ctx->stream));
    b->launches += b->graph_launches;
    b->decoded = true;
    return JPGPU_OK;
} JPGPU_CATCH_ALL

static int enqueue_decode(jpgpu_batch* b) {
    jpgpu_ctx* ctx = b->ctx;
    if (b->plan.groups.size() <= 1) {
        int st = jpgpu_batch_entropy(b);
        if (st != JPGPU_OK) return st;
        return jpgpu_batch_idct(b);
    }
    CK(cudaSetDevice(ctx->device));
    int st = ensure_aux(ctx);
    if (st != JPGPU_OK) return st;
    CK(cudaEventRecord(ctx->fork, ctx->stream));
    for (int i = 0; i < jpgpu_ctx::kAux; i++) CK(cudaStreamWaitEvent(ctx->aux[i], ctx->fork, 0));
    for (size_t g = 0; g < b->plan.groups.size(); g++) {
        cudaStream_t s = ctx->aux[g % jpgpu_ctx::kAux];
        const BatchDev d = group_dev(b, b->plan.groups[g]);
        launch_prepass(d, s);
        CK(launch_sync(d, s));
        CK(launch_verify_scan(d, s));
        CK(launch_decode_write(d, s));
        b->launches += entropy_launches(d) + (uint64_t)launch_idct_colour(d, s);
    }
    for (int i = 0; i < jpgpu_ctx::kAux; i++) {
        CK(cudaEventRecord(ctx->join[i], ctx->aux[i]));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->join[i], 0));
    }
    CK(cudaGetLastError());
    b->decoded = true;
    return JPGPU_OK;
}

extern "C" int jpgpu_batch_download(jpgpu_batch* b, uint8_t* const* outs) try {
    if (!b || !outs) return JPGPU_ERR_INVALID_ARG;
    jpgpu_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    for (size_t i = 0; i < b->n; i++) {
        if (b->plan.status[i] != JPGPU_OK || !outs[i]) continue;
        const ImgDev& im = b->plan.imgs[i];
        const size_t sb = sample_bytes(b), nbytes = (size_t)b->plan.out_w[i] * b->plan.out_h[i] * 3 * sb;
        if (!nbytes) continue;   // a further scan of a frame: the pixels belong to the frame's first scan
        CK(cudaMemcpyAsync(outs[i], b->dev.rgb + im.rgb_off * sb, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" int jpgpu_batch_set_output_format(jpgpu_batch* b, uint32_t format) try {
    if (!b || format > JPGPU_OUT_F32_PLANAR) return JPGPU_ERR_INVALID_ARG;
    jpgpu_ctx* ctx = b->ctx;
    const size_t before = sample_bytes(b);
    b->dev.out_planar = format;
    b->dev_gen++;
    if (format == JPGPU_OUT_F32_PLANAR && b->dev.out_scale[0] == 0.0f && b->dev.out_scale[1] == 0.0f && b->dev.out_scale[2] == 0.0f)
        for (int k = 0; k < 3; k++) { b->dev.out_scale[k] = 1.0f / 255.0f; b->dev.out_bias[k] = 0.0f; }   // default: [0, 1]
    if (sample_bytes(b) != before) {
        // the output grows or shrinks four-fold: a caller-owned arena set for the other format no longer applies
        CK(cudaSetDevice(ctx->device));
        const int st = dev_ensure(b, jpgpu_batch::kRgb, &b->own_rgb, b->plan.rgb_bytes * sample_bytes(b) + 256);
        if (st != JPGPU_OK) return st;
        b->dev.rgb = b->own_rgb;
        b->ext_rgb_cap = 0;
        launch_zero_output_pads(b->dev, ctx->stream);
    }
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" int jpgpu_batch_set_normalisation(jpgpu_batch* b, const float scale[3], const float bias[3]) try {
    if (!b || !scale || !bias) return JPGPU_ERR_INVALID_ARG;
    for (int k = 0; k < 3; k++) { b->dev.out_scale[k] = scale[k]; b->dev.out_bias[k] = bias[k]; }
    b->dev_gen++;
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" void* jpgpu_batch_device_rgb(jpgpu_batch* b, size_t i, size_t* nbytes) {
    if (!b || i >= b->n || b->plan.status[i] != JPGPU_OK) return nullptr;
    const ImgDev& im = b->plan.imgs[i];
    if (nbytes) *nbytes = (size_t)b->plan.out_w[i] * b->plan.out_h[i] * 3 * sample_bytes(b);
    return b->dev.rgb + im.rgb_off * sample_bytes(b);
}

extern "C" size_t jpgpu_batch_output_bytes(const jpgpu_batch* b) { return b ? (size_t)b->plan.rgb_bytes * sample_bytes(b) : 0; }

extern "C" int jpgpu_batch_rgb_offset(const jpgpu_batch* b, size_t i, size_t* offset, size_t* nbytes) {
    if (!b || i >= b->n) return JPGPU_ERR_INVALID_ARG;
    const ImgDev& im = b->plan.imgs[i];
    if (offset) *offset = (size_t)im.rgb_off * sample_bytes(b);
    if (nbytes) *nbytes = b->plan.status[i] == JPGPU_OK ? (size_t)b->plan.out_w[i] * b->plan.out_h[i] * 3 * sample_bytes(b) : 0;
    return JPGPU_OK;
}

extern "C" int jpgpu_batch_results(jpgpu_batch* b, int32_t* statuses, uint64_t* bytes_read) try {
    if (!b) return JPGPU_ERR_INVALID_ARG;
    jpgpu_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    std::vector<ImgDyn> dyn(b->n);
    if (b->n) CK(cudaMemcpyAsync(dyn.data(), b->dev.dyn, b->n * sizeof(ImgDyn), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < b->n; i++) {
        int32_t st = b->plan.status[i];
        uint64_t br = 0;
        if (st == JPGPU_OK) {
            const uint32_t f = dyn[i].status;
            if (f & kStDcSize) st = JPGPU_PANIC_READ_BITS_ASSERT;
            else if (f & kStBadCode) st = JPGPU_ERR_BAD_CODE;
            else if (f & kStRestart) st = JPGPU_ERR_RESTART;
            else if (!(f & kStDone)) st = JPGPU_ERR_TRUNCATED;
            br = ((uint64_t)dyn[i].bits_consumed + 7) / 8;  // decoder.rs:336-340
        }
        if (statuses) statuses[i] = st;
        if (bytes_read) bytes_read[i] = br;
    }
    // a frame of several scans is as good as its worst scan: the first scan (which owns the pixels) reports it
    if (statuses)
        for (size_t i = 0; i < b->n; i++) {
            if (b->plan.frame_part[i] != 1) continue;
            for (size_t k = i + 1; k < b->n && b->plan.frame_part[k] == 2; k++)
                if (statuses[i] == JPGPU_OK && statuses[k] != JPGPU_OK) statuses[i] = statuses[k];
        }
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" int jpgpu_batch_coefficients(jpgpu_batch* b, size_t i, int16_t* out, size_t cap, uint32_t nblocks[4]) try {
    if (!b || i >= b->n || !out || !nblocks) return JPGPU_ERR_INVALID_ARG;
    if (b->plan.status[i] != JPGPU_OK) return b->plan.status[i];
    jpgpu_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    const ImgDev& im = b->plan.imgs[i];
    if (cap < im.total_coefs) return JPGPU_ERR_INVALID_ARG;
    std::vector<int16_t> arena(im.total_coefs);
    CK(cudaMemcpyAsync(arena.data(), b->dev.coefs + im.coef_off, (size_t)im.total_coefs * 2, cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    export_reference_order(im, arena.data(), out, nblocks);
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" int jpgpu_batch_stats(jpgpu_batch* b, uint64_t stats[8]) try {
    if (!b || !stats) return JPGPU_ERR_INVALID_ARG;
    const HostPlan& p = b->plan;
    stats[0] = p.tot_scan_bytes;
    stats[1] = p.tot_blocks * 128;
    stats[2] = p.tot_rgb_bytes;
    stats[3] = p.tot_pixels;
    stats[4] = p.tot_blocks;
    stats[5] = p.seqs.size();
    stats[6] = p.sub_entries;
    stats[7] = p.raw_bytes + p.stream_words * 4 + p.coef_elems * 2 + p.rgb_bytes + p.sub_entries * sizeof(SubInfo);
    stats[2] = p.tot_rgb_bytes;
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" int jpgpu_batch_profile(jpgpu_batch* b, float ms[8]) try {
    if (!b || !ms) return JPGPU_ERR_INVALID_ARG;
    jpgpu_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    cudaEvent_t ev[8];
    for (auto& e : ev) CK(cudaEventCreate(&e));
    CK(cudaEventRecord(ev[0], s));
    for (int step = 0; step < 3; step++) { launch_prepass_step(b->dev, s, step); CK(cudaEventRecord(ev[1 + step], s)); }
    CK(launch_sync(b->dev, s));
    CK(cudaEventRecord(ev[4], s));
    CK(launch_verify_scan(b->dev, s));
    CK(cudaEventRecord(ev[5], s));
    CK(launch_decode_write(b->dev, s));
    CK(cudaEventRecord(ev[6], s));
    b->launches += entropy_launches(b->dev) + (uint64_t)launch_idct_colour(b->dev, s);
    CK(cudaEventRecord(ev[7], s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    for (int i = 0; i < 7; i++) CK(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
    ms[7] = 0.0f;
    for (auto& e : ev) cudaEventDestroy(e);
    b->decoded = true;
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" uint64_t jpgpu_batch_launch_count(const jpgpu_batch* b) { return b ? b->launches : 0; }

extern "C" int jpgpu_decode(jpgpu_ctx* ctx, const jpgpu_image_desc* desc, uint8_t* rgb_out, size_t* bytes_read) try {
    if (!ctx || !desc || !rgb_out) return JPGPU_ERR_INVALID_ARG;
    int st;
    if (!ctx->single) st = jpgpu_batch_create(ctx, desc, 1, &ctx->single);
    else st = jpgpu_batch_replan(ctx->single, desc, 1);
    if (st != JPGPU_OK) return st;
    jpgpu_batch* b = ctx->single;
    int32_t ist = JPGPU_OK;
    uint64_t br = 0;
    uint8_t* outs[1] = {rgb_out};
    st = jpgpu_batch_upload(b);
    if (st == JPGPU_OK) st = jpgpu_batch_decode(b);
    if (st == JPGPU_OK) st = jpgpu_batch_download(b, outs);
    if (st == JPGPU_OK) st = jpgpu_batch_results(b, &ist, &br);
    if (st != JPGPU_OK) return st;
    if (bytes_read) *bytes_read = (size_t)br;
    return ist;
} JPGPU_CATCH_ALL

extern "C" int jpgpu_decode_file(jpgpu_ctx* ctx, const uint8_t* file, size_t len, uint32_t ext_flags, uint32_t layout,
                                 uint8_t* rgb_out, size_t rgb_cap, uint32_t* width, uint32_t* height, size_t* bytes_read) try {
    if (!ctx || !file || !rgb_out) return JPGPU_ERR_INVALID_ARG;
    if (ext_flags & JPGPU_EXT_MULTISCAN) {
        // a file that may hold one non-interleaved scan per component: every scan is a descriptor, the frame's pixels
        // are the output of the first
        jpgpu_image_desc ds[4];
        size_t n = 0;
        int st = jpgpu_parse_scans(file, len, ext_flags, layout, ds, 4, &n);
        if (st != JPGPU_OK) return st;
        const uint32_t w = ds[0].frame_part ? ds[0].frame_width : ds[0].width, h = ds[0].frame_part ? ds[0].frame_height : ds[0].height;
        if (width) *width = w;
        if (height) *height = h;
        if ((size_t)w * h * 3 > rgb_cap) return JPGPU_ERR_INVALID_ARG;
        if (!ctx->single) st = jpgpu_batch_create(ctx, ds, n, &ctx->single);
        else st = jpgpu_batch_replan(ctx->single, ds, n);
        if (st != JPGPU_OK) return st;
        jpgpu_batch* b = ctx->single;
        int32_t ist[4] = {0, 0, 0, 0};
        uint64_t br[4] = {0, 0, 0, 0};
        uint8_t* outs[4] = {rgb_out, nullptr, nullptr, nullptr};
        st = jpgpu_batch_upload(b);
        if (st == JPGPU_OK) st = jpgpu_batch_decode(b);
        if (st == JPGPU_OK) st = jpgpu_batch_download(b, outs);
        if (st == JPGPU_OK) st = jpgpu_batch_results(b, ist, br);
        if (st != JPGPU_OK) return st;
        if (bytes_read) { *bytes_read = 0; for (size_t k = 0; k < n; k++) *bytes_read += (size_t)br[k]; }
        return ist[0];
    }
    jpgpu_image_desc d;
    int st = jpgpu_parse(file, len, ext_flags, layout, &d);
    if (st != JPGPU_OK) return st;
    if (width) *width = d.width;
    if (height) *height = d.height;
    if ((size_t)d.width * d.height * 3 > rgb_cap) return JPGPU_ERR_INVALID_ARG;
    return jpgpu_decode(ctx, &d, rgb_out, bytes_read);
} JPGPU_CATCH_ALL


// ====================================================================================================================
// Host buffers in, host buffers out (the reference-facing call for many files: what mod.rs:415 would
// call per image, for a whole directory at once).  The images are cut into chunks that alternate between two stream
// sets: the upload of one chunk, the kernels of another and the download of a third overlap, every transfer is one
// cudaMemcpyAsync.  Plans and device arenas are made once at creation; a run only enqueues work.
struct jpgpu_pipeline {
    int device = 0;
    jpgpu_ctx* ctx[2] = {nullptr, nullptr};
    std::vector<jpgpu_batch*> chunks;
    std::vector<size_t> first;        // first image of every chunk (+ n at the end)
    std::vector<size_t> out_off;      // where every chunk's output arena lies in the caller's host buffer (+ total)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_mid = nullptr;
    bool ran = false;
};

extern "C" void jpgpu_pipeline_destroy(jpgpu_pipeline* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (jpgpu_batch* b : p->chunks) jpgpu_batch_destroy(b);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->ev_mid) cudaEventDestroy(p->ev_mid);
    for (jpgpu_ctx* c : p->ctx) jpgpu_destroy(c);
    delete p;
}

extern "C" int jpgpu_pipeline_create(int device, const jpgpu_image_desc* descs, size_t n, size_t chunk_images, jpgpu_pipeline** out) try {
    if (!out || (!descs && n)) return JPGPU_ERR_INVALID_ARG;
    *out = nullptr;
    if (chunk_images == 0) chunk_images = 64;
    jpgpu_pipeline* p = new jpgpu_pipeline();
    p->device = device;
    int st = JPGPU_OK;
    for (int k = 0; k < 2 && st == JPGPU_OK; k++) st = jpgpu_create(device, &p->ctx[k]);
    size_t off = 0;
    for (size_t i0 = 0, m = 0; i0 < n && st == JPGPU_OK; i0 += m) {
        m = std::min(chunk_images, n - i0);
        while (i0 + m < n && descs[i0 + m].frame_part == 2) m++;   // the scans of one frame stay in one chunk
        jpgpu_batch* b = nullptr;
        st = jpgpu_batch_create(p->ctx[p->chunks.size() & 1], descs + i0, m, &b);
        if (st != JPGPU_OK) break;
        p->first.push_back(i0);
        p->out_off.push_back(off);
        off += (jpgpu_batch_output_bytes(b) + 255) & ~(size_t)255;
        p->chunks.push_back(b);
    }
    p->first.push_back(n);
    p->out_off.push_back(off);
    if (st == JPGPU_OK && (cudaEventCreate(&p->ev0) != cudaSuccess || cudaEventCreate(&p->ev1) != cudaSuccess ||
                           cudaEventCreateWithFlags(&p->ev_mid, cudaEventDisableTiming) != cudaSuccess))
        st = JPGPU_ERR_CUDA;
    if (st != JPGPU_OK) { jpgpu_pipeline_destroy(p); return st; }
    *out = p;
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" size_t jpgpu_pipeline_output_bytes(const jpgpu_pipeline* p) { return p ? p->out_off.back() : 0; }

extern "C" int jpgpu_pipeline_image_offset(const jpgpu_pipeline* p, size_t i, size_t* offset, size_t* nbytes) {
    if (!p || i >= p->first.back()) return JPGPU_ERR_INVALID_ARG;
    const size_t c = (size_t)(std::upper_bound(p->first.begin(), p->first.end(), i) - p->first.begin()) - 1;
    size_t o = 0, nb = 0;
    const int st = jpgpu_batch_rgb_offset(p->chunks[c], i - p->first[c], &o, &nb);
    if (offset) *offset = p->out_off[c] + o;
    if (nbytes) *nbytes = nb;
    return st;
}

// Enqueues the whole job: per chunk one upload (all its scans lie in [host_in, host_in + host_in_bytes)), the decode,
// one download to host_out + its chunk offset.  Asynchronous; jpgpu_pipeline_sync() waits.
extern "C" int jpgpu_pipeline_run(jpgpu_pipeline* p, const void* host_in, size_t host_in_bytes, void* host_out, size_t host_out_bytes) try {
    if (!p || !host_in || !host_out || host_out_bytes < p->out_off.back()) return JPGPU_ERR_INVALID_ARG;
    jpgpu_ctx* ctx = p->ctx[0];
    CK(cudaSetDevice(p->device));
    // both stream sets start together and end together, so that ev0..ev1 brackets the job on the device
    CK(cudaEventRecord(p->ev0, p->ctx[0]->stream));
    CK(cudaStreamWaitEvent(p->ctx[1]->stream, p->ev0, 0));
    for (size_t c = 0; c < p->chunks.size(); c++) {
        jpgpu_batch* b = p->chunks[c];
        int st = jpgpu_batch_upload_from(b, host_in, host_in_bytes);
        if (st == JPGPU_OK) st = jpgpu_batch_decode(b);
        if (st == JPGPU_OK) st = jpgpu_batch_download_contiguous(b, static_cast<uint8_t*>(host_out) + p->out_off[c], p->out_off[c + 1] - p->out_off[c]);
        if (st != JPGPU_OK) { ctx->err = p->ctx[c & 1]->err; return st; }
    }
    CK(cudaEventRecord(p->ev_mid, p->ctx[1]->stream));
    CK(cudaStreamWaitEvent(p->ctx[0]->stream, p->ev_mid, 0));
    CK(cudaEventRecord(p->ev1, p->ctx[0]->stream));
    p->ran = true;
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" int jpgpu_pipeline_sync(jpgpu_pipeline* p) try {
    if (!p) return JPGPU_ERR_INVALID_ARG;
    jpgpu_ctx* ctx = p->ctx[0];
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->ctx[1]->stream));
    CK(cudaStreamSynchronize(p->ctx[0]->stream));
    return JPGPU_OK;
} JPGPU_CATCH_ALL

// Device time of the last run in milliseconds (first upload enqueued .. last download finished). Synchronises.
extern "C" int jpgpu_pipeline_elapsed_ms(jpgpu_pipeline* p, float* ms) try {
    if (!p || !ms || !p->ran) return JPGPU_ERR_INVALID_ARG;
    jpgpu_ctx* ctx = p->ctx[0];
    int st = jpgpu_pipeline_sync(p);
    if (st != JPGPU_OK) return st;
    CK(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" int jpgpu_pipeline_results(jpgpu_pipeline* p, int32_t* statuses, uint64_t* bytes_read) try {
    if (!p) return JPGPU_ERR_INVALID_ARG;
    for (size_t c = 0; c < p->chunks.size(); c++) {
        const int st = jpgpu_batch_results(p->chunks[c], statuses ? statuses + p->first[c] : nullptr, bytes_read ? bytes_read + p->first[c] : nullptr);
        if (st != JPGPU_OK) return st;
    }
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" uint64_t jpgpu_pipeline_launch_count(const jpgpu_pipeline* p) {
    uint64_t n = 0;
    if (p) for (const jpgpu_batch* b : p->chunks) n += b->launches;
    return n;
}

extern "C" const char* jpgpu_pipeline_last_error(const jpgpu_pipeline* p) { return p && p->ctx[0] ? p->ctx[0]->err.c_str() : "null pipeline"; }

// One call: host files in, host pixels out (the SURVEY 8(b) proposal's jpgpu_decode_batch with memory_kind HOST on one device).
extern "C" int jpgpu_decode_batch_host(int device, const jpgpu_image_desc* descs, size_t n, const void* host_in, size_t host_in_bytes,
                                       void* host_out, size_t host_out_bytes, size_t* out_offsets, int32_t* statuses, uint64_t* bytes_read) try {
    jpgpu_pipeline* p = nullptr;
    int st = jpgpu_pipeline_create(device, descs, n, 0, &p);
    if (st != JPGPU_OK) return st;
    st = jpgpu_pipeline_run(p, host_in, host_in_bytes, host_out, host_out_bytes);
    if (st == JPGPU_OK) st = jpgpu_pipeline_results(p, statuses, bytes_read);   // synchronises the chunk streams
    if (st == JPGPU_OK) st = jpgpu_pipeline_sync(p);
    if (st == JPGPU_OK && out_offsets)
        for (size_t i = 0; i < n; i++) jpgpu_pipeline_image_offset(p, i, &out_offsets[i], nullptr);
    jpgpu_pipeline_destroy(p);
    return st;
} JPGPU_CATCH_ALL
