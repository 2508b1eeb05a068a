// jpgpu_multi.cpp — one process, several GPUs (SURVEY.md 8(b) "jpgpu_create(device_ids*, n_devices)" / 8(e)).
//
// The reference's caller (JPEGImage::parse, mod.rs:202) is one process.  A batch of independent images is cut into
// contiguous ranges of about equal scan bytes, one range per device; every device has its own context, stream set,
// batch object and a worker thread that plans and enqueues its range, so that planning and kernel launches of the
// devices run side by side.  There is no communication between devices (no NCCL, no peer access): results stay where
// they were produced or go back to the caller's host buffers.  Built on the single-device C ABI only.
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/jpgpu.h"

namespace {

// A worker thread bound to one device: runs the jobs it is handed, one at a time, in order.
class Worker {
public:
    Worker() : th_([this] { loop(); }) {}
    ~Worker() {
        { std::lock_guard<std::mutex> l(mu_); quit_ = true; }
        cv_.notify_all();
        th_.join();
    }
    void post(std::function<int()> job) {
        { std::lock_guard<std::mutex> l(mu_); job_ = std::move(job); busy_ = true; }
        cv_.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [this] { return !busy_; });
        return status_;
    }

private:
    void loop() {
        std::unique_lock<std::mutex> l(mu_);
        for (;;) {
            cv_.wait(l, [this] { return quit_ || (busy_ && job_); });
            if (quit_) return;
            std::function<int()> job = std::move(job_);
            job_ = nullptr;
            l.unlock();
            int st;
            try { st = job(); } catch (const std::bad_alloc&) { st = JPGPU_ERR_OOM; } catch (...) { st = JPGPU_ERR_INVALID_ARG; }
            l.lock();
            status_ = st;
            busy_ = false;
            cv_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::function<int()> job_;
    bool busy_ = false, quit_ = false;
    int status_ = JPGPU_OK;
    std::thread th_;   // last: started when everything above exists
};

struct Dev {
    int device = 0;
    jpgpu_ctx* ctx = nullptr;
    jpgpu_batch* batch = nullptr;
    size_t first = 0, count = 0;   // its image range
    Worker* worker = nullptr;
};

}  // namespace

struct jpgpu_multi {
    std::vector<Dev> devs;
    size_t n = 0;
    std::atomic<int> gate{0};      // start line of jpgpu_multi_time_decode
    // runs fn(d) on every device's worker and returns the first status that is not JPGPU_OK
    int run_all(const std::function<int(Dev&)>& fn) {
        for (Dev& d : devs) d.worker->post([&fn, &d] { return fn(d); });
        int st = JPGPU_OK;
        for (Dev& d : devs) { const int s = d.worker->wait(); if (st == JPGPU_OK) st = s; }
        return st;
    }
    Dev* owner(size_t i) {
        for (Dev& d : devs) if (i >= d.first && i < d.first + d.count) return &d;
        return nullptr;
    }
};

#define JPGPU_CATCH_ALL                                       \
    catch (const std::bad_alloc&) { return JPGPU_ERR_OOM; }   \
    catch (...) { return JPGPU_ERR_INVALID_ARG; }

extern "C" void jpgpu_multi_destroy(jpgpu_multi* m) {
    if (!m) return;
    for (Dev& d : m->devs) {
        if (d.worker) {
            d.worker->post([&d] {
                if (d.batch) jpgpu_batch_destroy(d.batch);
                if (d.ctx) jpgpu_destroy(d.ctx);
                return (int)JPGPU_OK;
            });
            d.worker->wait();
            delete d.worker;
        }
    }
    delete m;
}

extern "C" int jpgpu_multi_create(const int* devices, int n_devices, jpgpu_multi** out) try {
    if (!out || !devices || n_devices <= 0 || n_devices > 64) return JPGPU_ERR_INVALID_ARG;
    *out = nullptr;
    for (int a = 0; a < n_devices; a++)
        for (int b = 0; b < a; b++) if (devices[a] == devices[b]) return JPGPU_ERR_INVALID_ARG;
    jpgpu_multi* m = new jpgpu_multi();
    m->devs.resize((size_t)n_devices);
    for (int k = 0; k < n_devices; k++) { m->devs[(size_t)k].device = devices[k]; m->devs[(size_t)k].worker = new Worker(); }
    const int st = m->run_all([](Dev& d) { return jpgpu_create(d.device, &d.ctx); });
    if (st != JPGPU_OK) { jpgpu_multi_destroy(m); return st; }
    *out = m;
    return JPGPU_OK;
} JPGPU_CATCH_ALL

extern "C" int jpgpu_multi_device_count(const jpgpu_multi* m) { return m ? (int)m->devs.size() : 0; }

// Host only: cuts n images into `parts` contiguous ranges of about equal scan bytes (the entropy stage, the longer
// one, scales with bytes, not pixels).  first[k] .. first[k+1] is range k; first has parts + 1 entries.
extern "C" int jpgpu_partition(const jpgpu_image_desc* descs, size_t n, size_t parts, size_t* first) {
    if ((!descs && n) || !parts || !first) return JPGPU_ERR_INVALID_ARG;
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) total += descs[i].scan_len;
    size_t i = 0;
    uint64_t acc = 0;
    for (size_t k = 0; k < parts; k++) {
        first[k] = i;
        const uint64_t goal = total / parts * (k + 1) + total % parts * (k + 1) / parts;
        // an image goes to the range in which its middle byte falls; the scans of one frame (jpgpu_parse_scans) stay together
        while (i < n && (k + 1 == parts || acc + descs[i].scan_len / 2 < goal || descs[i].frame_part == 2)) acc += descs[i++].scan_len;
    }
    first[parts] = n;
    return JPGPU_OK;
}

// Every device plans its own range of the partition.
extern "C" int jpgpu_multi_plan(jpgpu_multi* m, const jpgpu_image_desc* descs, size_t n) try {
    if (!m || (!descs && n)) return JPGPU_ERR_INVALID_ARG;
    m->n = n;
    const size_t nd = m->devs.size();
    std::vector<size_t> first(nd + 1);
    jpgpu_partition(descs, n, nd, first.data());
    for (size_t k = 0; k < nd; k++) { m->devs[k].first = first[k]; m->devs[k].count = first[k + 1] - first[k]; }
    return m->run_all([descs](Dev& d) {
        if (!d.batch) return jpgpu_batch_create(d.ctx, descs + d.first, d.count, &d.batch);
        return jpgpu_batch_replan(d.batch, descs + d.first, d.count);
    });
} JPGPU_CATCH_ALL

extern "C" int jpgpu_multi_range(const jpgpu_multi* m, int k, int* device, size_t* first, size_t* count) {
    if (!m || k < 0 || (size_t)k >= m->devs.size()) return JPGPU_ERR_INVALID_ARG;
    const Dev& d = m->devs[(size_t)k];
    if (device) *device = d.device;
    if (first) *first = d.first;
    if (count) *count = d.count;
    return JPGPU_OK;
}

extern "C" int jpgpu_multi_upload(jpgpu_multi* m) try {
    if (!m) return JPGPU_ERR_INVALID_ARG;
    return m->run_all([](Dev& d) { return d.batch ? jpgpu_batch_upload(d.batch) : (int)JPGPU_ERR_INVALID_ARG; });
} JPGPU_CATCH_ALL

extern "C" int jpgpu_multi_decode(jpgpu_multi* m) try {
    if (!m) return JPGPU_ERR_INVALID_ARG;
    return m->run_all([](Dev& d) { return d.batch ? jpgpu_batch_decode(d.batch) : (int)JPGPU_ERR_INVALID_ARG; });
} JPGPU_CATCH_ALL

extern "C" int jpgpu_multi_set_output_format(jpgpu_multi* m, uint32_t format) try {
    if (!m) return JPGPU_ERR_INVALID_ARG;
    return m->run_all([format](Dev& d) { return d.batch ? jpgpu_batch_set_output_format(d.batch, format) : (int)JPGPU_ERR_INVALID_ARG; });
} JPGPU_CATCH_ALL

extern "C" int jpgpu_multi_sync(jpgpu_multi* m) try {
    if (!m) return JPGPU_ERR_INVALID_ARG;
    return m->run_all([](Dev& d) { return jpgpu_sync(d.ctx); });
} JPGPU_CATCH_ALL

extern "C" int jpgpu_multi_download(jpgpu_multi* m, uint8_t* const* outs) try {
    if (!m || !outs) return JPGPU_ERR_INVALID_ARG;
    return m->run_all([outs](Dev& d) { return d.batch ? jpgpu_batch_download(d.batch, outs + d.first) : (int)JPGPU_ERR_INVALID_ARG; });
} JPGPU_CATCH_ALL

extern "C" int jpgpu_multi_results(jpgpu_multi* m, int32_t* statuses, uint64_t* bytes_read) try {
    if (!m) return JPGPU_ERR_INVALID_ARG;
    return m->run_all([statuses, bytes_read](Dev& d) {
        if (!d.batch) return (int)JPGPU_ERR_INVALID_ARG;
        return jpgpu_batch_results(d.batch, statuses ? statuses + d.first : nullptr, bytes_read ? bytes_read + d.first : nullptr);
    });
} JPGPU_CATCH_ALL

extern "C" void* jpgpu_multi_device_rgb(jpgpu_multi* m, size_t i, int* device, size_t* nbytes) {
    if (!m) return nullptr;
    Dev* d = m->owner(i);
    if (!d || !d->batch) return nullptr;
    if (device) *device = d->device;
    return jpgpu_batch_device_rgb(d->batch, i - d->first, nbytes);
}

extern "C" int jpgpu_multi_coefficients(jpgpu_multi* m, size_t i, int16_t* out, size_t cap, uint32_t nblocks[4]) try {
    if (!m) return JPGPU_ERR_INVALID_ARG;
    Dev* d = m->owner(i);
    if (!d || !d->batch) return JPGPU_ERR_INVALID_ARG;
    d->worker->post([=] { return jpgpu_batch_coefficients(d->batch, i - d->first, out, cap, nblocks); });
    return d->worker->wait();
} JPGPU_CATCH_ALL

extern "C" uint64_t jpgpu_multi_launch_count(const jpgpu_multi* m) {
    uint64_t n = 0;
    if (m) for (const Dev& d : m->devs) if (d.batch) n += jpgpu_batch_launch_count(d.batch);
    return n;
}

// Measurement aid: `steps` decodes of the planned (uploaded) images on every device, all devices released together;
// ms[k] = device time of device k's steps (CUDA events on its stream).  The job's time is the maximum.
extern "C" int jpgpu_multi_time_decode(jpgpu_multi* m, int steps, float* ms) try {
    if (!m || steps <= 0 || !ms) return JPGPU_ERR_INVALID_ARG;
    const int nd = (int)m->devs.size();
    m->gate.store(0);
    return m->run_all([m, steps, ms, nd](Dev& d) {
        int st = d.batch ? jpgpu_sync(d.ctx) : (int)JPGPU_ERR_INVALID_ARG;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (st == JPGPU_OK && (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess)) st = JPGPU_ERR_CUDA;
        cudaStream_t s = static_cast<cudaStream_t>(jpgpu_stream(d.ctx));
        m->gate.fetch_add(1);   // every worker passes the gate, failed or not, so that nobody waits for ever
        while (m->gate.load() < nd) std::this_thread::yield();   // every device idle and every worker here: go
        if (st != JPGPU_OK) { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); return st; }
        cudaEventRecord(e0, s);
        for (int k = 0; k < steps && st == JPGPU_OK; k++) st = jpgpu_batch_decode(d.batch);
        cudaEventRecord(e1, s);
        if (st == JPGPU_OK) st = jpgpu_sync(d.ctx);
        const size_t idx = (size_t)(&d - m->devs.data());
        if (st == JPGPU_OK && cudaEventElapsedTime(&ms[idx], e0, e1) != cudaSuccess) st = JPGPU_ERR_CUDA;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return st;
    });
} JPGPU_CATCH_ALL

// SURVEY 8(b) "jpgpu_decode_batch(handle, descs, n, outs, statuses, memory_kind)": plan, upload, decode on all devices;
// memory_kind HOST copies every image to outs[i] (W*H*3 bytes each), DEVICE leaves the pixels where they were produced
// (jpgpu_multi_device_rgb).  Synchronous.
extern "C" int jpgpu_multi_decode_batch(jpgpu_multi* m, const jpgpu_image_desc* descs, size_t n, uint8_t* const* outs,
                                        int32_t* statuses, uint64_t* bytes_read, uint32_t memory_kind) try {
    if (!m || (memory_kind == JPGPU_MEMORY_HOST && !outs) || memory_kind > JPGPU_MEMORY_DEVICE) return JPGPU_ERR_INVALID_ARG;
    int st = jpgpu_multi_plan(m, descs, n);
    if (st != JPGPU_OK) return st;
    st = m->run_all([outs, statuses, bytes_read, memory_kind](Dev& d) {
        int s = jpgpu_batch_upload(d.batch);
        if (s == JPGPU_OK) s = jpgpu_batch_decode(d.batch);
        if (s == JPGPU_OK && memory_kind == JPGPU_MEMORY_HOST) s = jpgpu_batch_download(d.batch, outs + d.first);
        if (s == JPGPU_OK) s = jpgpu_batch_results(d.batch, statuses ? statuses + d.first : nullptr, bytes_read ? bytes_read + d.first : nullptr);
        return s;
    });
    return st;
} JPGPU_CATCH_ALL
