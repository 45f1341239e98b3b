// staging.cu — host <-> device copies for PAGEABLE host arrays (what numpy hands us).
//
// cudaMemcpyAsync from pageable memory is staged through one driver bounce buffer by one thread (~11 GB/s measured
// on the B200 boxes, against ~55 GB/s from pinned memory). Large pageable copies are therefore split into chunks
// that a few host threads memcpy into their own pinned double buffers and DMA from there on their own streams, so
// the host-side memcpy of one chunk overlaps the DMA of another (the reference's `threads` argument sized a rayon
// pool; here host threads only ever move bytes). Pinned inputs and small arrays take the plain path.

#include <cstring>
#include <mutex>
#include <system_error>
#include <thread>

#include "common.cuh"

namespace pnbx {
namespace {

constexpr size_t CHUNK = 4u << 20;        // bytes per staged chunk
constexpr size_t STAGE_MIN = 16u << 20;   // below this the plain copy is as fast
constexpr int MAX_THREADS = 8;

struct Lane {  // per staging thread: two pinned buffers, one stream, two events
    void* buf[2] = {nullptr, nullptr};
    cudaStream_t s = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool ready = false;
};

// Lanes are per DEVICE (a multi-device call stages on all its devices at once); one staged transfer at a time per
// device.
constexpr int MAX_DEVICES = 64;
std::mutex g_mu[MAX_DEVICES];
Lane g_lanes[MAX_DEVICES][MAX_THREADS];
thread_local int tl_threads_override = -1;

int staging_threads() {
    static const int n = [] {
        const char* e = getenv("PNBX_STAGING_THREADS");
        int v = e ? atoi(e) : 8;  // measured: 4 threads ~25 GB/s, 8 threads ~50 GB/s (pinned-memory speed)
        unsigned hw = std::thread::hardware_concurrency();
        if (hw && (unsigned)v > hw) v = (int)hw;
        return std::max(0, std::min(v, MAX_THREADS));
    }();
    return tl_threads_override >= 0 ? std::min(tl_threads_override, n) : n;
}

bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

void teardown_lane(Lane& L) {
    if (L.s) cudaStreamDestroy(L.s);
    for (int i = 0; i < 2; ++i) {
        if (L.ev[i]) cudaEventDestroy(L.ev[i]);
        if (L.buf[i]) cudaFreeHost(L.buf[i]);
    }
    L = Lane{};
}
// The current device is the lane's device (staged_copy runs after make_exec selected it).
void prepare_lane(Lane& L) {
    if (L.ready) return;
    try {
        PNBX_CUDA(cudaStreamCreateWithFlags(&L.s, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            PNBX_CUDA(cudaHostAlloc(&L.buf[i], CHUNK, cudaHostAllocPortable));
            PNBX_CUDA(cudaEventCreateWithFlags(&L.ev[i], cudaEventDisableTiming));
        }
    } catch (...) {
        teardown_lane(L);  // never leave a half-built lane behind
        throw;
    }
    L.ready = true;
}

// dir = 0: host -> device, 1: device -> host. Runs the chunks k, k+nt, k+2nt, ... of lane k.
void lane_work(Lane& L, int device, int k, int nt, char* dev, char* host, size_t bytes, int dir, cudaError_t* err) {
    cudaSetDevice(device);
    const size_t nchunks = (bytes + CHUNK - 1) / CHUNK;
    cudaError_t e = cudaSuccess;
    int use = 0;
    size_t pending_off[2] = {0, 0}, pending_len[2] = {0, 0};
    bool pending[2] = {false, false};
    for (size_t c = (size_t)k; c < nchunks && e == cudaSuccess; c += (size_t)nt, use ^= 1) {
        const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
        if (pending[use]) {  // this pinned buffer is still in flight from two chunks ago
            e = cudaEventSynchronize(L.ev[use]);
            if (dir == 1 && e == cudaSuccess) memcpy(host + pending_off[use], L.buf[use], pending_len[use]);
            pending[use] = false;
        }
        if (e != cudaSuccess) break;
        if (dir == 0) {
            memcpy(L.buf[use], host + off, len);
            e = cudaMemcpyAsync(dev + off, L.buf[use], len, cudaMemcpyHostToDevice, L.s);
        } else {
            e = cudaMemcpyAsync(L.buf[use], dev + off, len, cudaMemcpyDeviceToHost, L.s);
        }
        if (e == cudaSuccess) e = cudaEventRecord(L.ev[use], L.s);
        pending[use] = true;
        pending_off[use] = off;
        pending_len[use] = len;
    }
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        const int b = use ^ i;  // drain in issue order
        if (!pending[b]) continue;
        e = cudaEventSynchronize(L.ev[b]);
        if (dir == 1 && e == cudaSuccess) memcpy(host + pending_off[b], L.buf[b], pending_len[b]);
    }
    if (e != cudaSuccess) cudaStreamSynchronize(L.s);  // nothing of this lane may still be in flight when we report
    *err = e;
}

struct EventHolder {
    cudaEvent_t e = nullptr;
    ~EventHolder() { if (e) cudaEventDestroy(e); }
};

void staged_copy(void* dev, void* host, size_t bytes, int dir, const Exec& ex) {
    const int nt = staging_threads();
    const int d = ex.device < MAX_DEVICES ? ex.device : 0;
    std::lock_guard<std::mutex> lock(g_mu[d]);
    EventHolder ready;
    PNBX_CUDA(cudaEventCreateWithFlags(&ready.e, cudaEventDisableTiming));
    PNBX_CUDA(cudaEventRecord(ready.e, ex.stream));  // allocation of `dev` / the kernels that produced it
    for (int k = 0; k < nt; ++k) {
        prepare_lane(g_lanes[d][k]);
        PNBX_CUDA(cudaStreamWaitEvent(g_lanes[d][k].s, ready.e, 0));
    }
    cudaError_t errs[MAX_THREADS];
    for (int k = 0; k < MAX_THREADS; ++k) errs[k] = cudaSuccess;
    std::thread th[MAX_THREADS];
    int started = 0;
    bool spawn_failed = false;
    for (int k = 0; k < nt; ++k) {
        try {
            th[k] = std::thread(lane_work, std::ref(g_lanes[d][k]), ex.device, k, nt, (char*)dev, (char*)host, bytes, dir, &errs[k]);
            ++started;
        } catch (const std::system_error&) {  // out of threads: the chunks of the missing lanes would be lost
            spawn_failed = true;
            break;
        }
    }
    for (int k = 0; k < started; ++k) th[k].join();
    if (spawn_failed) throw ArgError{PNBX_ERR_CUDA, "staged host<->device copy: could not start the staging threads"};
    for (int k = 0; k < nt; ++k)
        if (errs[k] != cudaSuccess) throw CudaError{errs[k], "staged host<->device copy", __FILE__, __LINE__};
    // every lane drained its stream before returning: the data is in place, later work on ex.stream is ordered by
    // program order on the host
}

}  // namespace

void set_staging_threads_for_this_thread(int n) { tl_threads_override = n; }

void copy_h2d(void* dev, const void* host, size_t bytes, const Exec& ex) {
    if (bytes >= STAGE_MIN && staging_threads() > 0 && is_pageable(host)) {
        staged_copy(dev, const_cast<void*>(host), bytes, 0, ex);
        return;
    }
    PNBX_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ex.stream));
}

void copy_d2h(void* host, const void* dev, size_t bytes, const Exec& ex) {
    if (bytes >= STAGE_MIN && staging_threads() > 0 && is_pageable(host)) {
        staged_copy(const_cast<void*>(dev), host, bytes, 1, ex);
        return;
    }
    PNBX_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ex.stream));
}

}  // namespace pnbx
