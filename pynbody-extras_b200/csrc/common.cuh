// common.cuh — shared host-side plumbing of libpnbx_gravity.so (error state, device buffers,
// stage timers, option handling). No CPU compute lives in this library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

#include "../../include/pnbx_gravity.h"

namespace pnbx {

// ---- thread-local error + timing state (pnbx_last_error / pnbx_last_timings) ----
std::string& last_error();
struct StageTimes {
    std::vector<std::pair<const char*, double>> v;
};
StageTimes& last_times();
bool timing_enabled();  // GRAVITY_TIMING env var, same rule as tree.rs:5-16

int fail(int code, const std::string& msg);

struct CudaError {
    cudaError_t e;
    const char* what;
    const char* file;
    int line;
};

#define PNBX_CUDA(expr)                                                             \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) throw ::pnbx::CudaError{_e, #expr, __FILE__, __LINE__}; \
    } while (0)

struct ArgError {
    int code;
    std::string msg;
};

// Restores the calling thread's current CUDA device on scope exit: entry points select the device they work on
// (pnbx_opts.device / the tree's device) and must not leave it changed for the host application.
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Run `body`, translating exceptions into status codes + last_error().
template <class F>
int guarded(F&& body) {
    DeviceGuard restore_device;
    try {
        last_error().clear();
        body();
        return PNBX_OK;
    } catch (const CudaError& ce) {
        char buf[512];
        snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", (int)ce.e, cudaGetErrorString(ce.e),
                 ce.file, ce.line, ce.what);
        cudaGetLastError();  // clear sticky-less errors
        return fail(PNBX_ERR_CUDA, buf);
    } catch (const ArgError& ae) {
        return fail(ae.code, ae.msg);
    } catch (const std::exception& ex) {
        return fail(PNBX_ERR_CUDA, ex.what());
    }
}

// ---- execution context derived from pnbx_opts ----
struct Exec {
    int device = 0;
    bool device_ptrs = false;
    bool f64 = false;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    bool kernel_events = false;
    bool tree_order = false;
    bool block_cyclic = false;
    int shard_rank = 0, shard_world = 1;
    int64_t shard_block = 4096;
};
Exec make_exec(const pnbx_opts* opts);  // selects the device; throws if none
void finish_exec(Exec& ex);             // sync (host mode) + release

// Stream-ordered allocations come from a LIBRARY-OWNED caching allocator (capi.cu): cudaMalloc blocks kept on a per
// device free list with a last-use event, so repeated calls reuse their memory without touching the process-wide
// default pool's settings, and peers can access the blocks (multi.cu). pnbx_trim_memory() returns the cache.
void* pool_alloc(size_t bytes, cudaStream_t s);  // on the current device
void pool_free(void* p, cudaStream_t s);
void pool_trim();

// Stream-ordered device buffer, or a non-owning view into an Arena slab.
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    bool owned = true;
    DevBuf() = default;
    DevBuf(size_t count, cudaStream_t stream) { alloc(count, stream); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s), owned(o.owned) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p; n = o.n; s = o.s; owned = o.owned;
            o.p = nullptr; o.n = 0;
        }
        return *this;
    }
    void alloc(size_t count, cudaStream_t stream) {
        release();
        s = stream;
        n = count;
        owned = true;
        if (count) p = static_cast<T*>(pool_alloc(count * sizeof(T), stream));
    }
    void view(T* ptr, size_t count) {  // memory owned by an Arena
        release();
        p = ptr; n = count; owned = false;
    }
    void release() {
        if (p && owned) pool_free(p, s);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    T* get() const { return p; }
    size_t bytes() const { return n * sizeof(T); }
};

// One slab, many arrays: the layout code runs twice, first measuring (take() returns nullptr and only adds up
// sizes), then — after commit() made the single allocation — handing out 256-byte aligned pointers. A tree build is
// a handful of slab allocations instead of ~60 pool calls whose cost depends on the pool's history.
struct Arena {
    char* base = nullptr;
    size_t cap = 0, off = 0;
    cudaStream_t s = nullptr;
    bool measuring = true;
    Arena() = default;
    Arena(const Arena&) = delete;
    Arena& operator=(const Arena&) = delete;
    template <class T>
    T* take(size_t count) {
        const size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
        T* r = measuring ? nullptr : reinterpret_cast<T*>(base + off);
        off += bytes;
        return r;
    }
    template <class T>
    void take(DevBuf<T>& b, size_t count) { b.view(take<T>(count), count); }
    void commit(cudaStream_t stream) {
        const size_t need = off;  // measured by the first layout pass
        release();
        s = stream;
        cap = need ? need : 256;
        base = static_cast<char*>(pool_alloc(cap, stream));
        off = 0;
        measuring = false;
    }
    void release() {
        if (base) pool_free(base, s);
        base = nullptr;
        cap = off = 0;
        measuring = true;
    }
    ~Arena() { release(); }
};

// staging.cu: host<->device copies; large PAGEABLE host arrays are staged through pinned buffers by a few host threads
void copy_h2d(void* dev, const void* host, size_t bytes, const Exec& ex);
void copy_d2h(void* host, const void* dev, size_t bytes, const Exec& ex);
void set_staging_threads_for_this_thread(int n);  // multi-device rank threads share the host cores; < 0 = default

// Input array that may live on the host (copied in, like gravity.rs:154-180) or on the device.
template <class T>
struct InArray {
    const T* d = nullptr;
    DevBuf<T> owned;
    void bind(const T* src, size_t count, const Exec& ex) {
        if (!src || count == 0) { d = nullptr; return; }
        if (ex.device_ptrs) { d = src; return; }
        owned.alloc(count, ex.stream);
        copy_h2d(owned.p, src, count * sizeof(T), ex);
        d = owned.p;
    }
};

// Output array: device pointer given by the caller, or a temp that is copied back to the host.
template <class T>
struct OutArray {
    T* d = nullptr;
    T* host = nullptr;
    size_t n = 0;
    DevBuf<T> owned;
    void bind(T* dst, size_t count, const Exec& ex) {
        n = count;
        if (!dst || count == 0) { d = nullptr; return; }
        if (ex.device_ptrs) { d = dst; return; }
        owned.alloc(count, ex.stream);
        d = owned.p;
        host = dst;
    }
    void finish(const Exec& ex) {
        if (host && n) copy_d2h(host, d, n * sizeof(T), ex);
    }
};

// CUDA-event stage timer; records into last_times() when GRAVITY_TIMING is set.
struct StageTimer {
    cudaStream_t s;
    bool on;
    cudaEvent_t a = nullptr, b = nullptr;
    const char* label = nullptr;
    explicit StageTimer(cudaStream_t stream);
    ~StageTimer();
    void begin(const char* l);
    void end();
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Every kernel launch of this library goes through PNBX_LAUNCH so bench.py can report `gpu_launches`.
std::atomic<int64_t>& launch_counter();
#define PNBX_LAUNCH(kernel, grid, block, smem, stream, ...)          \
    do {                                                             \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \
        ++::pnbx::launch_counter();                                  \
    } while (0)

// CUDA events around the dominant kernel of the last call (no sync inside the call): recorded on the
// call's stream when pnbx_opts.flags has PNBX_FLAG_KERNEL_EVENTS; read back by pnbx_last_kernel_ms().
struct KernelEvents {
    static constexpr int MAXDEV = 64;
    cudaEvent_t a[MAXDEV] = {}, b[MAXDEV] = {};  // one pair per device: events belong to the device they were created on
    int dev = 0;                                 // device of the last armed call
    bool armed = false, valid = false;
    void begin(cudaStream_t s);
    void end(cudaStream_t s);
};
KernelEvents& kernel_events();

// direct.cu: the direct sum on ONE device, device-resident float64 inputs, stream-ordered on ex.stream
void direct_on_device(const Exec& ex, const double* d_pos, const double* d_mass, const double* d_h, int64_t n,
                      const double* d_tgt, int64_t m, int64_t tgt_begin, int kernel, int want, double* d_pot,
                      double* d_acc, StageTimer& tm);

// multi.cu: PNBX_DEVICES — several GPUs behind the unchanged host API. multi_* return false when the call should take
// the single-device path (knob unset, problem too small, no peer access).
constexpr int MAX_SLICES = 8;
struct OutSlices {  // result exchange of the multi-device tree walk: slice r = original indices [bounds[r], bounds[r+1])
    int n = 0;      // lives in the memory of device r; 0 slices = plain output arrays
    int64_t bounds[MAX_SLICES + 1] = {};
    double* pot[MAX_SLICES] = {};
    double* acc[MAX_SLICES] = {};
};
std::vector<int> multi_devices();
bool multi_direct(const double* src_pos, const double* src_mass, const double* src_h, int64_t n, const double* tgt_pos,
                  int64_t m, int kernel, int want, double* out_pot, double* out_acc, const pnbx_opts* opts);

// Shared by direct.cu and tree.cu: float64 bounding box of (n,3) positions on the device.
// out6 = {minx,miny,minz,maxx,maxy,maxz} (device). Follows tree.rs:628-640.
void launch_bbox(const double* pos, int64_t n, double* out6, cudaStream_t s);

}  // namespace pnbx
