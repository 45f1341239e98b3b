// capi.cu — error/timing state, option handling and the shared bounding-box reduction.
#include <cfloat>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace pnbx {

std::string& last_error() {
    thread_local std::string e;
    return e;
}
StageTimes& last_times() {
    thread_local StageTimes t;
    return t;
}
bool timing_enabled() {
    static const bool on = [] {
        const char* v = getenv("GRAVITY_TIMING");
        if (!v) return false;
        std::string s(v);
        size_t a = s.find_first_not_of(" \t"), b = s.find_last_not_of(" \t");
        s = (a == std::string::npos) ? "" : s.substr(a, b - a + 1);
        if (s.empty() || s == "0") return false;
        std::string low;
        for (char c : s) low += (char)tolower(c);
        return low != "false";
    }();
    return on;
}

std::atomic<int64_t>& launch_counter() {  // concurrent calls from several host threads are legal
    static std::atomic<int64_t> c{0};
    return c;
}
KernelEvents& kernel_events() {
    thread_local KernelEvents k;
    return k;
}
void KernelEvents::begin(cudaStream_t s) {
    if (!armed) return;
    if (!a[dev]) {  // created on the device the call runs on (make_exec selected it)
        cudaEventCreate(&a[dev]);
        cudaEventCreate(&b[dev]);
    }
    cudaEventRecord(a[dev], s);
}
void KernelEvents::end(cudaStream_t s) {
    if (!armed) return;
    cudaEventRecord(b[dev], s);
    valid = true;
}

// ---- library-owned caching allocator
// Stream-ordered like cudaMallocAsync, but the cache belongs to this library (no settings of the process-wide default
// pool are touched) and the memory is plain cudaMalloc memory, which peers can read and write once peer access is
// enabled (multi.cu). free: the block goes to the device's free list together with an event recorded on the freeing
// stream; alloc: best fit from the list, the allocating stream waits for that event first. Blocks are returned to the
// driver only by pnbx_trim_memory() or when an allocation fails.
namespace {
struct CachedBlock {
    void* p;
    size_t bytes;
    cudaEvent_t ev;  // last use, recorded at free time
};
struct DeviceCache {
    std::mutex mu;
    std::multimap<size_t, CachedBlock> free_blocks;      // by size
    std::unordered_map<void*, CachedBlock> live;         // handed out
};
DeviceCache g_cache[KernelEvents::MAXDEV];

size_t round_size(size_t bytes) {
    const size_t g = bytes >= (size_t(1) << 20) ? (size_t(2) << 20) : 512;  // 2 MB granules for large blocks
    return (bytes + g - 1) / g * g;
}
void trim_locked(DeviceCache& c) {
    for (auto& kv : c.free_blocks) {
        cudaEventSynchronize(kv.second.ev);
        cudaFree(kv.second.p);
        cudaEventDestroy(kv.second.ev);
    }
    c.free_blocks.clear();
}
}  // namespace

void* pool_alloc(size_t bytes, cudaStream_t s) {
    int dev = 0;
    PNBX_CUDA(cudaGetDevice(&dev));
    if (dev >= KernelEvents::MAXDEV) throw ArgError{PNBX_ERR_ARG, "device ordinal out of range"};
    DeviceCache& c = g_cache[dev];
    const size_t need = round_size(bytes ? bytes : 1);
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.free_blocks.lower_bound(need);
    if (it != c.free_blocks.end() && it->first <= need + need / 2 + (size_t(1) << 20)) {  // do not burn a huge block on a small request
        CachedBlock b = it->second;
        c.free_blocks.erase(it);
        PNBX_CUDA(cudaStreamWaitEvent(s, b.ev, 0));  // whatever used the block before has to finish first
        c.live.emplace(b.p, b);
        return b.p;
    }
    CachedBlock b{nullptr, need, nullptr};
    cudaError_t e = cudaMalloc(&b.p, need);
    if (e != cudaSuccess) {  // give the cache back to the driver and try once more
        cudaGetLastError();
        trim_locked(c);
        e = cudaMalloc(&b.p, need);
    }
    if (e != cudaSuccess) throw CudaError{e, "cudaMalloc (library cache)", __FILE__, __LINE__};
    PNBX_CUDA(cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming));
    c.live.emplace(b.p, b);
    return b.p;
}
void pool_free(void* p, cudaStream_t s) {
    if (!p) return;
    cudaPointerAttributes at;
    int dev = 0;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess) dev = at.device;
    else cudaGetLastError();
    DeviceCache& c = g_cache[dev < KernelEvents::MAXDEV ? dev : 0];
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.live.find(p);
    if (it == c.live.end()) return;  // not ours
    CachedBlock b = it->second;
    c.live.erase(it);
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != dev) cudaSetDevice(dev);  // the event belongs to the block's device
    cudaEventRecord(b.ev, s);
    if (cur != dev) cudaSetDevice(cur);
    c.free_blocks.emplace(b.bytes, b);
}
void pool_trim() {
    for (auto& c : g_cache) {
        std::lock_guard<std::mutex> lock(c.mu);
        if (c.free_blocks.empty()) continue;
        cudaPointerAttributes at;
        int cur = 0;
        cudaGetDevice(&cur);
        if (cudaPointerGetAttributes(&at, c.free_blocks.begin()->second.p) == cudaSuccess) cudaSetDevice(at.device);
        trim_locked(c);
        cudaSetDevice(cur);
    }
}

int fail(int code, const std::string& msg) {
    last_error() = msg;
    return code;
}

Exec make_exec(const pnbx_opts* opts) {
    Exec ex;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        throw ArgError{PNBX_ERR_CUDA,
                       "no CUDA device available: libpnbx_gravity has no CPU fallback (B200 / sm_100a required)"};
    }
    int dev = opts ? opts->device : -1;
    if (dev < 0) PNBX_CUDA(cudaGetDevice(&dev));
    if (dev >= ndev) throw ArgError{PNBX_ERR_ARG, "pnbx_opts.device out of range"};
    PNBX_CUDA(cudaSetDevice(dev));
    ex.device = dev;
    ex.device_ptrs = opts && opts->mem_space == PNBX_MEM_DEVICE;
    ex.f64 = opts && opts->precision == PNBX_PREC_F64;
    ex.kernel_events = opts && (opts->flags & PNBX_FLAG_KERNEL_EVENTS);
    ex.tree_order = opts && (opts->flags & PNBX_FLAG_TREE_ORDER);
    ex.block_cyclic = ex.tree_order && (opts->flags & PNBX_FLAG_BLOCK_CYCLIC);
    if (ex.block_cyclic) {
        ex.shard_world = opts->shard_world;
        ex.shard_rank = opts->shard_rank;
        ex.shard_block = opts->shard_block > 0 ? opts->shard_block : 4096;
        if (ex.shard_world < 1 || ex.shard_rank < 0 || ex.shard_rank >= ex.shard_world)
            throw ArgError{PNBX_ERR_ARG, "bad shard_rank / shard_world"};
    }
    kernel_events().armed = ex.kernel_events;
    if (ex.kernel_events) { kernel_events().valid = false; kernel_events().dev = dev < KernelEvents::MAXDEV ? dev : 0; }
    if (opts && (opts->stream || ex.device_ptrs)) {
        // device pointers are ordered against the caller's stream: NULL means the CUDA default stream
        ex.stream = (cudaStream_t)opts->stream;
    } else {
        // one library stream per (thread, device), created lazily and kept
        thread_local cudaStream_t streams[64] = {};
        if (dev < 64) {
            if (!streams[dev]) PNBX_CUDA(cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking));
            ex.stream = streams[dev];
        } else {
            PNBX_CUDA(cudaStreamCreateWithFlags(&ex.stream, cudaStreamNonBlocking));
            ex.own_stream = true;
        }
    }
    return ex;
}

void finish_exec(Exec& ex) {
    if (!ex.device_ptrs) PNBX_CUDA(cudaStreamSynchronize(ex.stream));
    if (ex.own_stream) {
        cudaStreamDestroy(ex.stream);
        ex.own_stream = false;
    }
}

StageTimer::StageTimer(cudaStream_t stream) : s(stream), on(timing_enabled()) {
    if (on) {
        last_times().v.clear();
        cudaEventCreate(&a);
        cudaEventCreate(&b);
    }
}
StageTimer::~StageTimer() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
}
void StageTimer::begin(const char* l) {
    if (!on) return;
    label = l;
    cudaEventRecord(a, s);
}
void StageTimer::end() {
    if (!on) return;
    cudaEventRecord(b, s);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    last_times().v.emplace_back(label, (double)ms);
    fprintf(stderr, "[gravity-timing] %s: %.3f ms\n", label, ms);
}

// ---- bounding box: block-wide min/max in float64, one partial per block, final pass by one block.
namespace {
constexpr int BB_THREADS = 256;

__device__ inline void warp_minmax(double (&mn)[3], double (&mx)[3]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double a = __shfl_down_sync(0xffffffffu, mn[k], o);
            double b = __shfl_down_sync(0xffffffffu, mx[k], o);
            if (a < mn[k]) mn[k] = a;  // same comparisons as tree.rs:633-638
            if (b > mx[k]) mx[k] = b;
        }
    }
}

__global__ void __launch_bounds__(BB_THREADS) bbox_partial(const double* __restrict__ pos, int64_t n,
                                                           double* __restrict__ partial) {
    double mn[3] = {INFINITY, INFINITY, INFINITY};
    double mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    // one particle (24 contiguous bytes) per thread per sweep: a warp covers 768 contiguous bytes
    const int64_t total = 3 * n;
    const int64_t stride = (int64_t)gridDim.x * BB_THREADS * 3;
    for (int64_t base = ((int64_t)blockIdx.x * BB_THREADS + threadIdx.x) * 3; base < total; base += stride) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double v = pos[base + k];
            if (v < mn[k]) mn[k] = v;
            if (v > mx[k]) mx[k] = v;
        }
    }
    warp_minmax(mn, mx);
    __shared__ double sm[BB_THREADS / 32][6];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        for (int k = 0; k < 3; ++k) { sm[w][k] = mn[k]; sm[w][3 + k] = mx[k]; }
    }
    __syncthreads();
    if (w == 0) {
        for (int k = 0; k < 3; ++k) {
            mn[k] = l < BB_THREADS / 32 ? sm[l][k] : INFINITY;
            mx[k] = l < BB_THREADS / 32 ? sm[l][3 + k] : -INFINITY;
        }
        warp_minmax(mn, mx);
        if (l == 0)
            for (int k = 0; k < 3; ++k) {
                partial[blockIdx.x * 6 + k] = mn[k];
                partial[blockIdx.x * 6 + 3 + k] = mx[k];
            }
    }
}

__global__ void __launch_bounds__(BB_THREADS) bbox_final(const double* __restrict__ partial, int nparts,
                                                         double* __restrict__ out6) {
    double mn[3] = {INFINITY, INFINITY, INFINITY};
    double mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = threadIdx.x; i < nparts; i += BB_THREADS)
        for (int k = 0; k < 3; ++k) {
            double a = partial[i * 6 + k], b = partial[i * 6 + 3 + k];
            if (a < mn[k]) mn[k] = a;
            if (b > mx[k]) mx[k] = b;
        }
    warp_minmax(mn, mx);
    __shared__ double sm[BB_THREADS / 32][6];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0)
        for (int k = 0; k < 3; ++k) { sm[w][k] = mn[k]; sm[w][3 + k] = mx[k]; }
    __syncthreads();
    if (w == 0) {
        for (int k = 0; k < 3; ++k) {
            mn[k] = l < BB_THREADS / 32 ? sm[l][k] : INFINITY;
            mx[k] = l < BB_THREADS / 32 ? sm[l][3 + k] : -INFINITY;
        }
        warp_minmax(mn, mx);
        if (l == 0)
            for (int k = 0; k < 3; ++k) { out6[k] = mn[k]; out6[3 + k] = mx[k]; }
    }
}
}  // namespace

void launch_bbox(const double* pos, int64_t n, double* out6, cudaStream_t s) {
    int blocks = (int)std::min<int64_t>(148 * 4, std::max<int64_t>(1, ceil_div(n, BB_THREADS * 4)));
    DevBuf<double> partial((size_t)blocks * 6, s);
    PNBX_LAUNCH(bbox_partial, blocks, BB_THREADS, 0, s, pos, n, partial.get());
    PNBX_LAUNCH(bbox_final, 1, BB_THREADS, 0, s, partial.get(), blocks, out6);
    PNBX_CUDA(cudaGetLastError());
}

}  // namespace pnbx

extern "C" {

const char* pnbx_last_error(void) { return pnbx::last_error().c_str(); }
int pnbx_abi_version(void) { return PNBX_ABI_VERSION; }
int pnbx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
int64_t pnbx_shard_count(int64_t n, int64_t block, int32_t world, int32_t rank) {
    if (block <= 0) block = 4096;
    if (n <= 0 || world < 1 || rank < 0 || rank >= world) return 0;
    const int64_t full = n / block, tail = n % block;
    int64_t cnt = (full / world + (rank < full % world ? 1 : 0)) * block;
    if (tail && full % world == rank) cnt += tail;
    return cnt;
}
int pnbx_last_kernel_ms(double* ms) {
    auto& k = pnbx::kernel_events();
    if (!k.valid || !ms) return pnbx::fail(PNBX_ERR_STATE, "no kernel events recorded (PNBX_FLAG_KERNEL_EVENTS not set)");
    if (cudaEventSynchronize(k.b[k.dev]) != cudaSuccess) return pnbx::fail(PNBX_ERR_CUDA, "cudaEventSynchronize failed");
    float f = 0.f;
    if (cudaEventElapsedTime(&f, k.a[k.dev], k.b[k.dev]) != cudaSuccess) return pnbx::fail(PNBX_ERR_CUDA, "cudaEventElapsedTime failed");
    *ms = (double)f;
    return PNBX_OK;
}
int64_t pnbx_launch_count(void) { return pnbx::launch_counter().load(); }
void pnbx_trim_memory(void) { pnbx::pool_trim(); }
int pnbx_last_timings(const char** labels, double* ms, int cap) {
    auto& v = pnbx::last_times().v;
    int k = 0;
    for (; k < (int)v.size() && k < cap; ++k) {
        if (labels) labels[k] = v[k].first;
        if (ms) ms[k] = v[k].second;
    }
    return k;
}

}  // extern "C"
