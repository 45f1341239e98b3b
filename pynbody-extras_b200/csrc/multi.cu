// multi.cu — several GPUs of one box behind the UNCHANGED host API.
//
// The reference is one process with a rayon pool (gravity.rs:87-101); its callers hold numpy arrays and make one
// call. With PNBX_DEVICES=all (or a list "0,1,2,3") the same call — pnbx_direct / pnbx_tree_create / pnbx_tree_eval
// with HOST pointers and no explicit device — is spread over the listed GPUs by one host thread per GPU:
//   sources : every GPU uploads 1/W of the rows (W PCIe links in parallel) and pulls the other shards from its
//             peers over NVLink (cudaMemcpyPeerAsync into the same offsets): an all-gather without a collective
//             library, W-1 bulk copies per GPU;
//   targets : independent, so they are sharded with no reduction. Direct sums take contiguous shards;
//             tree self-evaluations take block-cyclic shards in TREE order (coherent warps, equal cost per GPU) and
//             the walk kernel itself stores every result into the slice of the GPU that owns the particle's
//             ORIGINAL index (peer stores over NVLink, fused into the kernel's epilogue), so each GPU finishes with
//             a contiguous slice of the caller's array and copies it back in one piece;
//   tree    : every GPU builds the identical tree from the gathered sources (deterministic kernels).
// Results are bit-identical to the single-GPU shard calls (tgt_begin / PNBX_FLAG_BLOCK_CYCLIC) they are made of.
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

#include "tree.cuh"

namespace pnbx {
namespace {

std::vector<int> parse_devices() {
    std::vector<int> out;
    const char* e = getenv("PNBX_DEVICES");
    if (!e || !*e) return out;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); return out; }
    std::string v(e);
    if (v == "all" || v == "ALL") {
        for (int i = 0; i < ndev; ++i) out.push_back(i);
    } else {
        size_t pos = 0;
        while (pos < v.size()) {
            size_t q = v.find(',', pos);
            if (q == std::string::npos) q = v.size();
            const std::string tok = v.substr(pos, q - pos);
            pos = q + 1;
            if (tok.empty()) continue;
            char* end = nullptr;
            long d = strtol(tok.c_str(), &end, 10);
            if (*end || d < 0 || d >= ndev) throw ArgError{PNBX_ERR_ARG, "PNBX_DEVICES: bad device list (use \"all\" or e.g. \"0,1,2\")"};
            bool dup = false;
            for (int x : out) dup |= x == (int)d;
            if (!dup) out.push_back((int)d);
        }
    }
    if ((int)out.size() > MAX_SLICES) out.resize(MAX_SLICES);
    if (out.size() < 2) out.clear();
    // the result exchange stores into peer memory: every pair must be peer-accessible
    for (int a : out)
        for (int b : out) {
            if (a == b) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can) { cudaGetLastError(); return {}; }
        }
    return out;
}

double env_number(const char* name, double dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atof(e) : dflt;
}

// Host-side barrier for the rank threads; abort() releases everybody when one rank failed.
struct HostBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int n, count = 0, gen = 0;
    bool aborted = false;
    explicit HostBarrier(int ranks) : n(ranks) {}
    void arrive_and_wait() {
        std::unique_lock<std::mutex> lk(mu);
        if (aborted) throw ArgError{PNBX_ERR_CUDA, "another device of the multi-device call failed"};
        const int g = gen;
        if (++count == n) { count = 0; ++gen; cv.notify_all(); return; }
        cv.wait(lk, [&] { return gen != g || aborted; });
        if (gen == g && aborted) throw ArgError{PNBX_ERR_CUDA, "another device of the multi-device call failed"};
    }
    void abort() {
        std::lock_guard<std::mutex> lk(mu);
        aborted = true;
        cv.notify_all();
    }
};

// Persistent per-device worker stream + event (rank threads are short-lived; their streams must not be).
struct RankCtx {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev = nullptr, ev2 = nullptr;
    bool peers_enabled = false;
};
RankCtx g_ctx[KernelEvents::MAXDEV];
std::mutex g_multi_mu;  // one multi-device call at a time: it uses every listed GPU anyway

RankCtx& rank_ctx(int dev, const std::vector<int>& devs) {  // current device == dev
    RankCtx& c = g_ctx[dev];
    if (!c.stream) {
        PNBX_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        PNBX_CUDA(cudaEventCreateWithFlags(&c.ev, cudaEventDisableTiming));
        PNBX_CUDA(cudaEventCreateWithFlags(&c.ev2, cudaEventDisableTiming));
    }
    if (!c.peers_enabled) {
        for (int p : devs)
            if (p != dev && cudaDeviceEnablePeerAccess(p, 0) != cudaSuccess) cudaGetLastError();  // "already enabled" is fine
        c.peers_enabled = true;
    }
    return c;
}

struct RankError {
    int code = PNBX_OK;
    std::string msg;
};

// Runs body(r) on one host thread per device; the first failure is re-thrown on the calling thread.
template <class F>
void run_ranks(const std::vector<int>& devs, HostBarrier& bar, F&& body) {
    const int W = (int)devs.size();
    std::vector<RankError> errs((size_t)W);
    std::vector<std::thread> th;
    th.reserve((size_t)W);
    // Staging lanes per rank thread, sized from the machine's core count. Sizing them from the affinity mask instead
    // (a 4-GPU container with few CPUs in its mask) was measured 6x slower on the 4 GB upload of config 4 (790 vs
    // 137 ms): the lanes spend their time in page faults and copies into pinned memory, not on the cores.
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int stage_threads = (int)std::max(1u, std::min(8u, hw / (unsigned)W));
    for (int r = 0; r < W; ++r) {
        th.emplace_back([&, r] {
            try {
                PNBX_CUDA(cudaSetDevice(devs[(size_t)r]));
                set_staging_threads_for_this_thread(stage_threads);
                body(r);
            } catch (const CudaError& ce) {
                char buf[512];
                snprintf(buf, sizeof(buf), "device %d: CUDA error %d (%s) at %s:%d: %s", devs[(size_t)r], (int)ce.e,
                         cudaGetErrorString(ce.e), ce.file, ce.line, ce.what);
                cudaGetLastError();
                errs[(size_t)r] = {PNBX_ERR_CUDA, buf};
                bar.abort();
            } catch (const ArgError& ae) {
                errs[(size_t)r] = {ae.code, ae.msg};
                bar.abort();
            } catch (const std::exception& ex) {
                errs[(size_t)r] = {PNBX_ERR_CUDA, ex.what()};
                bar.abort();
            }
        });
    }
    for (auto& t : th) t.join();
    const RankError* first = nullptr;
    for (const auto& e : errs)  // prefer the root cause over "another device failed"
        if (e.code != PNBX_OK && (!first || first->msg.find("another device") != std::string::npos)) first = &e;
    if (first) throw ArgError{first->code, first->msg};
}

Exec rank_exec(int dev, cudaStream_t s, const pnbx_opts* opts) {
    Exec ex;
    ex.device = dev;
    ex.device_ptrs = false;
    ex.f64 = opts && opts->precision == PNBX_PREC_F64;
    ex.stream = s;
    return ex;
}

// all-gather by peer pulls: `mine[r]` holds rows [b[r], b[r+1]) of an array with `width` doubles per row; after the
// call (stream-ordered on s) device r holds every row. ready[p] was recorded after p's own shard was in place.
void pull_shards(int r, const std::vector<int>& devs, const std::vector<double*>& ptr, const std::vector<cudaEvent_t>& ready,
                 const std::vector<int64_t>& b, int width, cudaStream_t s) {
    const int W = (int)devs.size();
    for (int k = 1; k < W; ++k) {
        const int p = (r + k) % W;  // rotated: at any moment every GPU is read by a different peer
        const int64_t rows = b[(size_t)p + 1] - b[(size_t)p];
        if (rows == 0) continue;
        PNBX_CUDA(cudaStreamWaitEvent(s, ready[(size_t)p], 0));
        const size_t off = (size_t)b[(size_t)p] * width;
        PNBX_CUDA(cudaMemcpyPeerAsync(ptr[(size_t)r] + off, devs[(size_t)r], ptr[(size_t)p] + off, devs[(size_t)p],
                                      (size_t)rows * width * sizeof(double), s));
    }
}

std::vector<int64_t> even_bounds(int64_t n, int W) {
    std::vector<int64_t> b((size_t)W + 1);
    for (int r = 0; r <= W; ++r) b[(size_t)r] = (n * r) / W;
    return b;
}

}  // namespace

std::vector<int> multi_devices() { return parse_devices(); }

// ------------------------------------------------------------------------------------------ direct
bool multi_direct(const double* src_pos, const double* src_mass, const double* src_h, int64_t n, const double* tgt_pos,
                  int64_t m, int kernel, int want, double* out_pot, double* out_acc, const pnbx_opts* opts) {
    std::vector<int> devs = multi_devices();
    if (devs.size() < 2) return false;
    if ((double)n * (double)m < env_number("PNBX_MULTI_MIN_WORK", 1e10)) return false;
    if ((int64_t)devs.size() > m) devs.resize((size_t)std::max<int64_t>(m, 1));
    const int W = (int)devs.size();
    if (W < 2) return false;
    std::lock_guard<std::mutex> lock(g_multi_mu);
    const bool self = tgt_pos == nullptr;
    const std::vector<int64_t> sb = even_bounds(n, W), tb = even_bounds(m, W);
    std::vector<double*> p_pos((size_t)W, nullptr), p_mass((size_t)W, nullptr), p_h((size_t)W, nullptr);
    std::vector<cudaEvent_t> ready((size_t)W, nullptr);
    HostBarrier bar(W);
    run_ranks(devs, bar, [&](int r) {
        const int dev = devs[(size_t)r];
        RankCtx& c = rank_ctx(dev, devs);
        cudaStream_t s = c.stream;
        Exec ex = rank_exec(dev, s, opts);
        StageTimer tm(s);
        DevBuf<double> pos((size_t)3 * n, s), mass, h, tgt, dpot, dacc;
        if (src_mass) mass.alloc((size_t)n, s);
        if (src_h) h.alloc((size_t)n, s);
        const int64_t lo = sb[(size_t)r], cnt = sb[(size_t)r + 1] - lo;
        if (cnt) {
            copy_h2d(pos.p + 3 * lo, src_pos + 3 * lo, (size_t)cnt * 3 * sizeof(double), ex);
            if (src_mass) copy_h2d(mass.p + lo, src_mass + lo, (size_t)cnt * sizeof(double), ex);
            if (src_h) copy_h2d(h.p + lo, src_h + lo, (size_t)cnt * sizeof(double), ex);
        }
        PNBX_CUDA(cudaEventRecord(c.ev, s));
        p_pos[(size_t)r] = pos.p; p_mass[(size_t)r] = mass.p; p_h[(size_t)r] = h.p; ready[(size_t)r] = c.ev;
        bar.arrive_and_wait();
        pull_shards(r, devs, p_pos, ready, sb, 3, s);
        if (src_mass) pull_shards(r, devs, p_mass, ready, sb, 1, s);
        if (src_h) pull_shards(r, devs, p_h, ready, sb, 1, s);
        const int64_t tlo = tb[(size_t)r], tcnt = tb[(size_t)r + 1] - tlo;
        if (tcnt) {
            if (!self) {
                tgt.alloc((size_t)3 * tcnt, s);
                copy_h2d(tgt.p, tgt_pos + 3 * tlo, (size_t)tcnt * 3 * sizeof(double), ex);
            }
            if (want & PNBX_WANT_POT) dpot.alloc((size_t)tcnt, s);
            if (want & PNBX_WANT_ACC) dacc.alloc((size_t)3 * tcnt, s);
            direct_on_device(ex, pos.p, mass.p, h.p, n, self ? nullptr : tgt.p, tcnt, self ? tlo : 0, kernel, want, dpot.p,
                             dacc.p, tm);
            if (want & PNBX_WANT_POT) copy_d2h(out_pot + tlo, dpot.p, (size_t)tcnt * sizeof(double), ex);
            if (want & PNBX_WANT_ACC) copy_d2h(out_acc + 3 * tlo, dacc.p, (size_t)tcnt * 3 * sizeof(double), ex);
        }
        PNBX_CUDA(cudaStreamSynchronize(s));
        bar.arrive_and_wait();  // nobody frees its shard while a peer may still be pulling from it
    });
    return true;
}

// ------------------------------------------------------------------------------------------ tree
bool multi_tree_create(pnbx_tree_impl& primary, const double* pos, const double* mass, const double* h, int64_t n,
                       int64_t leaf_capacity, int multipole_order, int kernel, const pnbx_opts* opts) {
    std::vector<int> devs = multi_devices();
    if (devs.size() < 2) return false;
    if ((double)n < env_number("PNBX_MULTI_MIN_N", 2e6)) return false;
    const int W = (int)devs.size();
    std::lock_guard<std::mutex> lock(g_multi_mu);
    primary.replicas.clear();
    for (int r = 1; r < W; ++r) primary.replicas.emplace_back(new pnbx_tree_impl());
    auto part = [&](int r) -> pnbx_tree_impl& { return r == 0 ? primary : *primary.replicas[(size_t)r - 1]; };
    const std::vector<int64_t> sb = even_bounds(n, W);
    std::vector<double*> p_pos((size_t)W, nullptr), p_mass((size_t)W, nullptr), p_h((size_t)W, nullptr);
    std::vector<cudaEvent_t> ready((size_t)W, nullptr);
    HostBarrier bar(W);
    try {
        run_ranks(devs, bar, [&](int r) {
            const int dev = devs[(size_t)r];
            rank_ctx(dev, devs);
            pnbx_tree_impl& t = part(r);
            tree_init(t, dev, n, leaf_capacity, multipole_order, kernel, mass != nullptr, h != nullptr);
            Exec ex = rank_exec(dev, t.stream, opts);
            StageTimer tm(t.stream);
            const int64_t lo = sb[(size_t)r], cnt = sb[(size_t)r + 1] - lo;
            tm.begin("octree.copy_in");
            if (cnt) {
                copy_h2d(t.pos.p + 3 * lo, pos + 3 * lo, (size_t)cnt * 3 * sizeof(double), ex);
                if (mass) copy_h2d(t.mass.p + lo, mass + lo, (size_t)cnt * sizeof(double), ex);
                if (h) copy_h2d(t.h.p + lo, h + lo, (size_t)cnt * sizeof(double), ex);
            }
            tree_mark_ready(t);  // "my shard is in place"
            p_pos[(size_t)r] = t.pos.p; p_mass[(size_t)r] = t.mass.p; p_h[(size_t)r] = t.h.p; ready[(size_t)r] = t.ready;
            bar.arrive_and_wait();
            pull_shards(r, devs, p_pos, ready, sb, 3, t.stream);
            if (mass) pull_shards(r, devs, p_mass, ready, sb, 1, t.stream);
            if (h) pull_shards(r, devs, p_h, ready, sb, 1, t.stream);
            tm.end();
            bar.arrive_and_wait();  // every pull is queued before anybody re-records its `ready` event
            tree_build(t, tm);
            tree_mark_ready(t);
            PNBX_CUDA(cudaStreamSynchronize(t.stream));
        });
    } catch (...) {
        primary.replicas.clear();
        throw;
    }
    primary.multi_devs = devs;
    return true;
}

// Replays a setter on every replica (each on its own device and stream).
void multi_tree_for_each(pnbx_tree_impl& primary, const std::function<void(pnbx_tree_impl&)>& fn) {
    std::lock_guard<std::mutex> lock(g_multi_mu);
    HostBarrier bar((int)primary.multi_devs.size());
    run_ranks(primary.multi_devs, bar, [&](int r) { fn(r == 0 ? primary : *primary.replicas[(size_t)r - 1]); });
}

bool multi_tree_eval(pnbx_tree_impl& primary, const double* tgt_pos, int64_t m, double theta, int want, double* out_pot,
                     double* out_acc, const pnbx_opts* opts) {
    const std::vector<int>& devs = primary.multi_devs;
    const int W = (int)devs.size();
    if (W < 2) return false;
    const bool self = tgt_pos == nullptr;
    if (!self && (double)m < env_number("PNBX_MULTI_MIN_TARGETS", 65536)) return false;
    std::lock_guard<std::mutex> lock(g_multi_mu);
    auto part = [&](int r) -> pnbx_tree_impl& { return r == 0 ? primary : *primary.replicas[(size_t)r - 1]; };
    const std::vector<int64_t> ob = even_bounds(m, W);  // slice of the caller's arrays that rank r returns
    OutSlices slices{};
    slices.n = W;
    for (int r = 0; r <= W; ++r) slices.bounds[r] = ob[(size_t)r];
    std::vector<cudaEvent_t> alloc_ready((size_t)W, nullptr), walk_done((size_t)W, nullptr);
    HostBarrier bar(W);
    run_ranks(devs, bar, [&](int r) {
        const int dev = devs[(size_t)r];
        RankCtx& c = rank_ctx(dev, devs);
        cudaStream_t s = c.stream;
        pnbx_tree_impl& t = part(r);
        Exec ex = rank_exec(dev, s, opts);
        StageTimer tm(s);
        const int64_t lo = ob[(size_t)r], cnt = ob[(size_t)r + 1] - lo;
        DevBuf<double> dpot, dacc, tgt;
        if (want & PNBX_WANT_POT) dpot.alloc((size_t)std::max<int64_t>(cnt, 1), s);
        if (want & PNBX_WANT_ACC) dacc.alloc((size_t)std::max<int64_t>(3 * cnt, 1), s);
        if (!self) {  // every device holds all query points: it walks a balanced share of their path-key order
            tgt.alloc((size_t)3 * m, s);
            copy_h2d(tgt.p, tgt_pos, (size_t)m * 3 * sizeof(double), ex);
        }
        tree_begin_use(t, s);
        // results go straight to the device that owns the particle's / point's ORIGINAL index (peer stores)
        slices.pot[r] = dpot.p;
        slices.acc[r] = dacc.p;
        PNBX_CUDA(cudaEventRecord(c.ev, s));
        alloc_ready[(size_t)r] = c.ev;
        bar.arrive_and_wait();
        for (int p = 0; p < W; ++p)
            if (p != r) PNBX_CUDA(cudaStreamWaitEvent(s, alloc_ready[(size_t)p], 0));  // peers' slices exist
        ex.block_cyclic = true;
        ex.tree_order = self;
        ex.shard_rank = r; ex.shard_world = W;
        ex.shard_block = self ? 4096 : 256;  // tree-order particles / path-key-ordered query points
        const int64_t mr = pnbx_shard_count(m, ex.shard_block, W, r);
        const OutSlices sl = slices;  // complete after the barrier
        if (mr > 0) tree_walk(t, ex, self ? nullptr : tgt.p, self ? mr : m, 0, theta, want, nullptr, nullptr, tm, nullptr, &sl);
        PNBX_CUDA(cudaEventRecord(c.ev2, s));
        walk_done[(size_t)r] = c.ev2;
        bar.arrive_and_wait();
        for (int p = 0; p < W; ++p)
            if (p != r) PNBX_CUDA(cudaStreamWaitEvent(s, walk_done[(size_t)p], 0));  // everybody stored into my slice
        tree_end_use(t, s);
        if (cnt) {
            if (want & PNBX_WANT_POT) copy_d2h(out_pot + lo, dpot.p, (size_t)cnt * sizeof(double), ex);
            if (want & PNBX_WANT_ACC) copy_d2h(out_acc + 3 * lo, dacc.p, (size_t)cnt * 3 * sizeof(double), ex);
        }
        PNBX_CUDA(cudaStreamSynchronize(s));
        bar.arrive_and_wait();  // slices stay allocated until every peer's stores have landed and been copied out
    });
    return true;
}

}  // namespace pnbx
