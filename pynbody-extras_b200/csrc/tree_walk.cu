// tree_walk.cu — K7: warp-cooperative, lane-exact tree walk (tree.rs:1069-1370, leaf sums :97-417).
//
// One warp walks 32 key-adjacent targets through the *union* of their reference paths with the
// reference's stackless links (first_subnode / next_branch). Every node record is fetched once per
// warp (all lanes read the same address: one broadcast transaction), each lane evaluates its OWN
// opening criterion in float64 on the float64 COM — `size2 < theta^2 * dist2` plus the hmax
// softening gate (tree.rs:55-71, 1117-1126) — so each target's interaction list is exactly the
// reference's (leaves are always summed particle by particle). A lane that accepts a node
// which other lanes must open is parked until the warp reaches that node's next_branch.
// Interactions (leaf pairs, multipole M2P) run in fp32 (or fp64 in verification mode) and are folded
// into float64 accumulators once per node.
#include <cub/cub.cuh>

#include "multipole.cuh"
#include "spline.cuh"
#include "tree.cuh"

namespace pnbx {
namespace {

#ifndef PNBX_WT
#define PNBX_WT 128
#endif
#ifndef PNBX_WALK_MINB
#define PNBX_WALK_MINB 10  // 48 registers: 40 of 64 warps resident. With the record loads issued together (load_rec) this
                           // beats 12 blocks / 40 registers (which then spills) and 8 / 63: 38.7 / 38.2 ms vs 42.1 / 38.9 and
                           // 40.3 / 40.8 (potentials / accelerations, N = 1e7, profiles/r02_walk_kernel_ncu.md)
#endif
constexpr int WT = PNBX_WT;  // threads per block (independent warps)
#ifndef PNBX_LEAF_UNROLL
#define PNBX_LEAF_UNROLL 2  // measured: 2 and 4 ~2 % ahead of 1 once sibling leaves are merged into runs
#endif
constexpr int LEAF_UNROLL = PNBX_LEAF_UNROLL;  // pass-1 leaf loop unroll (leaves hold <= leaf_capacity particles)

template <class T>
struct Vec4T {
    T x, y, z, w;
};

constexpr int WALK_VISIT_COST = 16;  // hybrid split: one node visit of a warp ~ 16 pair-loop iterations

template <class T>
struct WalkArgs {
    const NodeRec* rec;
    int gated;            // hmax payload present (tree.rs:57-59)
    unsigned long long* counters;  // counting pass only: visits, accepts, leaf visits, leaf particles, warp visits
    const T* moments;     // (nn, K): float64 coefficients (T=double) or fp32 walk records (T=float)
    int K;
    const void* src;      // float4 (T=float) or double spos/smass (T=double)
    const double* spos;   // sorted float64 positions
    const double* smass;  // sorted float64 masses (nullable -> 1)
    const T* src_h;       // sorted softenings (nullable): float64 as given (T=double), max(h,0)^2 in float32 (T=float)
    const double* sh;     // sorted softenings float64 (nullable)
    const uint32_t* perm; // sorted position -> original index
    // targets
    int64_t m;
    int self;                 // 1: targets are tree particles
    const uint32_t* tlist;    // self: sorted positions to evaluate (nullable = identity)
    int64_t tgt_begin;        // self: output slot = perm[s] - tgt_begin
    int tree_order;           // self: targets are sorted positions [tgt_begin, tgt_begin+m), output slot k
    int64_t cyc_block;        // > 0: block-cyclic tree-order shard: position of target k =
    int cyc_rank, cyc_world;  //      ((k / B) * world + rank) * B + k % B
    const double* tgt;        // points: (m,3) float64
    const uint32_t* torder;   // points: walk order -> point index
    double theta2;
    const double* rc;         // root cube {cx,cy,cz,half} on the device (origin of the float64-mode coordinates)
    int kernel;               // PNBX_KERNEL_PLUMMER | PNBX_KERNEL_SPLINE
    double* out_pot;
    double* out_acc;
    OutSlices slices;         // multi-device self evaluation: results go to the owner of the ORIGINAL index (peer stores)
    // hybrid evaluation of query points (tree_walk): a warp of the lane-per-target kernel whose walk outgrows `budget`
    // (node visits x WALK_VISIT_COST + particles of the leaves it loops over) gives up, appends its points to
    // `over_list` and leaves them to the warp-per-target kernel, which reads its target count from the device
    int budget;               // INT_MAX: no limit
    uint32_t* over_list;      // points handed over (any order), *n_over of them
    int* n_over;
    const int* n_front;       // warp-per-target kernel, nullable: number of valid entries of torder (else m)
};

// One node record = four 16-byte loads issued TOGETHER at the top of a visit. Left to the compiler, the loads of the
// fields a visit needs later (the COM for the opening test) were sunk below the `kind` dispatch and the warp paid two
// dependent memory latencies per visit (ncu source page: 10 % + 12 % of the stall samples on the two first uses).
__device__ __forceinline__ NodeRec load_rec(const NodeRec* __restrict__ p) {
    uint4 q0, q1, q2, q3;
    const uint4* r = reinterpret_cast<const uint4*>(p);
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w) : "l"(r));
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w) : "l"(r + 1));
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q2.x), "=r"(q2.y), "=r"(q2.z), "=r"(q2.w) : "l"(r + 2));
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q3.x), "=r"(q3.y), "=r"(q3.z), "=r"(q3.w) : "l"(r + 3));
    NodeRec c;
    c.com[0] = __hiloint2double((int)q0.y, (int)q0.x);
    c.com[1] = __hiloint2double((int)q0.w, (int)q0.z);
    c.com[2] = __hiloint2double((int)q1.y, (int)q1.x);
    c.size2 = __hiloint2double((int)q1.w, (int)q1.z);
    c.gate2 = __hiloint2double((int)q2.y, (int)q2.x);
    c.next_branch = (int)q2.z;
    c.first = (int)q2.w;
    c.kind = (int)q3.x;
    c.nleaf = (int)q3.y;
    c.ref = (int)q3.z;
    c.pad_ = 0;
    return c;
}

template <class T>
__device__ __forceinline__ T tiny_v();
template <>
__device__ __forceinline__ float tiny_v<float>() { return FLT_MIN; }
template <>
__device__ __forceinline__ double tiny_v<double>() { return DBL_MIN; }

template <class T>
__device__ __forceinline__ T w2_in(T u) {  // kernel.rs:84-102, u < 1
    T u2 = u * u;
    if (u < T(0.5)) return T(16.0 / 3.0) * u2 + u2 * u2 * (T(32.0 / 5.0) * u - T(48.0 / 5.0)) - T(14.0 / 5.0);
    return T(1.0 / 15.0) / u + u2 * (T(32.0 / 3.0) + u * (T(-16.0) + u * (T(48.0 / 5.0) - T(32.0 / 15.0) * u))) - T(16.0 / 5.0);
}
template <class T>
__device__ __forceinline__ T w2p_in(T u) {  // kernel.rs:108-124, u < 1
    T u2 = u * u;
    if (u < T(0.5)) return u * (T(32.0 / 3.0) + u2 * (T(32.0) * u - T(192.0 / 5.0)));
    return T(-1.0 / 15.0) / u2 + u * (T(64.0 / 3.0) + u * (T(-48.0) + u * (T(192.0 / 5.0) - T(32.0 / 3.0) * u)));
}

// ---- fp32 leaf sums (tree.rs:97-417). `sp` = the leaf's sources relative to the leaf's COM, `h2p` = their clamped
// softenings SQUARED (max(h,0)^2; h = max(h_source, h_target) <=> h^2 = max of the squares), (lx,ly,lz) = the target in
// the same frame, th2 = the target's clamped softening squared, skip_rel = index of the target itself inside this leaf
// (anything outside [0,n) if it is not here: skip_self by index, tree.rs:130).
// Pass 1, branch-free: every pair as Newtonian / Plummer. For the spline kernel, pairs with r < h contribute nothing
// here and are flagged through the return value (true: this lane has at least one such pair) ...
template <int WANT, int SMODE, bool CHECK>
__device__ __forceinline__ bool leaf_pass1_f32(const float4* __restrict__ sp, const float* __restrict__ h2p, int n,
                                               int skip_rel, float lx, float ly, float lz, float th2, float& pot,
                                               float& ax, float& ay, float& az) {
    float margin = 0.f;  // min over pairs of r^2 - h^2, clipped at 0
#pragma unroll LEAF_UNROLL
    for (int i = 0; i < n; ++i) {
        const float4 s = __ldg(sp + i);
        const float dx = s.x - lx, dy = s.y - ly, dz = s.z - lz;
        float r2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, FLT_MIN)));  // + R2_TINY (tree.rs:139)
        float m = s.w;
        if (CHECK && i == skip_rel) {  // the target itself: contributes exactly nothing
            m = 0.f;
            r2 = 1e30f;
        }
        if (SMODE != 0) {
            const float h2 = fmaxf(__ldg(h2p + i), th2);  // tree.rs:234-235
            if (SMODE == 2) {
                const float d = r2 - h2;  // < 0: inside the spline radius, left to pass 2 (no add-then-subtract)
                margin = fminf(margin, d);
                m = d < 0.f ? 0.f : m;
            } else {
                r2 += h2;  // Plummer: -1/sqrt(r^2+h^2); h = 0 is Newtonian
            }
        }
        const float rinv = mp::inv_sqrt<float>(r2);
        const float mr = m * rinv;
        if (WANT & PNBX_WANT_POT) pot -= mr;
        if (WANT & PNBX_WANT_ACC) {
            const float mg = mr * (rinv * rinv);
            ax = fmaf(dx, mg, ax);
            ay = fmaf(dy, mg, ay);
            az = fmaf(dz, mg, az);
        }
    }
    return margin < 0.f;
}
// ... and pass 2 adds the W2-kernel terms of exactly those pairs (tree.rs:237-243 / 367-378), branch-free: all lanes
// that enter run every pair, pairs outside their radius select 0.
template <int WANT>
__device__ __forceinline__ void leaf_pass2_f32(const float4* __restrict__ sp, const float* __restrict__ h2p, int n,
                                               int skip_rel, float lx, float ly, float lz, float th2, float& pot,
                                               float& ax, float& ay, float& az) {
#pragma unroll 2
    for (int i = 0; i < n; ++i) {
        const float4 s = __ldg(sp + i);
        const float h2 = fmaxf(__ldg(h2p + i), th2);
        const float dx = s.x - lx, dy = s.y - ly, dz = s.z - lz;
        const float r2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, FLT_MIN)));
        const float d = r2 - h2;  // the same expression as pass 1: the two passes partition the pairs
        const bool in = (d < 0.f) && (i != skip_rel);
        const float rinv = mp::inv_sqrt<float>(r2), hinv = mp::inv_sqrt<float>(h2);
        float kpot, kacc;
        w2_terms_f32(r2, rinv, h2 * hinv, hinv, kpot, kacc);
        if (WANT & PNBX_WANT_POT) pot = fmaf(s.w, in ? kpot : 0.f, pot);
        if (WANT & PNBX_WANT_ACC) {
            const float mg = s.w * (in ? kacc : 0.f);
            ax = fmaf(dx, mg, ax);
            ay = fmaf(dy, mg, ay);
            az = fmaf(dz, mg, az);
        }
    }
}

// source particle p in T: float relative to its leaf's COM, double relative to the root centre
template <class T>
__device__ __forceinline__ Vec4T<T> load_src(const WalkArgs<T>& a, int p);
template <>
__device__ __forceinline__ Vec4T<float> load_src<float>(const WalkArgs<float>& a, int p) {
    float4 v = __ldg(reinterpret_cast<const float4*>(a.src) + p);
    return {v.x, v.y, v.z, v.w};
}
template <>
__device__ __forceinline__ Vec4T<double> load_src<double>(const WalkArgs<double>& a, int p) {
    return {a.spos[3 * (int64_t)p] - a.rc[0], a.spos[3 * (int64_t)p + 1] - a.rc[1], a.spos[3 * (int64_t)p + 2] - a.rc[2],
            a.smass ? a.smass[p] : 1.0};
}

// SMODE: 0 = no softening anywhere (no per-particle h, no hmax gate), 1 = Plummer, 2 = cubic spline,
//        3 = decided at run time (float64 verification mode and the counting pass).
// WANT == 0 is the counting pass: traversal decisions only, totals into a.counters.
template <int ORDER, int WANT, class T, int SMODE, bool BUDGET = false>
__global__ void __launch_bounds__(WT, (ORDER >= 4 || sizeof(T) == 8) ? 4 : PNBX_WALK_MINB) walk_kernel(const WalkArgs<T> a) {
    constexpr int DORD = (WANT & PNBX_WANT_ACC) ? (ORDER < 1 ? 1 : ORDER) : (ORDER < 2 ? 0 : ORDER);
    constexpr unsigned FULL = 0xffffffffu;
    const int64_t k = (int64_t)blockIdx.x * WT + threadIdx.x;
    const bool valid = k < a.m;
    // ---- this lane's target
    double tx = 0, ty = 0, tz = 0, th64 = 0;
    int skip = -1;
    int64_t oslot = 0;
    bool has_th = false;
    if (valid) {
        if (a.self) {
            uint32_t s;
            if (a.cyc_block > 0) s = (uint32_t)(((k / a.cyc_block) * a.cyc_world + a.cyc_rank) * a.cyc_block + k % a.cyc_block);
            else s = a.tree_order ? (uint32_t)(a.tgt_begin + k) : a.tlist ? a.tlist[k] : (uint32_t)k;
            tx = a.spos[3 * (int64_t)s]; ty = a.spos[3 * (int64_t)s + 1]; tz = a.spos[3 * (int64_t)s + 2];
            skip = (int)s;
            oslot = a.tree_order ? k : (int64_t)a.perm[s] - a.tgt_begin;
            if (a.sh) { th64 = a.sh[s]; has_th = true; }  // target_h_opt = softenings[i] (tree.rs:1439)
        } else {
            // walk order = path-key order of the query points; a multi-device call takes a block-cyclic share of it
            const int64_t sp = a.cyc_block > 0 ? ((k / a.cyc_block) * a.cyc_world + a.cyc_rank) * a.cyc_block + k % a.cyc_block : k;
            const uint32_t q = a.torder[sp];
            tx = a.tgt[3 * (int64_t)q]; ty = a.tgt[3 * (int64_t)q + 1]; tz = a.tgt[3 * (int64_t)q + 2];
            oslot = q;
        }
    }
    const T fx = (T)(tx - a.rc[0]), fy = (T)(ty - a.rc[1]), fz = (T)(tz - a.rc[2]);
    const T th = (T)fmax(th64, 0.0);  // clamped target softening (tree.rs:115)
    const float th2 = (float)th * (float)th;  // fp32 leaf sums compare squared softenings
    const bool soft = SMODE == 0 ? false : SMODE == 3 ? a.src_h != nullptr : true;  // softenings_opt.is_some() (tree.rs:116)
    const bool spline = SMODE == 3 ? a.kernel == PNBX_KERNEL_SPLINE : SMODE == 2;
    const bool gated = SMODE == 0 ? false : a.gated != 0;
    // softening gate threshold of this target: (c * max(h_t, 0))^2, c = 2.8 / 1.0 (kernel.rs:20-28). The node's
    // own (c * max(hmax,0))^2 is precomputed; max() commutes with the monotone rounded ops, so
    // max(gate_node, gate_t) == (c*h)*(c*h) for h = max(hmax, h_t) bit for bit (tree.rs:61-70).
    double gate_t = 0.0;
    if (gated && has_th) {
        const double ch = __dmul_rn(spline ? 1.0 : 2.8, fmax(th64, 0.0));
        gate_t = __dmul_rn(ch, ch);
    }
    double P = 0.0, Ax = 0.0, Ay = 0.0, Az = 0.0;
    long long n_visit = 0, n_accept = 0, n_leaf = 0, n_leafp = 0, n_wvisit = 0;
    int cost = 0;  // warp-uniform

    bool active = valid;
    int resume = INT_MIN;
    int idx = __ballot_sync(FULL, valid) ? 0 : -1;
    while (idx >= 0) {
        const NodeRec c = load_rec(a.rec + idx);  // one 64-byte record: one memory round trip per visit
        const NodeRec& gm = c;
        if (!active && resume == idx) active = true;
        if constexpr (BUDGET) {  // separate instantiation: the self evaluations' loop does not carry the counter
            cost += c.kind >= 0 ? WALK_VISIT_COST + c.kind : WALK_VISIT_COST;
            if (cost > a.budget) {  // hand this warp's points over to the warp-per-target kernel (query points only)
                const unsigned vm = __ballot_sync(FULL, valid);
                const int lane = threadIdx.x & 31;
                int base = 0;
                if (lane == 0) base = atomicAdd(a.n_over, __popc(vm));
                base = __shfl_sync(FULL, base, 0);
                if (valid) a.over_list[base + __popc(vm & ((1u << lane) - 1u))] = (uint32_t)oslot;
                return;
            }
        }
        if (WANT == 0) {  // a leaf-run record stands for `nleaf` reference nodes (tree.cuh)
            const int nodes = c.kind >= 0 ? c.nleaf : 1;
            if (active) n_visit += nodes;
            if ((threadIdx.x & 31) == 0) n_wvisit += nodes;  // nodes the WARP visits: union of its lanes' paths
        }
        if (c.kind == -2) {  // zero mass: skip the subtree (tree.rs:1087-1090)
            idx = c.next_branch;
            continue;
        }
        if (c.kind >= 0) {  // leaf (or a run of sibling leaves, tree.cuh): always summed directly (tree.rs:1094-1112)
            if (WANT == 0) {
                if (active) { n_leaf += c.nleaf; n_leafp += c.kind; }
            } else if (active) {
                if constexpr (sizeof(T) == 4) {
                    // fp32: sources are stored relative to their leaf run's origin (close pairs keep their separation)
                    float pot = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
                    const float lx = (float)(tx - gm.com[0]), ly = (float)(ty - gm.com[1]), lz = (float)(tz - gm.com[2]);
                    const float4* sp = reinterpret_cast<const float4*>(a.src) + c.first;
                    const float* h2p = SMODE != 0 ? a.src_h + c.first : nullptr;
                    const int skip_rel = skip - c.first;
                    bool inside;
                    if ((unsigned)skip_rel < (unsigned)c.kind)  // the target's own leaf: the loop with the index check
                        inside = leaf_pass1_f32<WANT, SMODE, true>(sp, h2p, c.kind, skip_rel, lx, ly, lz, th2, pot, ax, ay, az);
                    else
                        inside = leaf_pass1_f32<WANT, SMODE, false>(sp, h2p, c.kind, skip_rel, lx, ly, lz, th2, pot, ax, ay, az);
                    if (SMODE == 2 && inside) leaf_pass2_f32<WANT>(sp, h2p, c.kind, skip_rel, lx, ly, lz, th2, pot, ax, ay, az);
                    if (WANT & PNBX_WANT_POT) P += (double)pot;
                    if (WANT & PNBX_WANT_ACC) { Ax += (double)ax; Ay += (double)ay; Az += (double)az; }
                } else {
                    // float64 verification mode: sources relative to the root centre, runtime softening mode
                    T pot = T(0), ax = T(0), ay = T(0), az = T(0);
                    const T lx = fx, ly = fy, lz = fz;
                    bool inside = false;
                    for (int p = c.first; p < c.first + c.kind; ++p) {
                        Vec4T<T> s = load_src<T>(a, p);
                        const T dx = s.x - lx, dy = s.y - ly, dz = s.z - lz;
                        T r2 = fma(dx, dx, fma(dy, dy, dz * dz));
                        if (p == skip) {  // skip_self by index (tree.rs:130): contributes exactly nothing
                            s.w = T(0);
                            r2 = T(1);
                        }
                        if (soft) {
                            const T h = max(max(a.src_h[p], T(0)), th);  // tree.rs:234-235
                            if (spline) {
                                const bool in = r2 < h * h;  // h > 0 is implied by r2 >= 0
                                inside |= in;
                                if (in) s.w = T(0);          // left to pass 2 (no add-then-subtract cancellation)
                            } else {
                                r2 = fma(h, h, r2);          // Plummer: -1/sqrt(r^2+h^2); h = 0 is Newtonian
                            }
                        }
                        const T rinv = mp::inv_sqrt<T>(r2 + tiny_v<T>());
                        const T mr = s.w * rinv;
                        if (WANT & PNBX_WANT_POT) pot -= mr;
                        if (WANT & PNBX_WANT_ACC) {
                            const T mg = mr * (rinv * rinv);
                            ax = fma(dx, mg, ax);
                            ay = fma(dy, mg, ay);
                            az = fma(dz, mg, az);
                        }
                    }
                    if (spline && inside) {  // pass 2: the W2-kernel terms of the pairs with r < h (tree.rs:237-243 / 367-378)
                        for (int p = c.first; p < c.first + c.kind; ++p) {
                            if (p == skip) continue;
                            const Vec4T<T> s = load_src<T>(a, p);
                            const T h = max(max(a.src_h[p], T(0)), th);
                            const T dx = s.x - lx, dy = s.y - ly, dz = s.z - lz;
                            const T r2 = fma(dx, dx, fma(dy, dy, dz * dz));
                            if (!(r2 < h * h)) continue;
                            const T rinv = mp::inv_sqrt<T>(r2 + tiny_v<T>());
                            const T hinv = T(1) / h;
                            const T u = (r2 + tiny_v<T>()) * rinv * hinv;
                            if (WANT & PNBX_WANT_POT) pot = fma(s.w, w2_in(u) * hinv, pot);
                            if (WANT & PNBX_WANT_ACC) {
                                const T mg = s.w * (w2p_in(u) * (hinv * hinv) * rinv);
                                ax = fma(dx, mg, ax);
                                ay = fma(dy, mg, ay);
                                az = fma(dz, mg, az);
                            }
                        }
                    }
                    if (WANT & PNBX_WANT_POT) P += (double)pot;
                    if (WANT & PNBX_WANT_ACC) { Ax += (double)ax; Ay += (double)ay; Az += (double)az; }
                }
            }
            idx = c.next_branch;
            continue;
        }
        // ---- internal node: per-lane opening decision in float64 (tree.rs:1114-1126)
        // evaluated by every lane without a branch (parked lanes compute a value nobody uses: free under SIMT)
        const double dx = gm.com[0] - tx, dy = gm.com[1] - ty, dz = gm.com[2] - tz;
        // (+ R2_TINY of tree.rs:1117 can only matter for dist2 < 1e-300, where nothing is accepted anyway)
        const double dist2 = fma(dx, dx, fma(dy, dy, __dmul_rn(dz, dz)));
        bool accept = gm.size2 < __dmul_rn(a.theta2, dist2);
        // node_soft_ok (tree.rs:55-71): dist2 > max(gates). Without an hmax payload both gates are 0 (records and gate_t),
        // and dist2 > 0 can only fail where the opening test fails too, so softened kernels test unconditionally.
        if (SMODE == 3 ? gated : SMODE != 0) accept = accept && dist2 > c.gate2 && dist2 > gate_t;
        accept = accept && active;
        const bool need_open = __any_sync(FULL, active && !accept);
        if (WANT == 0) {
            if (accept) ++n_accept;
        } else if (accept) {
            if (sizeof(T) == 4) {
                // fp32: contracted closed forms on the per-node fp32 record (multipole.cuh m2p_fast / m2p_fast45)
                float pot = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
                const float* rec = reinterpret_cast<const float*>(a.moments) + (int64_t)idx * mp::fast_rec_floats(ORDER);
                if (ORDER <= 3) mp::m2p_fast<(ORDER <= 3 ? ORDER : 3), (WANT == 0 ? 1 : WANT)>(rec, (float)dx, (float)dy, (float)dz, pot, ax, ay, az);
                else mp::m2p_fast45<(ORDER >= 4 ? ORDER : 4), (WANT == 0 ? 1 : WANT)>(rec, (float)dx, (float)dy, (float)dz, pot, ax, ay, az);
                if (WANT & PNBX_WANT_POT) P += (double)pot;
                if (WANT & PNBX_WANT_ACC) { Ax += (double)ax; Ay += (double)ay; Az += (double)az; }
            } else {
                const T* M = a.moments + (int64_t)c.ref * a.K;  // float64 payloads are in the reference numbering
                T D[mp::NCOEF];
                mp::derivatives<DORD, T>((T)dx, (T)dy, (T)dz, tiny_v<T>(), D);
                if (ORDER <= 1) {
                    const T m0 = M[0];
                    if (WANT & PNBX_WANT_POT) P += (double)(-m0 * D[mp::I000]);
                    if (WANT & PNBX_WANT_ACC) {
                        Ax += (double)(-m0 * D[mp::I100]);
                        Ay += (double)(-m0 * D[mp::I010]);
                        Az += (double)(-m0 * D[mp::I001]);
                    }
                } else {
                    T Mr[mp::stored_coeffs(ORDER)];
#pragma unroll
                    for (int i = 0; i < mp::stored_coeffs(ORDER); ++i) Mr[i] = M[i];
                    if (WANT & PNBX_WANT_POT) P += (double)mp::m2p_potential<ORDER, T>(Mr, D);
                    if (WANT & PNBX_WANT_ACC) {
                        T ax, ay, az;
                        mp::m2p_accel<ORDER, T>(Mr, D, ax, ay, az);
                        Ax += (double)ax; Ay += (double)ay; Az += (double)az;
                    }
                }
            }
        }
        if (need_open) {
            if (accept) {  // done with this subtree; wait for the warp at its next_branch
                active = false;
                resume = c.next_branch;
            }
            idx = c.first;
        } else {
            idx = c.next_branch;
        }
    }
    if (WANT == 0) {  // warp-reduce, one atomic per warp per counter (integers: order-independent)
        for (int o = 16; o > 0; o >>= 1) {
            n_visit += __shfl_down_sync(FULL, n_visit, o);
            n_accept += __shfl_down_sync(FULL, n_accept, o);
            n_leaf += __shfl_down_sync(FULL, n_leaf, o);
            n_leafp += __shfl_down_sync(FULL, n_leafp, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(a.counters + 0, (unsigned long long)n_visit);
            atomicAdd(a.counters + 1, (unsigned long long)n_accept);
            atomicAdd(a.counters + 2, (unsigned long long)n_leaf);
            atomicAdd(a.counters + 3, (unsigned long long)n_leafp);
            atomicAdd(a.counters + 4, (unsigned long long)n_wvisit);
        }
        return;
    }
    if (valid) {
        double* op = a.out_pot;
        double* oa = a.out_acc;
        if (a.slices.n > 0) {  // the slice (a peer GPU's memory, or ours) that owns this particle's original index
            const int64_t gi = a.self ? (int64_t)a.perm[skip] : oslot;  // original particle index / query point index
            int o = 0;
            while (o + 1 < a.slices.n && gi >= a.slices.bounds[o + 1]) ++o;
            oslot = gi - a.slices.bounds[o];
            op = a.slices.pot[o];
            oa = a.slices.acc[o];
        }
        if (WANT & PNBX_WANT_POT) op[oslot] = P;
        if (WANT & PNBX_WANT_ACC) {
            oa[3 * oslot] = Ax; oa[3 * oslot + 1] = Ay; oa[3 * oslot + 2] = Az;
        }
    }
}

template <class T, int SMODE, bool BUDGET = false>
void launch_walk(int order, int want, const WalkArgs<T>& a, cudaStream_t s) {
    const unsigned grid = (unsigned)std::max<int64_t>(1, ceil_div(a.m, WT));
#define PNBX_W(O, W)                                                        \
    if (order == O && want == W) {                                          \
        PNBX_LAUNCH((walk_kernel<O, W, T, SMODE, BUDGET>), grid, WT, 0, s, a); \
        return;                                                             \
    }
    PNBX_W(1, 1) PNBX_W(1, 2) PNBX_W(1, 3)
    PNBX_W(2, 1) PNBX_W(2, 2) PNBX_W(2, 3)
    PNBX_W(3, 1) PNBX_W(3, 2) PNBX_W(3, 3)
    PNBX_W(4, 1) PNBX_W(4, 2) PNBX_W(4, 3)
    PNBX_W(5, 1) PNBX_W(5, 2) PNBX_W(5, 3)
#undef PNBX_W
    throw ArgError{PNBX_ERR_ARG, "internal: no walk kernel variant"};
}

// ------------------------------------------------------------------------------------------------------------
// Warp-per-target walk (fp32 arithmetic) for calls with FEW targets: query grids (rotation curves, binding-energy
// profiles; BASELINE config 5) hand the lane-per-target kernel a few thousand warps whose cost is set by the heaviest
// ones — a target in a softened core sums thousands of leaf particles serially while most of the GPU idles. Here one
// warp owns one target: the traversal (same records, same float64 opening test and softening gate, so the
// interaction list is again exactly the reference's) is warp-uniform, and the WORK it generates is spread over the
// lanes through two 32-entry queues — accepted nodes (one M2P per lane) and leaf particles (one pair per lane) —
// that are drained whenever they fill. Every lane folds its fp32 terms into float64 accumulators; one fixed-order
// butterfly at the end adds the 32 partials. The result is a pure function of (tree, target).
template <int ORDER, int WANT, int SMODE>
__global__ void __launch_bounds__(WT) walk_wpt_kernel(const WalkArgs<float> a) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t k = ((int64_t)blockIdx.x * WT + threadIdx.x) >> 5;  // one target per warp
    if (k >= (a.n_front ? (int64_t)*a.n_front : a.m)) return;  // whole warp
    double tx, ty, tz, th64 = 0.0;
    int skip = -1;
    int64_t oslot;
    bool has_th = false;
    if (a.self) {
        uint32_t s;
        if (a.cyc_block > 0) s = (uint32_t)(((k / a.cyc_block) * a.cyc_world + a.cyc_rank) * a.cyc_block + k % a.cyc_block);
        else s = a.tree_order ? (uint32_t)(a.tgt_begin + k) : a.tlist ? a.tlist[k] : (uint32_t)k;
        tx = a.spos[3 * (int64_t)s]; ty = a.spos[3 * (int64_t)s + 1]; tz = a.spos[3 * (int64_t)s + 2];
        skip = (int)s;
        oslot = a.tree_order ? k : (int64_t)a.perm[s] - a.tgt_begin;
        if (a.sh) { th64 = a.sh[s]; has_th = true; }
    } else {
        const int64_t sp = a.cyc_block > 0 ? ((k / a.cyc_block) * a.cyc_world + a.cyc_rank) * a.cyc_block + k % a.cyc_block : k;
        const uint32_t q = a.torder[sp];
        tx = a.tgt[3 * (int64_t)q]; ty = a.tgt[3 * (int64_t)q + 1]; tz = a.tgt[3 * (int64_t)q + 2];
        oslot = q;
    }
    const float thc = (float)fmax(th64, 0.0);
    const float th2 = thc * thc;
    const bool spline = SMODE == 2;
    double gate_t = 0.0;
    if (SMODE != 0 && a.gated && has_th) {
        const double ch = __dmul_rn(spline ? 1.0 : 2.8, fmax(th64, 0.0));
        gate_t = __dmul_rn(ch, ch);
    }
    double P = 0.0, Ax = 0.0, Ay = 0.0, Az = 0.0;  // this lane's share
    // queues: one pending accepted node / one pending leaf particle per lane
    int qn_node = -1;
    float qn_dx = 0.f, qn_dy = 0.f, qn_dz = 0.f;
    int n_nodes = 0;
    int qp_idx = -1;
    float qp_lx = 0.f, qp_ly = 0.f, qp_lz = 0.f;
    int n_parts = 0;

    auto drain_nodes = [&]() {
        if (qn_node >= 0) {
            float pot = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
            const float* rec = reinterpret_cast<const float*>(a.moments) + (int64_t)qn_node * mp::fast_rec_floats(ORDER);
            if (ORDER <= 3) mp::m2p_fast<(ORDER <= 3 ? ORDER : 3), WANT>(rec, qn_dx, qn_dy, qn_dz, pot, ax, ay, az);
            else mp::m2p_fast45<(ORDER >= 4 ? ORDER : 4), WANT>(rec, qn_dx, qn_dy, qn_dz, pot, ax, ay, az);
            if (WANT & PNBX_WANT_POT) P += (double)pot;
            if (WANT & PNBX_WANT_ACC) { Ax += (double)ax; Ay += (double)ay; Az += (double)az; }
        }
        qn_node = -1;
        n_nodes = 0;
    };
    auto drain_parts = [&]() {
        if (qp_idx >= 0) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(a.src) + qp_idx);
            const float dx = s4.x - qp_lx, dy = s4.y - qp_ly, dz = s4.z - qp_lz;
            float r2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, FLT_MIN)));
            float m = s4.w;
            if (qp_idx == skip) {  // skip_self by index (tree.rs:130): contributes exactly nothing
                m = 0.f;
                r2 = 1e30f;
            }
            float kpot, kacc;
            bool inside = false;
            float h2 = 0.f;
            if (SMODE != 0) {
                h2 = fmaxf(__ldg(a.src_h + qp_idx), th2);
                if (SMODE == 2) inside = r2 < h2;  // the lane-per-target kernel's predicate (d = r2 - h2 < 0)
                else r2 += h2;
            }
            const float rinv = mp::inv_sqrt<float>(r2);
            kpot = -rinv;
            kacc = rinv * rinv * rinv;
            if (SMODE == 2 && inside) {
                const float hinv = mp::inv_sqrt<float>(h2);
                w2_terms_f32(r2, rinv, h2 * hinv, hinv, kpot, kacc);
            }
            if (WANT & PNBX_WANT_POT) P += (double)(m * kpot);
            if (WANT & PNBX_WANT_ACC) {
                const float g = m * kacc;
                Ax += (double)(dx * g); Ay += (double)(dy * g); Az += (double)(dz * g);
            }
        }
        qp_idx = -1;
        n_parts = 0;
    };

    int idx = 0;
    while (idx >= 0) {
        const NodeRec c = a.rec[idx];
        if (c.kind == -2) { idx = c.next_branch; continue; }
        if (c.kind >= 0) {  // leaf run: queue its particles, 32 at a time
            const float lx = (float)(tx - c.com[0]), ly = (float)(ty - c.com[1]), lz = (float)(tz - c.com[2]);
            int done = 0;
            while (done < c.kind) {
                const int room = 32 - n_parts;
                const int take = min(room, c.kind - done);
                if (lane >= n_parts && lane < n_parts + take) {
                    qp_idx = c.first + done + (lane - n_parts);
                    qp_lx = lx; qp_ly = ly; qp_lz = lz;
                }
                n_parts += take;
                done += take;
                if (n_parts == 32) drain_parts();
            }
            idx = c.next_branch;
            continue;
        }
        const double dx = c.com[0] - tx, dy = c.com[1] - ty, dz = c.com[2] - tz;
        const double dist2 = fma(dx, dx, fma(dy, dy, __dmul_rn(dz, dz)));
        bool accept = c.size2 < __dmul_rn(a.theta2, dist2);
        if (SMODE != 0) accept = accept && dist2 > c.gate2 && dist2 > gate_t;
        if (accept) {
            if (lane == n_nodes) { qn_node = idx; qn_dx = (float)dx; qn_dy = (float)dy; qn_dz = (float)dz; }
            if (++n_nodes == 32) drain_nodes();
            idx = c.next_branch;
        } else {
            idx = c.first;
        }
    }
    drain_nodes();
    drain_parts();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {  // fixed-order butterfly: deterministic
        if (WANT & PNBX_WANT_POT) P += __shfl_xor_sync(FULL, P, o);
        if (WANT & PNBX_WANT_ACC) {
            Ax += __shfl_xor_sync(FULL, Ax, o);
            Ay += __shfl_xor_sync(FULL, Ay, o);
            Az += __shfl_xor_sync(FULL, Az, o);
        }
    }
    if (lane == 0) {
        double* op = a.out_pot;
        double* oa = a.out_acc;
        if (a.slices.n > 0) {
            const int64_t gi = a.self ? (int64_t)a.perm[skip] : oslot;
            int o = 0;
            while (o + 1 < a.slices.n && gi >= a.slices.bounds[o + 1]) ++o;
            oslot = gi - a.slices.bounds[o];
            op = a.slices.pot[o];
            oa = a.slices.acc[o];
        }
        if (WANT & PNBX_WANT_POT) op[oslot] = P;
        if (WANT & PNBX_WANT_ACC) { oa[3 * oslot] = Ax; oa[3 * oslot + 1] = Ay; oa[3 * oslot + 2] = Az; }
    }
}

template <int SMODE>
void launch_walk_wpt(int order, int want, const WalkArgs<float>& a, cudaStream_t s) {
    const unsigned grid = (unsigned)std::max<int64_t>(1, ceil_div(a.m * 32, WT));
#define PNBX_W(O, W)                                                        \
    if (order == O && want == W) {                                          \
        PNBX_LAUNCH((walk_wpt_kernel<O, W, SMODE>), grid, WT, 0, s, a);     \
        return;                                                             \
    }
    PNBX_W(1, 1) PNBX_W(1, 2) PNBX_W(1, 3)
    PNBX_W(2, 1) PNBX_W(2, 2) PNBX_W(2, 3)
    PNBX_W(3, 1) PNBX_W(3, 2) PNBX_W(3, 3)
    PNBX_W(4, 1) PNBX_W(4, 2) PNBX_W(4, 3)
    PNBX_W(5, 1) PNBX_W(5, 2) PNBX_W(5, 3)
#undef PNBX_W
    throw ArgError{PNBX_ERR_ARG, "internal: no walk kernel variant"};
}

__global__ void point_keys(const double* __restrict__ pos, int64_t n, const double* __restrict__ root4,
                           uint64_t* __restrict__ key) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
    double cx = root4[0], cy = root4[1], cz = root4[2], hf = root4[3];
    uint64_t k = 0;
    for (int l = 1; l <= KEY_LEVELS_HI; ++l) {
        const unsigned ox = x >= cx, oy = y >= cy, oz = z >= cz;
        k |= (uint64_t)(ox | (oy << 1) | (oz << 2)) << (3 * (KEY_LEVELS_HI - l));
        const double off = hf * 0.5;
        cx += ox ? off : -off;
        cy += oy ? off : -off;
        cz += oz ? off : -off;
        hf = off;
    }
    key[i] = k;
}
__global__ void iota32(uint32_t* p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}
__global__ void shard_flags(const uint32_t* __restrict__ perm, int64_t n, int64_t lo, int64_t hi, uint8_t* __restrict__ f) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) f[s] = perm[s] >= lo && perm[s] < hi;
}

inline unsigned nb(int64_t n) { return (unsigned)std::max<int64_t>(1, ceil_div(n, 256)); }

}  // namespace

void tree_walk(const pnbx_tree_impl& t, const Exec& ex, const double* d_tgt, int64_t m, int64_t tgt_begin, double theta,
               int want, double* d_pot, double* d_acc, StageTimer& tm, unsigned long long* d_counters,
               const OutSlices* slices) {
    cudaStream_t s = ex.stream;
    const bool self = d_tgt == nullptr;
    // query points of a multi-device call: `m` points are sorted by path key on every device (identical order) and this
    // device walks its block-cyclic share of that order (balanced whatever the layout of the caller's array)
    const bool pts_cyclic = !self && ex.block_cyclic;
    const int64_t m_walk = pts_cyclic ? pnbx_shard_count(m, ex.shard_block, ex.shard_world, ex.shard_rank) : m;
    // kernel choice: self evaluations go by the size of the WHOLE call (a multi-device share then runs the kernel the
    // single-device call would run: bit-identical results); query points go by what THIS device walks — a grid spread
    // over 8 GPUs is 8 small, heavy-tailed calls, which is exactly what the warp-per-target kernel is for
    const int64_t wpt_basis = self ? (ex.block_cyclic ? t.n : m) : m_walk;
    DevBuf<uint32_t> tlist, torder;
    tm.begin("octree.walk.prepare_targets");
    const bool tree_order = self && ex.tree_order;
    if (self) {
        if (!tree_order && !(tgt_begin == 0 && m == t.n)) {  // shard: sorted positions whose particle lies in [tgt_begin, tgt_begin+m)
            DevBuf<uint8_t> flag((size_t)t.n, s);
            DevBuf<int32_t> nsel(1, s);
            tlist.alloc((size_t)t.n, s);
            PNBX_LAUNCH(shard_flags, nb(t.n), 256, 0, s, t.perm.p, t.n, tgt_begin, tgt_begin + m, flag.p);
            size_t bytes = 0;
            cub::CountingInputIterator<uint32_t> it(0);
            PNBX_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes, it, flag.p, tlist.p, nsel.p, (int)t.n, s));
            DevBuf<uint8_t> tmp(bytes, s);
            PNBX_CUDA(cub::DeviceSelect::Flagged(tmp.get(), bytes, it, flag.p, tlist.p, nsel.p, (int)t.n, s));
            ++launch_counter();
        }
    } else {
        // walk the query points in path-key order so the 32 lanes of a warp share most of their path
        DevBuf<uint64_t> key((size_t)m, s), key_s((size_t)m, s);
        DevBuf<uint32_t> iota((size_t)m, s);
        torder.alloc((size_t)m, s);
        PNBX_LAUNCH(point_keys, nb(m), 256, 0, s, d_tgt, m, t.root4.p, key.p);
        PNBX_LAUNCH(iota32, nb(m), 256, 0, s, iota.p, m);
        size_t bytes = 0;
        PNBX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, key.p, key_s.p, iota.p, torder.p, (int)m, 0, 63, s));
        DevBuf<uint8_t> tmp(bytes, s);
        PNBX_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(), bytes, key.p, key_s.p, iota.p, torder.p, (int)m, 0, 63, s));
        ++launch_counter();
    }
    tm.end();

    // SMODE 1/2 kernels read the softening array unconditionally: give them zeros if the tree has an hmax payload
    // but no softenings any more (set_softenings(None) after the build)
    DevBuf<float> zero_h;
    if (!t.has_h && t.has_hmax && !ex.f64 && !d_counters) {
        zero_h.alloc((size_t)t.n + 4, s);
        PNBX_CUDA(cudaMemsetAsync(zero_h.p, 0, ((size_t)t.n + 4) * sizeof(float), s));
    }
    auto fill = [&](auto& a) {
        a.rec = t.rec.p;
        a.gated = t.has_hmax ? 1 : 0;
        a.counters = d_counters;
        a.spos = t.spos.p; a.smass = t.has_mass ? t.smass.p : nullptr;
        a.sh = t.has_h ? t.sh.p : nullptr;
        a.perm = t.perm.p;
        a.m = m_walk; a.self = self ? 1 : 0;
        a.tlist = tlist.p; a.tgt_begin = tgt_begin; a.tgt = d_tgt; a.torder = torder.p;
        a.tree_order = tree_order ? 1 : 0;
        a.cyc_block = ((tree_order && ex.block_cyclic) || pts_cyclic) ? ex.shard_block : 0;
        a.cyc_rank = ex.shard_rank; a.cyc_world = ex.shard_world;
        a.theta2 = theta * theta;
        a.rc = t.root4.p;
        a.kernel = t.kernel;
        a.out_pot = d_pot; a.out_acc = d_acc;
        a.slices = slices ? *slices : OutSlices{};
        a.budget = INT_MAX; a.over_list = nullptr; a.n_over = nullptr; a.n_front = nullptr;
    };
    const int order = std::max(1, t.order);  // order 0 and 1 are both monopoles
    tm.begin("octree.walk");
    kernel_events().begin(s);
    if (d_counters) {  // counting pass (order-independent): traversal decisions only
        WalkArgs<double> a;
        fill(a);
        a.moments = t.moments.p; a.K = t.n_moments; a.src = nullptr; a.src_h = t.has_h ? t.sh.p : nullptr;
        PNBX_LAUNCH((walk_kernel<1, 0, double, 3>), (unsigned)std::max<int64_t>(1, ceil_div(m_walk, WT)), WT, 0, s, a);
    } else if (ex.f64) {
        WalkArgs<double> a;
        fill(a);
        a.moments = t.moments.p; a.K = t.n_moments; a.src = nullptr; a.src_h = t.has_h ? t.sh.p : nullptr;
        launch_walk<double, 3>(order, want, a, s);
    } else {
        WalkArgs<float> a;
        fill(a);
        a.moments = t.moments32.p; a.K = t.rec32; a.src = t.src32.p; a.src_h = t.has_h ? t.sh32.p : zero_h.p;
        const bool any_soft = t.has_h || t.has_hmax;
        // few targets: one warp per target
        // switch-over sizes (measured, profiles/r02_walk_kernel_ncu.md): the warp-per-target kernel costs ~40 warp
        // instructions per visit and target, so for the balanced self evaluations it only wins on very small sets
        // (N = 15 682: 0.61 vs 0.72 ms; N = 60 000: 2.8 vs 0.87 ms). The lane-per-target kernel's time on a query grid is
        // set by its heaviest warps (points in a softened core sum thousands of particles): grids switch much later
        // (N = 1e7 zoom set: 1.25e5 points 6.6 vs 8.1 ms, 2e5 points 10.4 vs 7.2 ms; N = 1e8 on 8 GPUs, 1.25e5 points
        // per GPU: 14 vs 128 ms)
        const char* wpt_env = getenv("PNBX_WPT_MAX_TARGETS");
        const int64_t wpt_max = wpt_env ? atoll(wpt_env) : (int64_t)(self ? 16384 : 131072);
        const bool wpt = wpt_basis <= wpt_max;
        // Larger point sets, HYBRID: the lane-per-target kernel's time on a query grid is set by its slowest warps —
        // 32 neighbouring points of a sparse log-spaced grid whose paths diverge (the warp visits the union), or points
        // in a softened core where the gate opens whole subtrees. A warp whose walk outgrows the budget (a few times
        // the cost of a typical warp of a self evaluation) gives up and its points are walked one per warp instead.
        // Which warps give up is a deterministic function of their 32 points. Measured (zoom set, ms, lane / warp /
        // hybrid): N = 1e7: 2.5e5 points 7.7 / 12.9 / 5.7, 1e6 points 9.9 / 52.3 / 9.1; N = 1e8: 2.5e5 points
        // 141 / 21.2 / 24.0, 1e6 points 129 / 79.3 / 52.8. Giving up early on warps with few active lanes (sampled
        // every budget / 8) made it slower (1e6 points: 13.2 at N = 1e7, 63.3 at N = 1e8): not built.
        const char* hy_env = getenv("PNBX_WALK_HYBRID_COST");
        const int64_t budget = hy_env ? atoll(hy_env) : 64000;
        if (!wpt && !self && budget > 0 && m_walk < ((int64_t)1 << 31)) {  // the hand-over count is an int
            DevBuf<uint32_t> over((size_t)m_walk, s);
            DevBuf<int> n_over(1, s);
            PNBX_CUDA(cudaMemsetAsync(n_over.p, 0, sizeof(int), s));
            a.budget = (int)std::min<int64_t>(budget, INT_MAX / 2);
            a.over_list = over.p;
            a.n_over = n_over.p;
            if (!any_soft) launch_walk<float, 0, true>(order, want, a, s);
            else if (t.kernel == PNBX_KERNEL_SPLINE) launch_walk<float, 2, true>(order, want, a, s);
            else launch_walk<float, 1, true>(order, want, a, s);
            a.budget = INT_MAX;
            a.torder = over.p;
            a.cyc_block = 0;
            a.n_front = n_over.p;
            if (!any_soft) launch_walk_wpt<0>(order, want, a, s);
            else if (t.kernel == PNBX_KERNEL_SPLINE) launch_walk_wpt<2>(order, want, a, s);
            else launch_walk_wpt<1>(order, want, a, s);
        } else if (wpt) {
            if (!any_soft) launch_walk_wpt<0>(order, want, a, s);
            else if (t.kernel == PNBX_KERNEL_SPLINE) launch_walk_wpt<2>(order, want, a, s);
            else launch_walk_wpt<1>(order, want, a, s);
        } else if (!any_soft) launch_walk<float, 0>(order, want, a, s);
        else if (t.kernel == PNBX_KERNEL_SPLINE) launch_walk<float, 2>(order, want, a, s);
        else launch_walk<float, 1>(order, want, a, s);
    }
    kernel_events().end(s);
    PNBX_CUDA(cudaGetLastError());
    tm.end();
}

}  // namespace pnbx

using namespace pnbx;

extern "C" int pnbx_tree_eval(pnbx_tree* tp, const double* tgt_pos, int64_t m, int64_t tgt_begin, double theta, int want,
                              double* out_pot, double* out_acc, const pnbx_opts* opts) {
    return guarded([&] {
        if (!tp) throw ArgError{PNBX_ERR_ARG, "tree is NULL"};
        auto& t = *reinterpret_cast<pnbx_tree_impl*>(tp);
        if (!t.has_payload) throw ArgError{PNBX_ERR_STATE, "mass payload not built; call build_mass() before compute"};
        if (!(want & (PNBX_WANT_POT | PNBX_WANT_ACC)) || (want & ~3)) throw ArgError{PNBX_ERR_ARG, "bad `want` mask"};
        if (m < 0) throw ArgError{PNBX_ERR_ARG, "negative size"};
        if ((want & PNBX_WANT_POT) && !out_pot && m > 0) throw ArgError{PNBX_ERR_ARG, "out_pot is NULL"};
        if ((want & PNBX_WANT_ACC) && !out_acc && m > 0) throw ArgError{PNBX_ERR_ARG, "out_acc is NULL"};
        const bool self = tgt_pos == nullptr;
        const bool cyc = self && opts && (opts->flags & PNBX_FLAG_TREE_ORDER) && (opts->flags & PNBX_FLAG_BLOCK_CYCLIC);
        if (cyc) {
            if (m != pnbx_shard_count(t.n, opts->shard_block, opts->shard_world, opts->shard_rank))
                throw ArgError{PNBX_ERR_ARG, "m must equal pnbx_shard_count(n, block, world, rank) for a block-cyclic shard"};
        } else if (self && (tgt_begin < 0 || tgt_begin + m > t.n)) {
            throw ArgError{PNBX_ERR_ARG, "target shard outside [0, N)"};
        }
        if (m >= ((int64_t)1 << 31) - 16) throw ArgError{PNBX_ERR_ARG, "M must be < 2^31"};
        // PNBX_DEVICES tree, whole-array host call without an explicit device: all GPUs (multi.cu)
        if (!t.replicas.empty() && m > 0 && (!opts || (opts->mem_space == PNBX_MEM_HOST && opts->device < 0 &&
                                                        !(opts->flags & PNBX_FLAG_TREE_ORDER))) &&
            (!self || (tgt_begin == 0 && m == t.n)) && multi_tree_eval(t, tgt_pos, m, theta, want, out_pot, out_acc, opts))
            return;
        pnbx_opts o = opts ? *opts : pnbx_opts{-1, PNBX_MEM_HOST, 0, 0, nullptr};
        if (o.device < 0) o.device = t.device;
        if (o.device != t.device) throw ArgError{PNBX_ERR_ARG, "tree lives on a different device"};
        Exec ex = make_exec(&o);
        StageTimer tm(ex.stream);
        if (m == 0) { finish_exec(ex); return; }
        OutArray<double> o_pot, o_acc;
        if (want & PNBX_WANT_POT) o_pot.bind(out_pot, (size_t)m, ex);
        if (want & PNBX_WANT_ACC) o_acc.bind(out_acc, (size_t)3 * m, ex);
        InArray<double> i_tgt;
        if (!self) i_tgt.bind(tgt_pos, (size_t)3 * m, ex);
        tree_begin_use(t, ex.stream);
        tree_walk(t, ex, self ? nullptr : i_tgt.d, m, tgt_begin, theta, want, o_pot.d, o_acc.d, tm, nullptr);
        tree_end_use(t, ex.stream);
        o_pot.finish(ex);
        o_acc.finish(ex);
        finish_exec(ex);
    });
}

// Traversal statistics of the walk tree.rs:1069-1370 would do for these targets: totals over all targets of
// node visits, accepted nodes, leaf visits and leaf particles (the oracle's counters; input to the work model
// of BASELINE.md §3), plus out5[4] = nodes visited per WARP summed over warps (union of 32 paths: the walk's real cost). One extra decisions-only kernel; results are exact integers.
extern "C" int pnbx_tree_walk_counters(pnbx_tree* tp, const double* tgt_pos, int64_t m, int64_t tgt_begin, double theta,
                                       int64_t* out5, const pnbx_opts* opts) {
    return guarded([&] {
        if (!tp || !out5) throw ArgError{PNBX_ERR_ARG, "NULL argument"};
        auto& t = *reinterpret_cast<pnbx_tree_impl*>(tp);
        if (!t.has_payload) throw ArgError{PNBX_ERR_STATE, "mass payload not built; call build_mass() before compute"};
        const bool self = tgt_pos == nullptr;
        const bool cyc = self && opts && (opts->flags & PNBX_FLAG_TREE_ORDER) && (opts->flags & PNBX_FLAG_BLOCK_CYCLIC);
        if (m < 0 || (self && !cyc && (tgt_begin < 0 || tgt_begin + m > t.n))) throw ArgError{PNBX_ERR_ARG, "bad target range"};
        pnbx_opts o = opts ? *opts : pnbx_opts{-1, PNBX_MEM_HOST, 0, 0, nullptr};
        if (o.device < 0) o.device = t.device;
        Exec ex = make_exec(&o);
        StageTimer tm(ex.stream);
        DevBuf<unsigned long long> cnt(5, ex.stream);
        PNBX_CUDA(cudaMemsetAsync(cnt.p, 0, 5 * sizeof(unsigned long long), ex.stream));
        InArray<double> i_tgt;
        if (!self) i_tgt.bind(tgt_pos, (size_t)3 * m, ex);
        tree_begin_use(t, ex.stream);
        if (m > 0) tree_walk(t, ex, self ? nullptr : i_tgt.d, m, tgt_begin, theta, 1, nullptr, nullptr, tm, cnt.p);
        tree_end_use(t, ex.stream);
        unsigned long long h[5];
        PNBX_CUDA(cudaMemcpyAsync(h, cnt.p, sizeof(h), cudaMemcpyDeviceToHost, ex.stream));
        PNBX_CUDA(cudaStreamSynchronize(ex.stream));
        for (int i = 0; i < 5; ++i) out5[i] = (int64_t)h[i];
        finish_exec(ex);
    });
}

// Original particle indices of sorted (tree-order) positions [begin, begin+m): the scatter map for results of
// pnbx_tree_eval(..., PNBX_FLAG_TREE_ORDER). `out` is a host or device int64 array per opts->mem_space.
namespace pnbx {
__global__ void perm_to_i64(const uint32_t* __restrict__ perm, int64_t begin, int64_t m, int64_t cyc_block, int cyc_rank,
                            int cyc_world, int64_t* __restrict__ out) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const int64_t s = cyc_block > 0 ? ((k / cyc_block) * cyc_world + cyc_rank) * cyc_block + k % cyc_block : begin + k;
    out[k] = perm[s];
}
}  // namespace pnbx
extern "C" int pnbx_tree_get_order(const pnbx_tree* tp, int64_t begin, int64_t m, int64_t* out, const pnbx_opts* opts) {
    return guarded([&] {
        if (!tp || (!out && m > 0)) throw ArgError{PNBX_ERR_ARG, "NULL argument"};
        const auto& t = *reinterpret_cast<const pnbx_tree_impl*>(tp);
        pnbx_opts o = opts ? *opts : pnbx_opts{-1, PNBX_MEM_HOST, 0, 0, nullptr};
        if (o.device < 0) o.device = t.device;
        o.flags |= (o.flags & PNBX_FLAG_BLOCK_CYCLIC) ? PNBX_FLAG_TREE_ORDER : 0;
        Exec ex = make_exec(&o);
        if (ex.block_cyclic) {
            if (m != pnbx_shard_count(t.n, ex.shard_block, ex.shard_world, ex.shard_rank))
                throw ArgError{PNBX_ERR_ARG, "m must equal pnbx_shard_count(n, block, world, rank)"};
        } else if (begin < 0 || m < 0 || begin + m > t.n) {
            throw ArgError{PNBX_ERR_ARG, "range outside [0, N)"};
        }
        if (m > 0) {
            OutArray<int64_t> oa;
            oa.bind(out, (size_t)m, ex);
            tree_begin_use(t, ex.stream);
            PNBX_LAUNCH(perm_to_i64, (unsigned)ceil_div(m, 256), 256, 0, ex.stream, t.perm.p, begin, m,
                        ex.block_cyclic ? ex.shard_block : (int64_t)0, ex.shard_rank, ex.shard_world, oa.d);
            PNBX_CUDA(cudaGetLastError());
            tree_end_use(t, ex.stream);
            oa.finish(ex);
        }
        finish_exec(ex);
    });
}
