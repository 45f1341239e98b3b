// tree_build.cu — K2..K6: GPU construction of the reference's octree (tree.rs:628-1067).
//
//   K2 bbox            tree.rs:628-654   (launch_bbox + root_from_bbox, float64)
//   K3 path keys       tree.rs:815-838   float64 descent with the reference's own centre arithmetic
//   K4 radix sort      (cub::DeviceRadixSort, stable)  = the reference's stable octant bucketing
//   K5 node emission   tree.rs:804-864, 736-776: level-by-level from sorted key ranges, then the
//                      reference's creation-order numbering, first_subnode / next_branch links
//   K6 payloads        tree.rs:866-965, 1014-1067: mass/COM, hmax, P2M/M2M, bottom-up per level,
//                      float64 with the reference's operation order (no FMA contraction)
#include <cstring>
#include <mutex>

#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/reverse_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "multipole.cuh"
#include "tree.cuh"

namespace pnbx {
namespace {

// ------------------------------------------------------------------------------------------ K2
__global__ void root_from_bbox(const double* __restrict__ bb, double* __restrict__ root4) {
    double half = 0.0;
    for (int i = 0; i < 3; ++i) {
        root4[i] = __ddiv_rn(__dadd_rn(bb[i], bb[3 + i]), 2.0);
        half = fmax(half, __ddiv_rn(__dsub_rn(bb[3 + i], bb[i]), 2.0));
    }
    if (half == 0.0) half = 1e-6;  // tree.rs:650-652
    root4[3] = half;
}

// ------------------------------------------------------------------------------------------ K3
// Octant digits of levels 1..42 by the reference's descent: oct = (x>=cx) | (y>=cy)<<1 | (z>=cz)<<2
// against the *rounded* child centres centre +/- half/2 (tree.rs:818-838).
// `packed` (nullable): (x, y, z, m) as one 32-byte record per particle, written here because the positions are in
// registers anyway; the gather into tree order then fetches ONE sector per particle instead of two or three.
__global__ void path_keys(const double* __restrict__ pos, const double* __restrict__ mass, int64_t n,
                          const double* __restrict__ root4, int levels, uint64_t* __restrict__ key_hi,
                          uint64_t* __restrict__ key_lo, double4* __restrict__ packed) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
    if (packed) packed[i] = make_double4(x, y, z, mass ? mass[i] : 1.0);
    double cx = root4[0], cy = root4[1], cz = root4[2], hf = root4[3];
    uint64_t hi = 0, lo = 0;
    for (int l = 1; l <= levels; ++l) {
        const unsigned ox = x >= cx, oy = y >= cy, oz = z >= cz;
        const uint64_t oct = ox | (oy << 1) | (oz << 2);
        if (l <= KEY_LEVELS_HI) hi |= oct << (3 * (KEY_LEVELS_HI - l));
        else lo |= oct << (3 * (KEY_LEVELS - l));
        const double off = hf * 0.5;  // exact (== hf / 2.0)
        cx = __dadd_rn(cx, ox ? off : -off);
        cy = __dadd_rn(cy, oy ? off : -off);
        cz = __dadd_rn(cz, oz ? off : -off);
        hf = off;
    }
    key_hi[i] = hi;
    if (key_lo) key_lo[i] = lo;
}

__global__ void iota_u32(uint32_t* p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}
template <class T>
__global__ void gather_u32(const T* __restrict__ in, const uint32_t* __restrict__ idx, int64_t n, T* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[idx[i]];
}

// ------------------------------------------------------------------------------------------ K5
// Topology from the sorted path keys without any level-by-level host round trip.
//
// A cell at level L is the set of particles sharing their first L octant digits; it is a NODE of the reference's
// tree iff it is the root or its parent cell holds more than leaf_capacity particles (tree.rs:848-851), and an
// INTERNAL node iff it holds more than leaf_capacity itself. Populations only shrink down a path, so with
//   b[i] = digits shared by sorted particles i-1 and i                      (b[0] = -1)
//   D[i] = deepest level at which particle i's cell is still over-full
//        = max over windows j in [i-cap, i] of the digits shared by particles j and j+cap   (sorted => the cap+1
//          particles of a window all share that prefix; -1 if there is no window)
// particle i's leaf sits at level D[i]+1, particle i starts a new leaf ("head") iff b[i] < D[i]+1, and the nodes that
// START at a head i are exactly the levels b[i]+1 .. D[i]+1 (the last one is the leaf). Listing the heads in
// ascending order and each head's levels in ascending order IS the depth-first pre-order of the tree, so one
// exclusive scan numbers every node, and a single 8-byte-per-level D2H (sizes for the allocations) is the only
// synchronisation of the build.
constexpr int MAX_LEVELS = KEY_LEVELS + 2;  // histogram bins: levels 0 .. 43

__device__ __forceinline__ int shared_digits(uint64_t ahi, uint64_t alo, uint64_t bhi, uint64_t blo, bool two_words) {
    const uint64_t x = ahi ^ bhi;
    if (x) return (__clzll((long long)x) - 1) / 3;  // bit 63 is never set: digit l sits at bits 3*(21-l)+{0,1,2}
    if (!two_words) return KEY_LEVELS_HI;
    const uint64_t y = alo ^ blo;
    return y ? KEY_LEVELS_HI + (__clzll((long long)y) - 1) / 3 : KEY_LEVELS;
}
// do sorted particles j and i share their first L digits?
__device__ __forceinline__ bool same_prefix(const uint64_t* __restrict__ khi, const uint64_t* __restrict__ klo,
                                            int64_t j, uint64_t ihi, uint64_t ilo, int L) {
    if (L <= KEY_LEVELS_HI) return ((khi[j] ^ ihi) >> (63 - 3 * L)) == 0;  // L = 0: shift 63, bit 63 is clear
    return khi[j] == ihi && ((klo[j] ^ ilo) >> (63 - 3 * (L - KEY_LEVELS_HI))) == 0;
}

// W[j] = digits shared by sorted particles j and j+cap (the over-full witness of every particle in between)
__global__ void window_digits(const uint64_t* __restrict__ khi, const uint64_t* __restrict__ klo, int64_t nw,
                              int64_t cap, int8_t* __restrict__ W) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nw) return;
    W[j] = (int8_t)shared_digits(khi[j], klo ? klo[j] : 0, khi[j + cap], klo ? klo[j + cap] : 0, klo != nullptr);
}
struct MaxI8 {
    __device__ __forceinline__ int8_t operator()(int8_t a, int8_t b) const { return a > b ? a : b; }
};
struct BlockOf {  // segment key of the van Herk sliding-window maximum: blocks of `w` consecutive windows
    int64_t w;
    __host__ __device__ __forceinline__ int64_t operator()(int64_t j) const { return j / w; }
};
struct RevBlockOf {  // the same blocks, enumerated from the end (for the suffix maxima)
    int64_t w, last;
    __host__ __device__ __forceinline__ int64_t operator()(int64_t r) const { return (last - r) / w; }
};

// Per sorted particle: b, D, head flag, number of nodes that start here (0 for non-heads); per-level node and
// internal-node histograms; overflow flag if an over-full cell survives at the last key level.
// Small capacities evaluate the windows directly; large ones read the prefix / suffix maxima (P, S) of W.
__global__ void __launch_bounds__(256) leaf_levels(const uint64_t* __restrict__ khi, const uint64_t* __restrict__ klo,
                                                   int64_t n, int64_t cap, int max_level,
                                                   const int8_t* __restrict__ P, const int8_t* __restrict__ S,
                                                   int8_t* __restrict__ b_out, int8_t* __restrict__ D_out,
                                                   int32_t* __restrict__ nodes_here,
                                                   unsigned long long* __restrict__ hist /* [2][MAX_LEVELS] + overflow */) {
    __shared__ unsigned int sh[2][MAX_LEVELS];
    for (int t = threadIdx.x; t < 2 * MAX_LEVELS; t += blockDim.x) (&sh[0][0])[t] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool two = klo != nullptr;
    if (i < n) {
        const uint64_t ihi = khi[i], ilo = two ? klo[i] : 0;
        const int b = i == 0 ? -1 : shared_digits(khi[i - 1], two ? klo[i - 1] : 0, ihi, ilo, two);
        int D = -1;
        const int64_t nw = n - cap;  // number of windows
        if (nw > 0) {
            const int64_t lo = i - cap > 0 ? i - cap : 0, hi = i < nw - 1 ? i : nw - 1;
            if (P == nullptr) {
                for (int64_t j = lo; j <= hi; ++j) {
                    const int d = shared_digits(khi[j], two ? klo[j] : 0, khi[j + cap], two ? klo[j + cap] : 0, two);
                    D = d > D ? d : D;
                }
            } else if (lo <= hi) {
                const int64_t w = cap + 1;
                if (lo / w != hi / w) D = max((int)S[lo], (int)P[hi]);
                else D = (lo % w == 0) ? (int)P[hi] : (int)S[lo];  // same block: a clipped window at either end
            }
        }
        const int leaf_level = D + 1;
        const bool head = b < leaf_level;
        b_out[i] = (int8_t)b;
        D_out[i] = (int8_t)D;
        nodes_here[i] = head ? leaf_level - b : 0;
        if (head) {
            if (leaf_level > max_level) atomicOr(&hist[2 * MAX_LEVELS], 1ull);
            else {
                for (int L = b + 1; L < leaf_level; ++L) { atomicAdd(&sh[0][L], 1u); atomicAdd(&sh[1][L], 1u); }
                atomicAdd(&sh[0][leaf_level], 1u);
            }
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * MAX_LEVELS; t += blockDim.x) {
        const unsigned int v = (&sh[0][0])[t];
        if (v) atomicAdd(&hist[t], (unsigned long long)v);
    }
}

// Nodes in depth-first order (temporary): everything the renumbering and the final scatter need.
struct DfsNodes {
    uint32_t* start; uint32_t* count; uint32_t* parent; uint32_t* mask;  // mask: octants of the existing children
    uint8_t* level; uint8_t* digit;
    double* center; double* half; uint64_t* phi; uint64_t* plo;
};
// One thread per head: replays the reference's descent (tree.rs:818-838: centre +/- half/2, rounded additions) along
// the particle's key digits and emits the nodes that start here — levels b+1 .. D+1 — with their particle ranges
// (galloping searches in the sorted keys), their parent and their octant bit in the parent's child mask.
__global__ void __launch_bounds__(128) emit_nodes(const uint64_t* __restrict__ khi, const uint64_t* __restrict__ klo,
                                                  int64_t n, const int8_t* __restrict__ bq, const int8_t* __restrict__ Dq,
                                                  const int32_t* __restrict__ nodes_here, const int32_t* __restrict__ base,
                                                  const double* __restrict__ root4, DfsNodes d) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || nodes_here[i] == 0) return;
    const bool two = klo != nullptr;
    const uint64_t ihi = khi[i], ilo = two ? klo[i] : 0;
    const int b = bq[i], leaf_level = Dq[i] + 1;
    const int64_t id0 = base[i];  // DFS index of the node at level b+1
    auto digit_of = [&](int L) -> unsigned {
        return L <= KEY_LEVELS_HI ? (unsigned)((ihi >> (3 * (KEY_LEVELS_HI - L))) & 7u)
                                  : (unsigned)((ilo >> (3 * (KEY_LEVELS - L))) & 7u);
    };
    // ---- particle ranges, deepest first: the leaf ends at the next head, every ancestor extends its child's range
    int64_t e = i + 1;
    while (e < n && nodes_here[e] == 0) ++e;
    d.count[id0 + (leaf_level - b - 1)] = (uint32_t)(e - i);
    for (int L = leaf_level - 1; L > b; --L) {
        if (L == 0) e = n;
        else if (e < n && same_prefix(khi, klo, e, ihi, ilo, L)) {
            int64_t lo = e, step = 1, hi = e + 1;
            while (hi < n && same_prefix(khi, klo, hi, ihi, ilo, L)) { lo = hi; step <<= 1; hi = lo + step; }
            if (hi > n) hi = n;
            while (hi - lo > 1) {  // invariant: lo shares the prefix, hi does not (or hi == n)
                const int64_t mid = lo + ((hi - lo) >> 1);
                if (same_prefix(khi, klo, mid, ihi, ilo, L)) lo = mid; else hi = mid;
            }
            e = hi;
        }
        d.count[id0 + (L - b - 1)] = (uint32_t)(e - i);
    }
    // ---- parent of the first node that starts here: the level-b cell around i; it starts at the first particle
    // sharing b digits with i (i-1 does, by the definition of b)
    uint32_t parent0 = 0xffffffffu;
    if (i > 0) {
        int64_t lo = i - 1, step = 1, p = lo - 1;
        while (p >= 0 && same_prefix(khi, klo, p, ihi, ilo, b)) { lo = p; step <<= 1; p = lo - step; }
        if (p < -1) p = -1;
        while (lo - p > 1) {  // invariant: lo shares the prefix, p does not (or p == -1)
            const int64_t mid = p + ((lo - p) >> 1);
            if (same_prefix(khi, klo, mid, ihi, ilo, b)) lo = mid; else p = mid;
        }
        parent0 = (uint32_t)(base[lo] + (b - bq[lo] - 1));
    }
    // ---- geometry by the reference's descent, nodes from level b+1 on
    double cx = root4[0], cy = root4[1], cz = root4[2], hf = root4[3];
    for (int L = 0; L <= leaf_level; ++L) {
        if (L > 0) {
            const unsigned o = digit_of(L);
            const double off = hf * 0.5;  // exact (== hf / 2.0)
            cx = __dadd_rn(cx, (o & 1) ? off : -off);
            cy = __dadd_rn(cy, (o & 2) ? off : -off);
            cz = __dadd_rn(cz, (o & 4) ? off : -off);
            hf = off;
        }
        if (L <= b) continue;
        const int64_t id = id0 + (L - b - 1);
        const unsigned o = L > 0 ? digit_of(L) : 0u;
        const uint32_t par = L == b + 1 ? parent0 : (uint32_t)(id - 1);
        d.start[id] = (uint32_t)i;
        d.level[id] = (uint8_t)L;
        d.digit[id] = (uint8_t)o;
        d.parent[id] = par;
        if (par != 0xffffffffu) atomicOr(&d.mask[par], 1u << o);  // integer OR: order-independent
        d.center[3 * id] = cx; d.center[3 * id + 1] = cy; d.center[3 * id + 2] = cz;
        d.half[id] = hf;
        uint64_t ph = 0, pl = 0;  // octant path of the node: the particle's key truncated to L digits
        if (L > 0) {
            if (L <= KEY_LEVELS_HI) ph = ihi >> (63 - 3 * L) << (63 - 3 * L);
            else { ph = ihi; pl = ilo >> (63 - 3 * (L - KEY_LEVELS_HI)) << (63 - 3 * (L - KEY_LEVELS_HI)); }
        }
        d.phi[id] = ph;
        d.plo[id] = pl;
    }
}
__global__ void child_counts(const uint32_t* __restrict__ mask, int64_t nn, int32_t* __restrict__ nchild) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nn) nchild[i] = __popc(mask[i]);
}
// Reference creation order (tree.rs:804-864): a node's children get consecutive ids when the node is subdivided,
// before any of them is visited => id(first child of X) = 1 + (children of all internal nodes before X in depth-first
// pre-order) = 1 + exclusive scan of the child counts in DFS order; a child's rank among its siblings is the number of
// existing octants below its own.
__global__ void assign_ref_ids(const uint32_t* __restrict__ parent, const uint8_t* __restrict__ digit,
                               const uint32_t* __restrict__ mask, const int32_t* __restrict__ scan, int64_t nn,
                               int32_t* __restrict__ ref) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const uint32_t p = parent[i];
    ref[i] = p == 0xffffffffu ? 0 : 1 + scan[p] + __popc(mask[p] & ((1u << digit[i]) - 1u));
}
struct FinalArrays {
    double* center; double* half; uint8_t* depth; uint32_t* start; uint32_t* count; int32_t* first_subnode;
    int32_t* next_branch; uint64_t* path_hi; uint64_t* path_lo; uint8_t* nchild;
};
// Scatter into the reference numbering, with the walk links of tree.rs:736-776: first_subnode = first child,
// next_branch = next sibling, else the parent's = the first node that starts at the particle right after this node's
// range (depth-first order), or "none" (-1) at the end of the particle list.
__global__ void scatter_nodes(DfsNodes d, const int32_t* __restrict__ scan, const int32_t* __restrict__ ref,
                              const int32_t* __restrict__ base, int64_t nn, int64_t n, FinalArrays f,
                              uint8_t* __restrict__ sortkey, int32_t* __restrict__ sortval,
                              int32_t* __restrict__ dfs_of_ref) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const int32_t r = ref[i];
    f.center[3 * r] = d.center[3 * i]; f.center[3 * r + 1] = d.center[3 * i + 1]; f.center[3 * r + 2] = d.center[3 * i + 2];
    f.half[r] = d.half[i];
    f.depth[r] = d.level[i];
    f.start[r] = d.start[i];
    f.count[r] = d.count[i];
    const int nc = __popc(d.mask[i]);
    f.first_subnode[r] = nc > 0 ? 1 + scan[i] : -1;
    const int64_t e = (int64_t)d.start[i] + d.count[i];
    f.next_branch[r] = e < n ? ref[base[e]] : -1;
    f.path_hi[r] = d.phi[i];
    f.path_lo[r] = d.plo[i];
    f.nchild[r] = (uint8_t)nc;
    sortkey[i] = nc > 0 ? d.level[i] : (uint8_t)64;  // internal nodes grouped by level, all leaves behind them
    sortval[i] = r;
    dfs_of_ref[r] = (int32_t)i;
}
__global__ void empty_root(const double* __restrict__ root4, FinalArrays f) {
    f.center[0] = root4[0]; f.center[1] = root4[1]; f.center[2] = root4[2]; f.half[0] = root4[3];
    f.depth[0] = 0; f.start[0] = 0; f.count[0] = 0; f.first_subnode[0] = -1; f.next_branch[0] = -1;
    f.path_hi[0] = 0; f.path_lo[0] = 0; f.nchild[0] = 0;
}

// ---- leaf-internal order: ascending original index (stable bucketing, tree.rs:813-828)
// The key sort orders the particles of a leaf by their deeper digits; the reference keeps them in original order.
// Leaves hold at most leaf_capacity particles: for small capacities every head sorts its own leaf in registers.
constexpr int LEAF_SORT_MAX = 32;
__global__ void __launch_bounds__(128) sort_leaves_small(const int32_t* __restrict__ nodes_here, int64_t n,
                                                         uint32_t* __restrict__ perm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || nodes_here[i] == 0) return;
    uint32_t v[LEAF_SORT_MAX];
    int c = 0;
    v[c++] = perm[i];
    for (int64_t j = i + 1; j < n && c < LEAF_SORT_MAX && nodes_here[j] == 0; ++j) {  // insertion sort while loading
        const uint32_t x = perm[j];
        int k = c++;
        while (k > 0 && v[k - 1] > x) { v[k] = v[k - 1]; --k; }
        v[k] = x;
    }
    for (int k = 0; k < c; ++k) perm[i + k] = v[k];
}
// Large capacities: stable sort of the particles by leaf ordinal from the original order.
__global__ void head_flags(const int32_t* __restrict__ nodes_here, int64_t n, uint32_t* __restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = nodes_here[i] != 0;
}
__global__ void leaf_ordinal_to_particles(const uint32_t* __restrict__ ord_sorted, const uint32_t* __restrict__ perm,
                                          int64_t n, uint32_t* __restrict__ ord_orig) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) ord_orig[perm[s]] = ord_sorted[s];
}
// ---- sorted copies
__global__ void gather_sources(const double* __restrict__ pos, const double* __restrict__ mass,
                               const uint32_t* __restrict__ perm, int64_t n, double* __restrict__ spos,
                               double* __restrict__ smass) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t i = perm[s];
    const double x = pos[3 * (int64_t)i], y = pos[3 * (int64_t)i + 1], z = pos[3 * (int64_t)i + 2];
    spos[3 * s] = x; spos[3 * s + 1] = y; spos[3 * s + 2] = z;
    if (smass) smass[s] = mass[i];
}
__global__ void gather_packed_sources(const double4* __restrict__ packed, const uint32_t* __restrict__ perm, int64_t n,
                                      double* __restrict__ spos, double* __restrict__ smass) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const double4 v = packed[perm[s]];
    spos[3 * s] = v.x; spos[3 * s + 1] = v.y; spos[3 * s + 2] = v.z;
    if (smass) smass[s] = v.w;
}
__global__ void gather_soft(const double* __restrict__ h, const uint32_t* __restrict__ perm, int64_t n,
                            double* __restrict__ sh, float* __restrict__ sh32) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const double v = h[perm[s]];
    sh[s] = v;
    const float hc = (float)fmax(v, 0.0);
    sh32[s] = hc * hc;  // the fp32 walk compares squares: max(h_s,0)^2 (tree_walk.cu leaf_pass1_f32)
}

// ------------------------------------------------------------------------------------------ K6
// fp32 walk records from the float64 moments (layout: multipole.cuh, m2p_fast). Called by the payload kernels right
// after a node's moments are final, so the float64 moments are not read again by a separate packing pass.
__device__ __forceinline__ void pack_walk_record(const double* __restrict__ m, int order, float* __restrict__ o) {
    using namespace mp;
    if (order <= 1) { o[0] = (float)m[I000]; return; }
    o[0] = (float)m[I000];
    if (order <= 3) {  // m2p_fast layout: traceless T = 3 (S - trS/3 I), folded cubic C' = 15 C - 3 (w.u)(u.u)
        const double tr = m[I200] + m[I020] + m[I002];
        o[1] = (float)(3.0 * m[I200] - tr); o[2] = (float)(3.0 * m[I020] - tr); o[3] = (float)(3.0 * m[I002] - tr);
        o[4] = (float)(1.5 * m[I110]); o[5] = (float)(1.5 * m[I101]); o[6] = (float)(1.5 * m[I011]);
        o[7] = 0.f;
        if (order == 3) {
            const double wx = 3.0 * m[I300] + m[I120] + m[I102];
            const double wy = 3.0 * m[I030] + m[I210] + m[I012];
            const double wz = 3.0 * m[I003] + m[I201] + m[I021];
            // field order: 300 030 003 210 201 120 102 021 012 111; the w component is that of the odd-power axis
            o[8]  = (float)(15.0 * m[I300] - 3.0 * wx); o[9]  = (float)(15.0 * m[I030] - 3.0 * wy);
            o[10] = (float)(15.0 * m[I003] - 3.0 * wz); o[11] = (float)(15.0 * m[I210] - 3.0 * wy);
            o[12] = (float)(15.0 * m[I201] - 3.0 * wz); o[13] = (float)(15.0 * m[I120] - 3.0 * wx);
            o[14] = (float)(15.0 * m[I102] - 3.0 * wx); o[15] = (float)(15.0 * m[I021] - 3.0 * wz);
            o[16] = (float)(15.0 * m[I012] - 3.0 * wy); o[17] = (float)(15.0 * m[I111]);
            o[18] = o[19] = 0.f;
        }
        return;
    }
    // orders 4, 5 (m2p_fast45): [1..6] 6S, [7] 3 trS, [8..17] octupole, [18..20] 9 v = 3 w
    o[1] = (float)(6.0 * m[I200]); o[2] = (float)(6.0 * m[I020]); o[3] = (float)(6.0 * m[I002]);
    o[4] = (float)(3.0 * m[I110]); o[5] = (float)(3.0 * m[I101]); o[6] = (float)(3.0 * m[I011]);
    o[7] = (float)(3.0 * (m[I200] + m[I020] + m[I002]));
    for (int t = 0; t < 10; ++t) o[8 + t] = (float)m[I300 + t];
    o[18] = (float)(3.0 * (3.0 * m[I300] + m[I120] + m[I102]));
    o[19] = (float)(3.0 * (3.0 * m[I030] + m[I210] + m[I012]));
    o[20] = (float)(3.0 * (3.0 * m[I003] + m[I201] + m[I021]));
    o[21] = o[22] = o[23] = 0.f;
    if (order >= 4) {
        for (int t = 0; t < 15; ++t) o[24 + t] = (float)m[I400 + t];
        o[39] = (float)(12.0 * m[I400] + 2.0 * m[I220] + 2.0 * m[I202]);
        o[40] = (float)(12.0 * m[I040] + 2.0 * m[I220] + 2.0 * m[I022]);
        o[41] = (float)(12.0 * m[I004] + 2.0 * m[I202] + 2.0 * m[I022]);
        o[42] = (float)(6.0 * m[I310] + 6.0 * m[I130] + 2.0 * m[I112]);
        o[43] = (float)(6.0 * m[I301] + 6.0 * m[I103] + 2.0 * m[I121]);
        o[44] = (float)(6.0 * m[I031] + 6.0 * m[I013] + 2.0 * m[I211]);
        o[45] = (float)(24.0 * (m[I400] + m[I040] + m[I004]) + 8.0 * (m[I220] + m[I202] + m[I022]));
        o[46] = o[47] = 0.f;
    }
    if (order >= 5) {
        for (int t = 0; t < 21; ++t) o[48 + t] = (float)m[I500 + t];
        const double x3 = 20.0 * m[I500] + 2.0 * m[I320] + 2.0 * m[I302];
        const double y3 = 20.0 * m[I050] + 2.0 * m[I230] + 2.0 * m[I032];
        const double z3 = 20.0 * m[I005] + 2.0 * m[I203] + 2.0 * m[I023];
        const double x2y = 12.0 * m[I410] + 6.0 * m[I230] + 2.0 * m[I212];
        const double x2z = 12.0 * m[I401] + 6.0 * m[I203] + 2.0 * m[I221];
        const double xy2 = 12.0 * m[I140] + 6.0 * m[I320] + 2.0 * m[I122];
        const double xz2 = 12.0 * m[I104] + 6.0 * m[I302] + 2.0 * m[I122];
        const double y2z = 12.0 * m[I041] + 6.0 * m[I023] + 2.0 * m[I221];
        const double yz2 = 12.0 * m[I014] + 6.0 * m[I032] + 2.0 * m[I212];
        const double xyz = 6.0 * (m[I311] + m[I131] + m[I113]);
        o[69] = (float)x3; o[70] = (float)y3; o[71] = (float)z3; o[72] = (float)x2y; o[73] = (float)x2z;
        o[74] = (float)xy2; o[75] = (float)xz2; o[76] = (float)y2z; o[77] = (float)yz2; o[78] = (float)xyz;
        o[79] = (float)(6.0 * x3 + 2.0 * xy2 + 2.0 * xz2);
        o[80] = (float)(6.0 * y3 + 2.0 * x2y + 2.0 * yz2);
        o[81] = (float)(6.0 * z3 + 2.0 * x2z + 2.0 * y2z);
        o[82] = o[83] = 0.f;
    }
}


// One thread per node of one level (deepest level first). Children are contiguous reference ids in
// octant order, so the sums run in the reference's order (tree.rs:876-929, 945-962, 1023-1063).
struct PayloadArgs {
    const int32_t* ids; int64_t count;
    const uint32_t* start; const uint32_t* pcount; const uint8_t* nchild; const int32_t* first_subnode;
    const double* spos; const double* smass; const double* sh;
    double* nmass; double* ncom; double* hmax; double* moments; int order; int ncoef;
    const int32_t* dfs; float* moments32; int rec32;  // fp32 walk records, written in depth-first order
};
// Leaves: one launch over ALL nodes (P2M has no dependencies), one thread per node, internal nodes return at once.
// Mass / COM / hmax in leaf-list order (tree.rs:881-905, 947-952), then P2M about the COM (tree.rs:1030-1042).
template <int ORDER>  // effective order: 0 (monopole storage, multipole_order <= 1), 2, 3, 4, 5
__global__ void __launch_bounds__(128) payload_leaves(PayloadArgs a) {
    constexpr int NC = mp::stored_coeffs(ORDER);
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= a.count || a.nchild[id] != 0) return;
    double mass = 0.0, cx = 0.0, cy = 0.0, cz = 0.0, hm = 0.0;
    const uint32_t s0 = a.start[id], c = a.pcount[id];
    for (uint32_t s = s0; s < s0 + c; ++s) {
        const double px = a.spos[3 * (int64_t)s], py = a.spos[3 * (int64_t)s + 1], pz = a.spos[3 * (int64_t)s + 2];
        if (a.smass) {
            const double m = a.smass[s];
            mass = __dadd_rn(mass, m);
            cx = __dadd_rn(cx, __dmul_rn(px, m));
            cy = __dadd_rn(cy, __dmul_rn(py, m));
            cz = __dadd_rn(cz, __dmul_rn(pz, m));
        } else {
            mass = __dadd_rn(mass, 1.0);
            cx = __dadd_rn(cx, px);
            cy = __dadd_rn(cy, py);
            cz = __dadd_rn(cz, pz);
        }
        if (a.hmax) hm = fmax(hm, fmax(a.sh[s], 0.0));
    }
    if (mass > 0.0) {
        cx = __ddiv_rn(cx, mass);
        cy = __ddiv_rn(cy, mass);
        cz = __ddiv_rn(cz, mass);
    }
    a.nmass[id] = mass;
    a.ncom[3 * id] = cx; a.ncom[3 * id + 1] = cy; a.ncom[3 * id + 2] = cz;
    if (a.hmax) a.hmax[id] = hm;
    double mom[NC];
#pragma unroll
    for (int t = 0; t < NC; ++t) mom[t] = 0.0;
    if (mass != 0.0) {
        for (uint32_t s = s0; s < s0 + c; ++s) {
            const double m = a.smass ? a.smass[s] : 1.0;
            mp::p2m_accumulate_ct<NC>(mom, m, __dsub_rn(a.spos[3 * (int64_t)s], cx), __dsub_rn(a.spos[3 * (int64_t)s + 1], cy),
                                      __dsub_rn(a.spos[3 * (int64_t)s + 2], cz));
        }
    }
#pragma unroll
    for (int t = 0; t < NC; ++t) a.moments[id * NC + t] = mom[t];
    pack_walk_record(mom, a.order, a.moments32 + (int64_t)a.dfs[id] * a.rec32);
}

// Internal nodes of one level, 8 lanes per node (lane l <-> child slot l): every lane loads its child's mass / COM /
// hmax, all lanes replay the reference's sequential sums over the children in octant order (identical bits on every
// lane, tree.rs:907-925, 953-961), each lane translates ITS child's moments to the node's COM (the expensive M2M, now
// 8-way parallel), and the lanes then add the translated sets coefficient-parallel, again in octant order
// (tree.rs:1046-1061) — same operations in the same order as a serial sweep, bit-equal results.
template <int ORDER>
__global__ void __launch_bounds__(mp::stored_coeffs(ORDER) > 35 ? 64 : 128) payload_internal(PayloadArgs a) {
    constexpr int NC = mp::stored_coeffs(ORDER);
    constexpr int GROUPS = NC > 35 ? 8 : 16;  // nodes per block
    __shared__ double s_tr[GROUPS][8][NC];
    const int lane8 = threadIdx.x & 7, g = threadIdx.x >> 3;
    const int64_t j = (int64_t)blockIdx.x * GROUPS + g;
    int32_t id = -1;
    int nc = 0;
    if (j < a.count) {
        id = a.ids[j];
        nc = a.nchild[id];
    }
    const bool node_ok = nc > 0;  // leaves were handled by payload_leaves<ORDER>
    const int32_t c0 = node_ok ? a.first_subnode[id] : 0;
    const bool have = node_ok && lane8 < nc;
    const int32_t c = c0 + lane8;
    double cm = 0.0, ccx = 0.0, ccy = 0.0, ccz = 0.0, chm = 0.0;
    if (have) {
        cm = a.nmass[c];
        ccx = a.ncom[3 * (int64_t)c]; ccy = a.ncom[3 * (int64_t)c + 1]; ccz = a.ncom[3 * (int64_t)c + 2];
        if (a.hmax) chm = a.hmax[c];
    }
    double mass = 0.0, cx = 0.0, cy = 0.0, cz = 0.0, hm = 0.0;
    unsigned nzmask = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double mk = __shfl_sync(0xffffffffu, cm, k, 8);
        const double xk = __shfl_sync(0xffffffffu, ccx, k, 8), yk = __shfl_sync(0xffffffffu, ccy, k, 8),
                     zk = __shfl_sync(0xffffffffu, ccz, k, 8), hk = __shfl_sync(0xffffffffu, chm, k, 8);
        if (k < nc) {
            hm = fmax(hm, hk);
            if (mk != 0.0) {
                nzmask |= 1u << k;
                mass = __dadd_rn(mass, mk);
                cx = __dadd_rn(cx, __dmul_rn(xk, mk));
                cy = __dadd_rn(cy, __dmul_rn(yk, mk));
                cz = __dadd_rn(cz, __dmul_rn(zk, mk));
            }
        }
    }
    if (mass > 0.0) {
        cx = __ddiv_rn(cx, mass);
        cy = __ddiv_rn(cy, mass);
        cz = __ddiv_rn(cz, mass);
    }
    if (node_ok && lane8 == 0) {
        a.nmass[id] = mass;
        a.ncom[3 * (int64_t)id] = cx; a.ncom[3 * (int64_t)id + 1] = cy; a.ncom[3 * (int64_t)id + 2] = cz;
        if (a.hmax) a.hmax[id] = hm;
    }
    if (have && cm != 0.0 && mass != 0.0) {
        double tr[NC];
#pragma unroll
        for (int t = 0; t < NC; ++t) tr[t] = 0.0;
        const double shift[3] = {__dsub_rn(cx, ccx), __dsub_rn(cy, ccy), __dsub_rn(cz, ccz)};
        mp::m2m_accumulate_ct<ORDER, NC>(tr, a.moments + (int64_t)c * NC, shift);
#pragma unroll
        for (int t = 0; t < NC; ++t) s_tr[g][lane8][t] = tr[t];
    }
    __syncwarp();
    if (node_ok) {
        for (int t = lane8; t < NC; t += 8) {
            double acc = 0.0;
            if (mass != 0.0)
                for (int k = 0; k < nc; ++k)
                    if (nzmask & (1u << k)) acc = __dadd_rn(acc, s_tr[g][k][t]);
            a.moments[(int64_t)id * NC + t] = acc;
        }
    }
    __syncwarp();  // the group's coefficient stores are visible to its lane 0
    if (node_ok && lane8 == 0) {
        double mom[NC];
#pragma unroll
        for (int t = 0; t < NC; ++t) mom[t] = a.moments[(int64_t)id * NC + t];
        pack_walk_record(mom, a.order, a.moments32 + (int64_t)a.dfs[id] * a.rec32);
    }
}

__global__ void build_walk_records(const double* __restrict__ nmass, const double* __restrict__ ncom,
                                   const double* __restrict__ half, const double* __restrict__ hmax, double csep,
                                   const uint8_t* __restrict__ nchild, const uint32_t* __restrict__ start,
                                   const uint32_t* __restrict__ count, const int32_t* __restrict__ first_subnode,
                                   const int32_t* __restrict__ next_branch, const int32_t* __restrict__ dfs, int64_t nn,
                                   NodeRec* __restrict__ rec) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    NodeRec r;
    r.com[0] = ncom[3 * i]; r.com[1] = ncom[3 * i + 1]; r.com[2] = ncom[3 * i + 2];
    const double s = half[i] * 2.0;
    r.size2 = __dmul_rn(s, s);
    r.gate2 = 0.0;
    if (hmax) {
        const double ch = __dmul_rn(csep, fmax(hmax[i], 0.0));
        r.gate2 = __dmul_rn(ch, ch);
    }
    r.next_branch = next_branch[i] < 0 ? -1 : dfs[next_branch[i]];
    if (nmass[i] == 0.0) { r.kind = -2; r.first = -1; }          // tree.rs:1087-1090
    else if (nchild[i] == 0) { r.kind = (int32_t)count[i]; r.first = (int32_t)start[i]; }
    else { r.kind = -1; r.first = dfs[first_subnode[i]]; }  // == dfs[i] + 1
    r.nleaf = 1;
    r.ref = (int32_t)i;
    r.pad_ = 0;
    rec[dfs[i]] = r;
}
// Leaf runs (tree.cuh) + the fp32 walk sources. One thread per sibling block (= per internal node, plus the root when
// the root itself is a leaf): merges every maximal run of consecutive non-zero-mass leaf children into the record of
// the run's first leaf and writes the run's particles relative to the run's origin (float64 subtraction, then the
// cast: close pairs keep their separation wherever the run sits in the box).
__global__ void merge_leaf_runs(const uint8_t* __restrict__ nchild, const int32_t* __restrict__ first_subnode,
                                const uint32_t* __restrict__ start, const uint32_t* __restrict__ count,
                                const int32_t* __restrict__ next_branch, const double* __restrict__ nmass,
                                const double* __restrict__ ncom, const double* __restrict__ spos,
                                const double* __restrict__ smass, const int32_t* __restrict__ dfs, int64_t nn,
                                NodeRec* __restrict__ rec, float4* __restrict__ src32) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nn) return;
    int64_t c0, c1;  // the sibling block [c0, c1)
    if (nchild[p] != 0) { c0 = first_subnode[p]; c1 = c0 + nchild[p]; }
    else if (p == 0) { c0 = 0; c1 = 1; }
    else return;
    auto is_leaf = [&](int64_t j) { return nchild[j] == 0 && nmass[j] != 0.0; };
    int64_t j = c0;
    while (j < c1) {
        if (!is_leaf(j)) { ++j; continue; }
        int64_t e = j + 1;
        while (e < c1 && is_leaf(e)) ++e;
        const int nl = (int)(e - j);
        double ox = 0.0, oy = 0.0, oz = 0.0;
        uint32_t total = 0;
        for (int64_t k = j; k < e; ++k) {
            ox += ncom[3 * k]; oy += ncom[3 * k + 1]; oz += ncom[3 * k + 2];
            total += count[k];
        }
        ox /= nl; oy /= nl; oz /= nl;
        if (nl > 1) {
            NodeRec r = rec[dfs[j]];  // sibling leaves are consecutive in depth-first order too (no descendants between)
            r.com[0] = ox; r.com[1] = oy; r.com[2] = oz;
            r.kind = (int32_t)total;
            r.next_branch = next_branch[e - 1] < 0 ? -1 : dfs[next_branch[e - 1]];
            r.nleaf = nl;
            rec[dfs[j]] = r;
        }
        if (src32) {
            const uint32_t s0 = start[j];
            for (uint32_t s = s0; s < s0 + total; ++s)
                src32[s] = make_float4((float)(spos[3 * (int64_t)s] - ox), (float)(spos[3 * (int64_t)s + 1] - oy),
                                       (float)(spos[3 * (int64_t)s + 2] - oz), smass ? (float)smass[s] : 1.0f);
        }
        j = e;
    }
}
__global__ void update_gates(const double* __restrict__ hmax, double csep, const int32_t* __restrict__ dfs, int64_t nn,
                             NodeRec* __restrict__ rec) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const double ch = __dmul_rn(csep, fmax(hmax[i], 0.0));
    rec[dfs[i]].gate2 = __dmul_rn(ch, ch);
}
inline unsigned nblk(int64_t n, int t = 256) { return (unsigned)std::max<int64_t>(1, ceil_div(n, t)); }

int bits_for(uint64_t v) {
    int b = 1;
    while (b < 64 && (v >> b)) ++b;
    return b;
}

// stable radix sort of (key, value) pairs on bits [0, end_bit)
template <class K>
void sort_pairs(const K* kin, K* kout, const uint32_t* vin, uint32_t* vout, int64_t n, int end_bit, cudaStream_t s) {
    size_t bytes = 0;
    PNBX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)n, 0, end_bit, s));
    DevBuf<uint8_t> tmp(bytes, s);
    PNBX_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(), bytes, kin, kout, vin, vout, (int)n, 0, end_bit, s));
    ++launch_counter();
}
void exclusive_sum(const int32_t* in, int32_t* out, int64_t n, cudaStream_t s) {
    size_t bytes = 0;
    PNBX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, s));
    DevBuf<uint8_t> tmp(bytes, s);
    PNBX_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), bytes, in, out, (int)n, s));
    ++launch_counter();
}
void inclusive_sum_u32(const uint32_t* in, uint32_t* out, int64_t n, cudaStream_t s) {
    size_t bytes = 0;
    PNBX_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, (int)n, s));
    DevBuf<uint8_t> tmp(bytes, s);
    PNBX_CUDA(cub::DeviceScan::InclusiveSum(tmp.get(), bytes, in, out, (int)n, s));
    ++launch_counter();
}

// Builds the topology from sorted keys. Returns false if `max_level` was reached with an over-full cell.
// `scratch` holds the n-sized temporaries (laid out by the caller together with the keys), `s2` the node-sized ones.
struct KeyScratch {
    uint64_t* skhi; uint64_t* sklo;  // sorted keys
    int8_t* bq; int8_t* Dq; int32_t* nodes_here; int32_t* base;
};
bool build_topology(pnbx_tree_impl& t, cudaStream_t s, const KeyScratch& k, int max_level, StageTimer& tm) {
    const int64_t n = t.n;
    const int64_t cap = t.leaf_capacity;
    const double* root4 = t.root4.p;
    const uint64_t* skhi = k.skhi;
    const uint64_t* sklo = k.sklo;
    FinalArrays f{};
    auto layout_final = [&](int64_t nn) {
        t.nn = nn;
        for (int pass = 0; pass < 2; ++pass) {
            Arena& A = t.a_topo;
            A.take(t.center, (size_t)3 * nn); A.take(t.half, (size_t)nn); A.take(t.node_depth, (size_t)nn);
            A.take(t.node_start, (size_t)nn); A.take(t.node_count, (size_t)nn);
            A.take(t.first_subnode, (size_t)nn); A.take(t.next_branch, (size_t)nn);
            A.take(t.path_hi, (size_t)nn); A.take(t.path_lo, (size_t)nn);
            A.take(t.node_nchild, (size_t)nn); A.take(t.level_ids, (size_t)nn); A.take(t.dfs_of_ref, (size_t)nn);
            if (pass == 0) A.commit(s);
        }
        f = FinalArrays{t.center.p, t.half.p, t.node_depth.p, t.node_start.p, t.node_count.p, t.first_subnode.p,
                        t.next_branch.p, t.path_hi.p, t.path_lo.p, t.node_nchild.p};
    };
    if (n == 0) {  // the reference's empty tree: one root leaf (tree.rs:658-734 with no points)
        layout_final(1);
        PNBX_LAUNCH(empty_root, 1, 1, 0, s, root4, f);
        PNBX_CUDA(cudaMemsetAsync(t.dfs_of_ref.p, 0, 4, s));
        t.n_leaves = 1; t.depth = 0; t.n_internal = 0;
        t.ilevel_off.assign(2, 0);
        return true;
    }
    tm.begin("octree.leaf_levels");
    DevBuf<int8_t> Pm, Sm;
    DevBuf<unsigned long long> hist(2 * MAX_LEVELS + 2, s);
    PNBX_CUDA(cudaMemsetAsync(hist.p, 0, hist.bytes(), s));
    const int64_t nw = n - cap;
    if (cap > LEAF_SORT_MAX && nw > 0) {
        // sliding-window maximum of W over windows of cap+1 entries (van Herk / Gil-Werman): prefix maxima P and
        // suffix maxima S inside blocks of cap+1 entries; any window is covered by one S and one P entry
        DevBuf<int8_t> W((size_t)nw, s);
        Pm.alloc((size_t)nw, s); Sm.alloc((size_t)nw, s);
        PNBX_LAUNCH(window_digits, nblk(nw), 256, 0, s, skhi, sklo, nw, cap, W.p);
        auto kf = thrust::make_transform_iterator(thrust::make_counting_iterator<int64_t>(0), BlockOf{cap + 1});
        auto kr = thrust::make_transform_iterator(thrust::make_counting_iterator<int64_t>(0), RevBlockOf{cap + 1, nw - 1});
        size_t bytes = 0;
        PNBX_CUDA(cub::DeviceScan::InclusiveScanByKey(nullptr, bytes, kf, W.p, Pm.p, MaxI8{}, (int)nw, cub::Equality{}, s));
        DevBuf<uint8_t> tmp(bytes, s);
        PNBX_CUDA(cub::DeviceScan::InclusiveScanByKey(tmp.get(), bytes, kf, W.p, Pm.p, MaxI8{}, (int)nw, cub::Equality{}, s));
        auto rin = thrust::make_reverse_iterator(W.p + nw);
        auto rout = thrust::make_reverse_iterator(Sm.p + nw);
        size_t bytes2 = 0;
        PNBX_CUDA(cub::DeviceScan::InclusiveScanByKey(nullptr, bytes2, kr, rin, rout, MaxI8{}, (int)nw, cub::Equality{}, s));
        DevBuf<uint8_t> tmp2(bytes2, s);
        PNBX_CUDA(cub::DeviceScan::InclusiveScanByKey(tmp2.get(), bytes2, kr, rin, rout, MaxI8{}, (int)nw, cub::Equality{}, s));
        launch_counter() += 2;
    }
    PNBX_LAUNCH(leaf_levels, nblk(n), 256, 0, s, skhi, sklo, n, cap, max_level, Pm.p, Sm.p, k.bq, k.Dq, k.nodes_here, hist.p);
    PNBX_CUDA(cudaMemsetAsync(k.base + n, 0, 4, s));
    exclusive_sum(k.nodes_here, k.base, n, s);
    // the one synchronisation of the build: per-level node counts (allocation sizes) and the overflow flag
    unsigned long long hh[2 * MAX_LEVELS + 2];
    PNBX_CUDA(cudaMemcpyAsync(hh, hist.p, sizeof(hh), cudaMemcpyDeviceToHost, s));
    PNBX_CUDA(cudaStreamSynchronize(s));
    tm.end();
    if (hh[2 * MAX_LEVELS]) return false;
    int64_t nn = 0, ni = 0;
    int depth = 0;
    t.ilevel_off.assign(1, 0);
    for (int L = 0; L < MAX_LEVELS; ++L) {
        nn += (int64_t)hh[L];
        if (hh[L]) depth = L;
    }
    for (int L = 0; L <= depth; ++L) { ni += (int64_t)hh[MAX_LEVELS + L]; t.ilevel_off.push_back(ni); }
    if (nn >= ((int64_t)1 << 31) - 16) throw ArgError{PNBX_ERR_ARG, "octree has more than 2^31 nodes"};

    tm.begin("octree.emit_nodes");
    layout_final(nn);
    Arena S2;  // node-sized temporaries: the depth-first node list and the renumbering arrays
    DfsNodes d{};
    int32_t *nchild = nullptr, *scan = nullptr, *ref = nullptr, *sortval = nullptr;
    uint8_t *sortkey = nullptr, *key_out = nullptr, *sort_tmp = nullptr;
    size_t sort_bytes = 0;
    PNBX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, sortkey, key_out, sortval, t.level_ids.p, (int)nn, 0, 7, s));
    for (int pass = 0; pass < 2; ++pass) {
        d.start = S2.take<uint32_t>((size_t)nn); d.count = S2.take<uint32_t>((size_t)nn);
        d.parent = S2.take<uint32_t>((size_t)nn); d.mask = S2.take<uint32_t>((size_t)nn);
        d.level = S2.take<uint8_t>((size_t)nn); d.digit = S2.take<uint8_t>((size_t)nn);
        d.center = S2.take<double>((size_t)3 * nn); d.half = S2.take<double>((size_t)nn);
        d.phi = S2.take<uint64_t>((size_t)nn); d.plo = S2.take<uint64_t>((size_t)nn);
        nchild = S2.take<int32_t>((size_t)nn); scan = S2.take<int32_t>((size_t)nn); ref = S2.take<int32_t>((size_t)nn);
        sortval = S2.take<int32_t>((size_t)nn); sortkey = S2.take<uint8_t>((size_t)nn); key_out = S2.take<uint8_t>((size_t)nn);
        sort_tmp = S2.take<uint8_t>(sort_bytes);
        if (pass == 0) S2.commit(s);
    }
    PNBX_CUDA(cudaMemsetAsync(d.mask, 0, (size_t)nn * 4, s));
    PNBX_LAUNCH(emit_nodes, nblk(n, 128), 128, 0, s, skhi, sklo, n, k.bq, k.Dq, k.nodes_here, k.base, root4, d);
    tm.end();

    // ---- reference numbering, links, scatter
    tm.begin("octree.renumber_links");
    PNBX_LAUNCH(child_counts, nblk(nn), 256, 0, s, d.mask, nn, nchild);
    exclusive_sum(nchild, scan, nn, s);
    PNBX_LAUNCH(assign_ref_ids, nblk(nn), 256, 0, s, d.parent, d.digit, d.mask, scan, nn, ref);
    PNBX_LAUNCH(scatter_nodes, nblk(nn), 256, 0, s, d, scan, ref, k.base, nn, n, f, sortkey, sortval, t.dfs_of_ref.p);
    // internal nodes grouped by level (bottom-up payload sweeps): stable 7-bit sort of the DFS list
    PNBX_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, sortkey, key_out, sortval, t.level_ids.p, (int)nn, 0, 7, s));
    ++launch_counter();
    t.depth = depth;
    t.n_internal = ni;
    t.n_leaves = nn - ni;
    tm.end();

    // ---- ascending original index inside every leaf
    tm.begin("octree.leaf_order");
    if (cap <= LEAF_SORT_MAX) {
        PNBX_LAUNCH(sort_leaves_small, nblk(n, 128), 128, 0, s, k.nodes_here, n, t.perm.p);
    } else {
        DevBuf<uint32_t> lflag((size_t)n, s), lord_sorted((size_t)n, s), lord_orig((size_t)n, s), lord_out((size_t)n, s);
        DevBuf<uint32_t> iota((size_t)n, s), perm2((size_t)n, s);
        PNBX_LAUNCH(head_flags, nblk(n), 256, 0, s, k.nodes_here, n, lflag.p);
        inclusive_sum_u32(lflag.p, lord_sorted.p, n, s);
        PNBX_LAUNCH(leaf_ordinal_to_particles, nblk(n), 256, 0, s, lord_sorted.p, t.perm.p, n, lord_orig.p);
        PNBX_LAUNCH(iota_u32, nblk(n), 256, 0, s, iota.p, n);
        sort_pairs<uint32_t>(lord_orig.p, lord_out.p, iota.p, perm2.p, n, bits_for((uint64_t)t.n_leaves + 1), s);
        t.perm = std::move(perm2);
    }
    tm.end();
    PNBX_CUDA(cudaGetLastError());
    return true;
}

void gather_sorted_sources(pnbx_tree_impl& t, cudaStream_t s) {
    const int64_t n = t.n;
    if (n == 0) return;
    PNBX_LAUNCH(gather_sources, nblk(n), 256, 0, s, t.pos.p, t.has_mass ? t.mass.p : nullptr, t.perm.p, n, t.spos.p,
                t.has_mass ? t.smass.p : nullptr);
}
void gather_sorted_soft(pnbx_tree_impl& t, cudaStream_t s) {
    const int64_t n = t.n;
    if (!t.has_h || n == 0) return;
    if (!t.sh.p) { t.sh.alloc((size_t)n, s); t.sh32.alloc((size_t)n + 4, s); }  // softenings given after the build
    PNBX_LAUNCH(gather_soft, nblk(n), 256, 0, s, t.h.p, t.perm.p, n, t.sh.p, t.sh32.p);
}

thread_local cudaEvent_t tl_order_event[KernelEvents::MAXDEV] = {};
cudaEvent_t order_event(int device) {
    cudaEvent_t& e = tl_order_event[device < KernelEvents::MAXDEV ? device : 0];
    if (!e) PNBX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return e;
}

}  // namespace

// `waiter` waits for everything queued on `on` so far. The per-thread event may be re-recorded right away: a wait
// captures the record that preceded it.
void stream_wait_stream(cudaStream_t waiter, cudaStream_t on, int device) {
    if (waiter == on) return;
    cudaEvent_t e = order_event(device);
    PNBX_CUDA(cudaEventRecord(e, on));
    PNBX_CUDA(cudaStreamWaitEvent(waiter, e, 0));
}
void tree_mark_ready(pnbx_tree_impl& t) { PNBX_CUDA(cudaEventRecord(t.ready, t.stream)); }
void tree_begin_use(const pnbx_tree_impl& t, cudaStream_t s) {
    if (s != t.stream) PNBX_CUDA(cudaStreamWaitEvent(s, t.ready, 0));
}
void tree_end_use(const pnbx_tree_impl& t, cudaStream_t s) { stream_wait_stream(t.stream, s, t.device); }

pnbx_tree_impl::~pnbx_tree_impl() {
    replicas.clear();  // each on its own device
    cudaSetDevice(device);
    // everything is freed on the tree's own stream, which is ordered after every evaluation (tree_end_use)
    DevBuf<double>* d64[] = {&pos, &mass, &h, &spos, &smass, &sh, &center, &half, &nmass, &ncom, &hmax, &moments, &root4};
    for (auto* b : d64) b->release();
    key_hi.release(); key_lo.release(); perm.release(); src32.release(); sh32.release();
    node_depth.release(); node_nchild.release(); node_start.release(); node_count.release();
    first_subnode.release(); next_branch.release(); path_hi.release(); path_lo.release(); level_ids.release();
    dfs_of_ref.release();
    moments32.release(); rec.release();
    a_payload.release(); a_topo.release(); a_src.release();
    if (ready) cudaEventDestroy(ready);
}

// One library-owned stream per device carries the memory traffic (allocate, build, free) of EVERY tree on that
// device. Blocks freed by one tree are then reusable by the next build at once (same-stream reuse in the pool; with
// a stream per tree the pool fell back to fresh driver allocations, 10-70 ms per GB-sized slab).
cudaStream_t device_tree_stream(int device) {
    static std::mutex mu;
    static cudaStream_t streams[KernelEvents::MAXDEV] = {};
    if (device < 0 || device >= KernelEvents::MAXDEV) throw ArgError{PNBX_ERR_ARG, "device ordinal out of range"};
    std::lock_guard<std::mutex> lock(mu);
    if (!streams[device]) PNBX_CUDA(cudaStreamCreateWithFlags(&streams[device], cudaStreamNonBlocking));
    return streams[device];
}

void tree_init(pnbx_tree_impl& t, int device, int64_t n, int64_t leaf_capacity, int multipole_order, int kernel,
               bool has_mass, bool has_h) {
    t.device = device;
    t.n = n;
    t.leaf_capacity = std::max<int64_t>(leaf_capacity, 1);  // tree.rs:701
    t.order_raw = multipole_order;
    t.order = std::min(multipole_order, 5);                 // tree.rs:1020
    t.kernel = kernel;
    t.has_mass = has_mass;
    t.has_h = has_h;
    t.stream = device_tree_stream(device);
    PNBX_CUDA(cudaEventCreateWithFlags(&t.ready, cudaEventDisableTiming));
    const size_t m = (size_t)std::max<int64_t>(n, 1);
    for (int pass = 0; pass < 2; ++pass) {
        Arena& A = t.a_src;
        A.take(t.root4, 4);
        A.take(t.pos, 3 * m); A.take(t.mass, m);
        if (has_h) { A.take(t.h, m); A.take(t.sh, m); A.take(t.sh32, m + 4); }
        A.take(t.key_hi, m); A.take(t.perm, m);
        A.take(t.spos, 3 * m); A.take(t.smass, m); A.take(t.src32, m);
        if (pass == 0) A.commit(t.stream);
    }
}

// tree.rs:968-1012
void tree_build_mass(pnbx_tree_impl& t, StageTimer& tm, bool regather) {
    cudaStream_t s = t.stream;
    const int64_t nn = t.nn;
    tm.begin("octree.payload.alloc_gather");
    if (regather) gather_sorted_sources(t, s);  // build_mass(): masses may have been replaced
    t.has_hmax = t.has_h;
    t.n_moments = mp::stored_coeffs(t.order);
    t.rec32 = mp::fast_rec_floats(t.order);
    for (int pass = 0; pass < 2; ++pass) {
        Arena& A = t.a_payload;
        if (pass == 0) A.release();  // rebuild: the old slab goes back to the pool in stream order
        A.take(t.nmass, (size_t)nn); A.take(t.ncom, (size_t)3 * nn);
        if (t.has_hmax) A.take(t.hmax, (size_t)nn); else t.hmax.release();
        A.take(t.moments, (size_t)nn * t.n_moments);
        A.take(t.rec, (size_t)nn);
        A.take(t.moments32, (size_t)nn * t.rec32);
        if (pass == 0) A.commit(s);
    }
    PayloadArgs a;
    a.start = t.node_start.p; a.pcount = t.node_count.p; a.nchild = t.node_nchild.p; a.first_subnode = t.first_subnode.p;
    a.spos = t.spos.p; a.smass = t.has_mass ? t.smass.p : nullptr; a.sh = t.has_h ? t.sh.p : nullptr;
    a.nmass = t.nmass.p; a.ncom = t.ncom.p; a.hmax = t.has_hmax ? t.hmax.p : nullptr; a.moments = t.moments.p;
    a.order = t.order; a.ncoef = t.n_moments;
    a.dfs = t.dfs_of_ref.p; a.moments32 = t.moments32.p; a.rec32 = t.rec32;
    const int eff = t.order <= 1 ? 0 : t.order;
    auto launch = [&](bool leaves) {
#define PNBX_P(O)                                                                                            \
    if (eff == O) {                                                                                          \
        if (leaves) PNBX_LAUNCH((payload_leaves<O>), nblk(a.count, 128), 128, 0, s, a);                      \
        else PNBX_LAUNCH((payload_internal<O>), nblk(a.count, mp::stored_coeffs(O) > 35 ? 8 : 16),            \
                         mp::stored_coeffs(O) > 35 ? 64 : 128, 0, s, a);                                     \
    }
        PNBX_P(0) PNBX_P(2) PNBX_P(3) PNBX_P(4) PNBX_P(5)
#undef PNBX_P
    };
    tm.end();
    tm.begin("octree.payload.leaves");
    a.ids = nullptr;
    a.count = nn;
    launch(true);  // every leaf of every level at once
    tm.end();
    tm.begin("octree.payload.internal");
    for (int d = (int)t.ilevel_off.size() - 2; d >= 0; --d) {  // internal nodes only, deepest level first
        a.ids = t.level_ids.p + t.ilevel_off[d];
        a.count = t.ilevel_off[d + 1] - t.ilevel_off[d];
        if (a.count > 0) launch(false);
    }
    tm.end();
    tm.begin("octree.payload.walk_records");
    PNBX_LAUNCH(build_walk_records, nblk(nn), 256, 0, s, t.nmass.p, t.ncom.p, t.half.p, t.has_hmax ? t.hmax.p : nullptr,
                t.kernel == PNBX_KERNEL_SPLINE ? 1.0 : 2.8, t.node_nchild.p, t.node_start.p, t.node_count.p,
                t.first_subnode.p, t.next_branch.p, t.dfs_of_ref.p, nn, t.rec.p);
    PNBX_LAUNCH(merge_leaf_runs, nblk(nn), 256, 0, s, t.node_nchild.p, t.first_subnode.p, t.node_start.p, t.node_count.p,
                t.next_branch.p, t.nmass.p, t.ncom.p, t.spos.p, t.has_mass ? t.smass.p : nullptr, t.dfs_of_ref.p, nn,
                t.rec.p, t.n > 0 ? t.src32.p : nullptr);
    PNBX_CUDA(cudaGetLastError());
    t.has_payload = true;
    tm.end();
}

void tree_build(pnbx_tree_impl& t, StageTimer& tm) {
    cudaStream_t s = t.stream;
    const int64_t n = t.n;
    if (n > 0) {
        tm.begin("octree.bbox");
        DevBuf<double> bb(6, s);
        launch_bbox(t.pos.p, n, bb.p, s);
        PNBX_LAUNCH(root_from_bbox, 1, 1, 0, s, bb.p, t.root4.p);
        tm.end();
    } else {
        // empty point set: the reference gets a NaN-centred, half = 1e-6 root leaf of mass 0
        static const double r4[4] = {0.0, 0.0, 0.0, 1e-6};
        PNBX_CUDA(cudaMemcpyAsync(t.root4.p, r4, sizeof(r4), cudaMemcpyHostToDevice, s));
    }
    // keys + sort; first with the 21-level word only, the second word only if a level-21 cell is over-full
    bool ok = false;
    for (int attempt = 0; attempt < 2 && !ok; ++attempt) {
        const bool two = attempt == 1;
        const int levels = two ? KEY_LEVELS : KEY_LEVELS_HI;
        const size_t m = (size_t)std::max<int64_t>(n, 1);
        Arena S1;  // n-sized temporaries
        KeyScratch k{};
        uint32_t *iota = nullptr, *idx1 = nullptr;
        uint64_t *tmpk = nullptr, *hi1 = nullptr;
        uint8_t* sort_tmp = nullptr;
        double4* packed = nullptr;
        size_t sort_bytes = 0;
        if (n > 0)
            PNBX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, k.skhi, k.skhi, iota, iota, (int)n, 0, 63, s));
        if (two) t.key_lo.alloc(m, s);
        for (int pass = 0; pass < 2; ++pass) {
            k.skhi = S1.take<uint64_t>(m);
            iota = S1.take<uint32_t>(m);
            sort_tmp = S1.take<uint8_t>(sort_bytes);
            k.bq = S1.take<int8_t>(m); k.Dq = S1.take<int8_t>(m);
            k.nodes_here = S1.take<int32_t>(m); k.base = S1.take<int32_t>(m + 1);
            packed = S1.take<double4>(m);
            if (two) {
                k.sklo = S1.take<uint64_t>(m); tmpk = S1.take<uint64_t>(m); hi1 = S1.take<uint64_t>(m);
                idx1 = S1.take<uint32_t>(m);
            }
            if (pass == 0) S1.commit(s);
        }
        if (n > 0) {
            tm.begin("octree.keys_sort");
            PNBX_LAUNCH(path_keys, nblk(n), 256, 0, s, t.pos.p, t.has_mass ? t.mass.p : nullptr, n, t.root4.p, levels,
                        t.key_hi.p, two ? t.key_lo.p : nullptr, packed);
            PNBX_LAUNCH(iota_u32, nblk(n), 256, 0, s, iota, n);
            auto sort63 = [&](const uint64_t* kin, uint64_t* kout, const uint32_t* vin, uint32_t* vout) {
                size_t bytes = sort_bytes;
                PNBX_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, bytes, kin, kout, vin, vout, (int)n, 0, 63, s));
                ++launch_counter();
            };
            if (!two) {
                sort63(t.key_hi.p, k.skhi, iota, t.perm.p);
            } else {  // LSD over the two words: low word first, then a stable sort by the high word
                sort63(t.key_lo.p, tmpk, iota, idx1);
                PNBX_LAUNCH(gather_u32<uint64_t>, nblk(n), 256, 0, s, t.key_hi.p, idx1, n, hi1);
                sort63(hi1, k.skhi, idx1, t.perm.p);
                PNBX_LAUNCH(gather_u32<uint64_t>, nblk(n), 256, 0, s, t.key_lo.p, t.perm.p, n, k.sklo);
            }
            tm.end();
        }
        ok = build_topology(t, s, k, levels, tm);
        if (ok && n > 0)  // sorted float64 copies, from the packed records while the scratch slab is alive
            PNBX_LAUNCH(gather_packed_sources, nblk(n), 256, 0, s, packed, t.perm.p, n, t.spos.p,
                        t.has_mass ? t.smass.p : nullptr);
    }
    if (!ok)
        throw ArgError{PNBX_ERR_DEPTH,
                       "octree deeper than 42 levels (more than leaf_capacity coincident or nearly coincident "
                       "points); the reference would recurse without bound here"};
    gather_sorted_soft(t, s);
    if (t.has_mass) tree_build_mass(t, tm, false);  // gravity.rs:210-220
}

}  // namespace pnbx

using namespace pnbx;

namespace {
Exec make_exec_for(int device, cudaStream_t s) {  // host-pointer context on an explicit device / stream (current device is set)
    Exec ex;
    ex.device = device;
    ex.stream = s;
    return ex;
}
pnbx_opts tree_opts(const pnbx_tree_impl& t, const pnbx_opts* opts) {
    pnbx_opts o = opts ? *opts : pnbx_opts{-1, PNBX_MEM_HOST, 0, 0, nullptr, 0, 1, 0};
    if (o.device < 0) o.device = t.device;
    if (o.device != t.device) throw ArgError{PNBX_ERR_ARG, "tree lives on a different device"};
    return o;
}
// copy an input array into a tree buffer on the tree's stream (host pointer: H2D, staged if pageable; device
// pointer: ordered after the caller's stream)
void copy_into_tree(pnbx_tree_impl& t, const Exec& ex, double* dst, const double* src, size_t count) {
    if (!count) return;
    if (ex.device_ptrs) {
        PNBX_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToDevice, t.stream));
    } else {
        Exec tex = ex;
        tex.stream = t.stream;
        copy_h2d(dst, src, count * sizeof(double), tex);
    }
}
// after a build / setter: publish, then either block (host arrays: the call is synchronous) or order the caller's
// stream behind the tree's (device arrays: the caller may reuse its inputs in stream order)
void tree_publish(pnbx_tree_impl& t, Exec& ex) {
    tree_mark_ready(t);
    if (ex.device_ptrs) PNBX_CUDA(cudaStreamWaitEvent(ex.stream, t.ready, 0));
    else PNBX_CUDA(cudaStreamSynchronize(t.stream));
    if (ex.own_stream) { cudaStreamDestroy(ex.stream); ex.own_stream = false; }
}
}  // namespace

extern "C" int pnbx_tree_create(pnbx_tree** out, const double* pos, const double* mass, const double* h, int64_t n,
                                int64_t leaf_capacity, int multipole_order, int kernel, const pnbx_opts* opts) {
    return guarded([&] {
        if (!out) throw ArgError{PNBX_ERR_ARG, "out is NULL"};
        *out = nullptr;
        if (n < 0) throw ArgError{PNBX_ERR_ARG, "negative size"};
        if (n > 0 && !pos) throw ArgError{PNBX_ERR_ARG, "positions must be (N,3) float64 array"};
        if (n >= ((int64_t)1 << 31) - 16) throw ArgError{PNBX_ERR_ARG, "N must be < 2^31"};
        if (kernel != PNBX_KERNEL_PLUMMER && kernel != PNBX_KERNEL_SPLINE)
            throw ArgError{PNBX_ERR_ARG, "kernel must be 0 (Plummer) or 1 (CubicSplineW2)"};
        if (leaf_capacity < 0 || multipole_order < 0) throw ArgError{PNBX_ERR_ARG, "negative leaf_capacity / multipole_order"};
        auto t = std::make_unique<pnbx_tree_impl>();
        // PNBX_DEVICES: host arrays, no explicit device -> identical trees on all listed GPUs (multi.cu)
        if ((!opts || (opts->mem_space == PNBX_MEM_HOST && opts->device < 0)) &&
            multi_tree_create(*t, pos, mass, h, n, leaf_capacity, multipole_order, kernel, opts)) {
            *out = reinterpret_cast<pnbx_tree*>(t.release());
            return;
        }
        Exec ex = make_exec(opts);
        tree_init(*t, ex.device, n, leaf_capacity, multipole_order, kernel, mass != nullptr, h != nullptr);
        StageTimer tm(t->stream);
        if (ex.device_ptrs) stream_wait_stream(t->stream, ex.stream, ex.device);  // inputs produced on the caller's stream
        tm.begin("octree.copy_in");
        copy_into_tree(*t, ex, t->pos.p, pos, (size_t)3 * n);
        if (mass) copy_into_tree(*t, ex, t->mass.p, mass, (size_t)n);
        if (h) copy_into_tree(*t, ex, t->h.p, h, (size_t)n);
        tm.end();
        tree_build(*t, tm);
        tree_publish(*t, ex);
        *out = reinterpret_cast<pnbx_tree*>(t.release());
    });
}

extern "C" int pnbx_tree_build_mass_ex(pnbx_tree* tp, const double* mass, const pnbx_opts* opts) {
    return guarded([&] {
        if (!tp) throw ArgError{PNBX_ERR_ARG, "tree is NULL"};
        auto& t = *reinterpret_cast<pnbx_tree_impl*>(tp);
        if (!t.replicas.empty()) {  // PNBX_DEVICES tree: host arrays only, replayed on every copy
            if (opts && opts->mem_space == PNBX_MEM_DEVICE)
                throw ArgError{PNBX_ERR_ARG, "a multi-device tree takes host arrays"};
            multi_tree_for_each(t, [&](pnbx_tree_impl& r) {
                Exec ex = make_exec_for(r.device, r.stream);
                StageTimer tm(r.stream);
                if (mass) { copy_into_tree(r, ex, r.mass.p, mass, (size_t)r.n); r.has_mass = true; }
                tree_build_mass(r, tm, true);
                tree_mark_ready(r);
                PNBX_CUDA(cudaStreamSynchronize(r.stream));
            });
            return;
        }
        pnbx_opts o = tree_opts(t, opts);
        Exec ex = make_exec(&o);
        StageTimer tm(t.stream);
        if (mass) {
            if (ex.device_ptrs) stream_wait_stream(t.stream, ex.stream, ex.device);
            copy_into_tree(t, ex, t.mass.p, mass, (size_t)t.n);
            t.has_mass = true;
        }
        tree_build_mass(t, tm, true);
        tree_publish(t, ex);
    });
}
extern "C" int pnbx_tree_build_mass(pnbx_tree* tp, const double* mass) { return pnbx_tree_build_mass_ex(tp, mass, nullptr); }

extern "C" int pnbx_tree_set_softenings_ex(pnbx_tree* tp, const double* h, const pnbx_opts* opts) {
    return guarded([&] {
        if (!tp) throw ArgError{PNBX_ERR_ARG, "tree is NULL"};
        auto& t = *reinterpret_cast<pnbx_tree_impl*>(tp);
        if (!t.replicas.empty()) {
            if (opts && opts->mem_space == PNBX_MEM_DEVICE)
                throw ArgError{PNBX_ERR_ARG, "a multi-device tree takes host arrays"};
            multi_tree_for_each(t, [&](pnbx_tree_impl& r) {
                Exec ex = make_exec_for(r.device, r.stream);
                if (h) {
                    if (!r.h.p) r.h.alloc((size_t)std::max<int64_t>(r.n, 1), r.stream);
                    copy_into_tree(r, ex, r.h.p, h, (size_t)r.n);
                    r.has_h = true;
                    gather_sorted_soft(r, r.stream);
                    PNBX_CUDA(cudaGetLastError());
                } else {
                    r.has_h = false;
                }
                tree_mark_ready(r);
                PNBX_CUDA(cudaStreamSynchronize(r.stream));
            });
            return;
        }
        pnbx_opts o = tree_opts(t, opts);
        Exec ex = make_exec(&o);
        if (h) {
            if (!t.h.p) t.h.alloc((size_t)std::max<int64_t>(t.n, 1), t.stream);  // tree built without softenings
            if (ex.device_ptrs) stream_wait_stream(t.stream, ex.stream, ex.device);
            copy_into_tree(t, ex, t.h.p, h, (size_t)t.n);
            t.has_h = true;
            gather_sorted_soft(t, t.stream);
            PNBX_CUDA(cudaGetLastError());
        } else {
            t.has_h = false;
        }
        // hmax is deliberately left as built (tree.rs:777-782 does not touch it)
        tree_publish(t, ex);
    });
}
extern "C" int pnbx_tree_set_softenings(pnbx_tree* tp, const double* h) { return pnbx_tree_set_softenings_ex(tp, h, nullptr); }

extern "C" int pnbx_tree_set_kernel(pnbx_tree* tp, int kernel) {
    return guarded([&] {
        if (!tp) throw ArgError{PNBX_ERR_ARG, "tree is NULL"};
        if (kernel != PNBX_KERNEL_PLUMMER && kernel != PNBX_KERNEL_SPLINE)
            throw ArgError{PNBX_ERR_ARG, "kernel must be 0 (Plummer) or 1 (CubicSplineW2)"};
        auto& primary = *reinterpret_cast<pnbx_tree_impl*>(tp);
        auto apply = [&](pnbx_tree_impl& t) {
            if (t.kernel != kernel && t.has_payload && t.has_hmax) {  // the gate factor c depends on the kernel (kernel.rs:20-28)
                PNBX_CUDA(cudaSetDevice(t.device));
                PNBX_LAUNCH(update_gates, nblk(t.nn), 256, 0, t.stream, t.hmax.p, kernel == PNBX_KERNEL_SPLINE ? 1.0 : 2.8,
                            t.dfs_of_ref.p, t.nn, t.rec.p);
                PNBX_CUDA(cudaGetLastError());
                tree_mark_ready(t);  // stream-ordered: evaluations wait for `ready`
            }
            t.kernel = kernel;
        };
        apply(primary);
        for (auto& r : primary.replicas) apply(*r);
    });
}

extern "C" void pnbx_tree_destroy(pnbx_tree* tp) {
    if (!tp) return;
    DeviceGuard restore_device;
    auto* t = reinterpret_cast<pnbx_tree_impl*>(tp);
    cudaSetDevice(t->device);
    delete t;
}

extern "C" int pnbx_tree_get_info(const pnbx_tree* tp, pnbx_tree_info* info) {
    return guarded([&] {
        if (!tp || !info) throw ArgError{PNBX_ERR_ARG, "NULL argument"};
        const auto& t = *reinterpret_cast<const pnbx_tree_impl*>(tp);
        info->n_particles = t.n;
        info->n_nodes = t.nn;
        info->n_leaves = t.n_leaves;
        info->depth = t.depth;
        info->multipole_order = t.order;
        info->n_moments = t.has_payload ? t.n_moments : 0;
        info->has_payload = t.has_payload;
        info->has_hmax = t.has_payload && t.has_hmax;
        info->kernel = t.kernel;
        info->leaf_capacity = t.leaf_capacity;
    });
}

namespace {
template <class T>
void d2h(T* dst, const T* src, size_t n, cudaStream_t s) {
    if (dst && n) PNBX_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, s));
}
}  // namespace

extern "C" int pnbx_tree_dump_topology(const pnbx_tree* tp, double* center, double* half, int32_t* depth,
                                       int64_t* first_subnode, int64_t* next_branch, int64_t* leaf_start,
                                       int64_t* leaf_count, int64_t* leaf_particles, uint64_t* path_hi,
                                       uint64_t* path_lo) {
    return guarded([&] {
        if (!tp) throw ArgError{PNBX_ERR_ARG, "tree is NULL"};
        const auto& t = *reinterpret_cast<const pnbx_tree_impl*>(tp);
        PNBX_CUDA(cudaSetDevice(t.device));
        cudaStream_t s = t.stream;  // ordered after the build and every setter
        const size_t nn = (size_t)t.nn;
        d2h(center, t.center.p, 3 * nn, s);
        d2h(half, t.half.p, nn, s);
        d2h(path_hi, t.path_hi.p, nn, s);
        d2h(path_lo, t.path_lo.p, nn, s);
        std::vector<uint8_t> dp(nn), nc(nn);
        std::vector<int32_t> fs(nn), nb(nn);
        std::vector<uint32_t> st(nn), ct(nn), pm((size_t)t.n);
        d2h(dp.data(), t.node_depth.p, nn, s);
        d2h(nc.data(), t.node_nchild.p, nn, s);
        d2h(fs.data(), t.first_subnode.p, nn, s);
        d2h(nb.data(), t.next_branch.p, nn, s);
        d2h(st.data(), t.node_start.p, nn, s);
        d2h(ct.data(), t.node_count.p, nn, s);
        d2h(pm.data(), t.perm.p, (size_t)t.n, s);
        PNBX_CUDA(cudaStreamSynchronize(s));
        for (size_t i = 0; i < nn; ++i) {
            if (depth) depth[i] = dp[i];
            if (first_subnode) first_subnode[i] = fs[i];
            if (next_branch) next_branch[i] = nb[i];
            const bool leaf = nc[i] == 0;
            if (leaf_start) leaf_start[i] = leaf ? (int64_t)st[i] : -1;
            if (leaf_count) leaf_count[i] = leaf ? (int64_t)ct[i] : -1;
        }
        if (leaf_particles)
            for (size_t k = 0; k < (size_t)t.n; ++k) leaf_particles[k] = pm[k];
    });
}

extern "C" int pnbx_tree_dump_payload(const pnbx_tree* tp, double* mass, double* com, double* hmax, double* moments) {
    return guarded([&] {
        if (!tp) throw ArgError{PNBX_ERR_ARG, "tree is NULL"};
        const auto& t = *reinterpret_cast<const pnbx_tree_impl*>(tp);
        if (!t.has_payload) throw ArgError{PNBX_ERR_STATE, "mass payload not built; call build_mass() first"};
        PNBX_CUDA(cudaSetDevice(t.device));
        const size_t nn = (size_t)t.nn;
        d2h(mass, t.nmass.p, nn, t.stream);
        d2h(com, t.ncom.p, 3 * nn, t.stream);
        if (t.has_hmax) d2h(hmax, t.hmax.p, nn, t.stream);
        d2h(moments, t.moments.p, nn * t.n_moments, t.stream);
        PNBX_CUDA(cudaStreamSynchronize(t.stream));
    });
}

extern "C" int pnbx_tree_dump_keys(const pnbx_tree* tp, uint64_t* key_hi, uint64_t* key_lo) {
    return guarded([&] {
        if (!tp) throw ArgError{PNBX_ERR_ARG, "tree is NULL"};
        const auto& t = *reinterpret_cast<const pnbx_tree_impl*>(tp);
        PNBX_CUDA(cudaSetDevice(t.device));
        d2h(key_hi, t.key_hi.p, (size_t)t.n, t.stream);
        if (t.key_lo.p) d2h(key_lo, t.key_lo.p, (size_t)t.n, t.stream);
        else if (key_lo) memset(key_lo, 0, (size_t)t.n * sizeof(uint64_t));  // 21 levels were enough: no second word
        PNBX_CUDA(cudaStreamSynchronize(t.stream));
    });
}
