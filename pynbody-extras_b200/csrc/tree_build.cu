// tree_build.cu — K2..K6: GPU construction of the reference's octree (tree.rs:628-1067).
//
//   K2 bbox            tree.rs:628-654   (launch_bbox + root_from_bbox, float64)
//   K3 path keys       tree.rs:815-838   float64 descent with the reference's own centre arithmetic
//   K4 radix sort      (cub::DeviceRadixSort, stable)  = the reference's stable octant bucketing
//   K5 node emission   tree.rs:804-864, 736-776: level-by-level from sorted key ranges, then the
//                      reference's creation-order numbering, first_subnode / next_branch links
//   K6 payloads        tree.rs:866-965, 1014-1067: mass/COM, hmax, P2M/M2M, bottom-up per level,
//                      float64 with the reference's operation order (no FMA contraction)
#include <cub/cub.cuh>

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
__global__ void path_keys(const double* __restrict__ pos, int64_t n, const double* __restrict__ root4, int levels,
                          uint64_t* __restrict__ key_hi, uint64_t* __restrict__ key_lo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
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
// BFS node arrays (struct of arrays, grown geometrically).
struct Bfs {
    DevBuf<uint32_t> start, count;
    DevBuf<int32_t> parent, child_base;
    DevBuf<uint8_t> rank, depth, nchild;
    DevBuf<double> center, half;  // (cap,3), (cap)
    DevBuf<uint64_t> path_hi, path_lo;
    int64_t cap = 0, size = 0;
    cudaStream_t s = nullptr;

    template <class T>
    static void grow(DevBuf<T>& b, int64_t old_elems, int64_t new_elems, cudaStream_t s) {
        DevBuf<T> nb((size_t)new_elems, s);
        if (old_elems) PNBX_CUDA(cudaMemcpyAsync(nb.p, b.p, (size_t)old_elems * sizeof(T), cudaMemcpyDeviceToDevice, s));
        b = std::move(nb);
    }
    void reserve(int64_t want) {
        if (want <= cap) return;
        int64_t nc = std::max<int64_t>(want, cap + cap / 2 + 1024);
        grow(start, size, nc, s); grow(count, size, nc, s); grow(parent, size, nc, s); grow(child_base, size, nc, s);
        grow(rank, size, nc, s); grow(depth, size, nc, s); grow(nchild, size, nc, s);
        grow(center, 3 * size, 3 * nc, s); grow(half, size, nc, s);
        grow(path_hi, size, nc, s); grow(path_lo, size, nc, s);
        cap = nc;
    }
};

__global__ void init_root(uint32_t* start, uint32_t* count, int32_t* parent, uint8_t* rank, uint8_t* depth,
                          double* center, double* half, uint64_t* phi, uint64_t* plo, const double* root4, uint32_t n) {
    start[0] = 0; count[0] = n; parent[0] = -1; rank[0] = 0; depth[0] = 0;
    center[0] = root4[0]; center[1] = root4[1]; center[2] = root4[2]; half[0] = root4[3];
    phi[0] = 0; plo[0] = 0;
}

__device__ __forceinline__ unsigned digit_at(const uint64_t* __restrict__ khi, const uint64_t* __restrict__ klo,
                                             uint32_t s, int level) {
    return level <= KEY_LEVELS_HI ? (unsigned)((khi[s] >> (3 * (KEY_LEVELS_HI - level))) & 7u)
                                  : (unsigned)((klo[s] >> (3 * (KEY_LEVELS - level))) & 7u);
}

// For every node of the current level that must split (count > leaf_capacity, tree.rs:848-851), find the 9
// octant boundaries inside its sorted range; one thread per (node, boundary).
__global__ void split_level(const uint32_t* __restrict__ start, const uint32_t* __restrict__ count, int64_t level_begin,
                            int64_t level_size, int child_level, uint32_t leaf_capacity,
                            const uint64_t* __restrict__ khi, const uint64_t* __restrict__ klo,
                            uint32_t* __restrict__ bounds /* level_size x 9 */, uint8_t* __restrict__ nchild,
                            int32_t* __restrict__ nchild_i32) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t j = t / 9;
    int b = (int)(t % 9);
    if (j >= level_size) return;
    const uint32_t s0 = start[level_begin + j], c = count[level_begin + j];
    if (c <= leaf_capacity) {
        if (b == 0) { nchild[level_begin + j] = 0; nchild_i32[j] = 0; }
        return;
    }
    // first position in [s0, s0+c) whose digit >= b
    uint32_t lo = s0, hi = s0 + c;
    if (b == 8) lo = hi;
    else if (b > 0) {
        while (lo < hi) {
            uint32_t mid = lo + ((hi - lo) >> 1);
            if (digit_at(khi, klo, mid, child_level) < (unsigned)b) lo = mid + 1;
            else hi = mid;
        }
    }
    bounds[j * 9 + b] = lo;
}
__global__ void count_children(const uint32_t* __restrict__ count, int64_t level_begin, int64_t level_size,
                               uint32_t leaf_capacity, const uint32_t* __restrict__ bounds, uint8_t* __restrict__ nchild,
                               int32_t* __restrict__ nchild_i32) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= level_size) return;
    if (count[level_begin + j] <= leaf_capacity) return;
    int k = 0;
    for (int o = 0; o < 8; ++o) k += bounds[j * 9 + o + 1] > bounds[j * 9 + o];
    nchild[level_begin + j] = (uint8_t)k;
    nchild_i32[j] = k;
}
// Children in octant order, contiguous per parent (tree.rs:830-843); centre = parent +/- half/2 with the
// reference's rounded additions.
__global__ void emit_children(uint32_t* __restrict__ start, uint32_t* __restrict__ count, int32_t* __restrict__ parent,
                              int32_t* __restrict__ child_base, uint8_t* __restrict__ rank, uint8_t* __restrict__ depth,
                              const uint8_t* __restrict__ nchild, double* __restrict__ center, double* __restrict__ half,
                              uint64_t* __restrict__ phi, uint64_t* __restrict__ plo, int64_t level_begin,
                              int64_t level_size, int64_t next_begin, const int32_t* __restrict__ child_off,
                              const uint32_t* __restrict__ bounds, int child_level) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= level_size) return;
    const int64_t p = level_begin + j;
    if (nchild[p] == 0) { child_base[p] = -1; return; }
    const int64_t base = next_begin + child_off[j];
    child_base[p] = (int32_t)base;
    const double cx = center[3 * p], cy = center[3 * p + 1], cz = center[3 * p + 2];
    const double off = half[p] * 0.5;
    int r = 0;
    for (int o = 0; o < 8; ++o) {
        const uint32_t a = bounds[j * 9 + o], b = bounds[j * 9 + o + 1];
        if (b <= a) continue;
        const int64_t c = base + r;
        start[c] = a;
        count[c] = b - a;
        parent[c] = (int32_t)p;
        rank[c] = (uint8_t)r;
        depth[c] = (uint8_t)child_level;
        center[3 * c + 0] = __dadd_rn(cx, (o & 1) ? off : -off);
        center[3 * c + 1] = __dadd_rn(cy, (o & 2) ? off : -off);
        center[3 * c + 2] = __dadd_rn(cz, (o & 4) ? off : -off);
        half[c] = off;
        uint64_t h = phi[p], l = plo[p];
        if (child_level <= KEY_LEVELS_HI) h |= (uint64_t)o << (3 * (KEY_LEVELS_HI - child_level));
        else l |= (uint64_t)o << (3 * (KEY_LEVELS - child_level));
        phi[c] = h;
        plo[c] = l;
        ++r;
    }
}

// ---- renumbering into the reference's creation order
__global__ void internal_flags_keys(const uint8_t* __restrict__ nchild, const uint32_t* __restrict__ start,
                                    const uint8_t* __restrict__ depth, int64_t nn, uint8_t* __restrict__ flag,
                                    uint64_t* __restrict__ key) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    flag[i] = nchild[i] > 0;
    key[i] = ((uint64_t)start[i] << 8) | depth[i];  // DFS pre-order of internal nodes = (start, depth) ascending
}
__global__ void gather_nchild(const uint8_t* __restrict__ nchild, const int32_t* __restrict__ idx, int64_t ni,
                              int32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ni) out[i] = nchild[idx[i]];
}
__global__ void scatter_first_child(const int32_t* __restrict__ idx, const int32_t* __restrict__ scan, int64_t ni,
                                    int32_t* __restrict__ first_child_ref) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ni) first_child_ref[idx[i]] = 1 + scan[i];  // root is node 0 (tree.rs:708-709)
}
__global__ void assign_ref_ids(const int32_t* __restrict__ parent, const uint8_t* __restrict__ rank,
                               const int32_t* __restrict__ first_child_ref, int64_t nn, int32_t* __restrict__ ref) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    ref[i] = i == 0 ? 0 : first_child_ref[parent[i]] + rank[i];
}
// next_branch (tree.rs:736-776) one level at a time, in BFS indexing
__global__ void links_level(const int32_t* __restrict__ parent, const uint8_t* __restrict__ rank,
                            const uint8_t* __restrict__ nchild, const int32_t* __restrict__ ref, int64_t level_begin,
                            int64_t level_size, int32_t* __restrict__ nb_bfs) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= level_size) return;
    const int64_t i = level_begin + j;
    if (i == 0) { nb_bfs[0] = -1; return; }
    const int32_t p = parent[i];
    nb_bfs[i] = (rank[i] + 1 < nchild[p]) ? ref[i] + 1 : nb_bfs[p];
}
struct FinalArrays {
    double* center; double* half; uint8_t* depth; uint32_t* start; uint32_t* count; int32_t* first_subnode;
    int32_t* next_branch; uint64_t* path_hi; uint64_t* path_lo; uint8_t* nchild; int32_t* level_ids;
};
__global__ void scatter_nodes(const uint32_t* __restrict__ start, const uint32_t* __restrict__ count,
                              const int32_t* __restrict__ parent, const uint8_t* __restrict__ depth,
                              const uint8_t* __restrict__ nchild, const double* __restrict__ center,
                              const double* __restrict__ half, const uint64_t* __restrict__ phi,
                              const uint64_t* __restrict__ plo, const int32_t* __restrict__ ref,
                              const int32_t* __restrict__ first_child_ref, const int32_t* __restrict__ nb_bfs, int64_t nn,
                              FinalArrays f) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const int32_t r = ref[i];
    f.center[3 * r] = center[3 * i]; f.center[3 * r + 1] = center[3 * i + 1]; f.center[3 * r + 2] = center[3 * i + 2];
    f.half[r] = half[i];
    f.depth[r] = depth[i];
    f.start[r] = start[i];
    f.count[r] = count[i];
    f.first_subnode[r] = nchild[i] > 0 ? first_child_ref[i] : -1;
    f.next_branch[r] = nb_bfs[i];
    f.path_hi[r] = phi[i];
    f.path_lo[r] = plo[i];
    f.nchild[r] = nchild[i];
    f.level_ids[i] = r;  // BFS order is level order
}

// ---- leaf-internal order: ascending original index (stable bucketing, tree.rs:813-828)
__global__ void mark_leaf_starts(const uint32_t* __restrict__ start, const uint8_t* __restrict__ nchild, int64_t nn,
                                 uint32_t* __restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    if (nchild[i] == 0) flag[start[i]] = 1;
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
// One thread per node of one level (deepest level first). Children are contiguous reference ids in
// octant order, so the sums run in the reference's order (tree.rs:876-929, 945-962, 1023-1063).
struct PayloadArgs {
    const int32_t* ids; int64_t count;
    const uint32_t* start; const uint32_t* pcount; const uint8_t* nchild; const int32_t* first_subnode;
    const double* spos; const double* smass; const double* sh;
    double* nmass; double* ncom; double* hmax; double* moments; int order; int ncoef;
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
}

__global__ void build_walk_records(const double* __restrict__ nmass, const double* __restrict__ ncom,
                                   const double* __restrict__ half, const double* __restrict__ hmax, double csep,
                                   const uint8_t* __restrict__ nchild, const uint32_t* __restrict__ start,
                                   const uint32_t* __restrict__ count, const int32_t* __restrict__ first_subnode,
                                   const int32_t* __restrict__ next_branch, int64_t nn, NodeRec* __restrict__ rec) {
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
    r.next_branch = next_branch[i];
    if (nmass[i] == 0.0) { r.kind = -2; r.first = -1; }          // tree.rs:1087-1090
    else if (nchild[i] == 0) { r.kind = (int32_t)count[i]; r.first = (int32_t)start[i]; }
    else { r.kind = -1; r.first = first_subnode[i]; }
    r.nleaf = 1;
    rec[i] = r;
}
// Leaf runs (tree.cuh) + the fp32 walk sources. One thread per sibling block (= per internal node, plus the root when
// the root itself is a leaf): merges every maximal run of consecutive non-zero-mass leaf children into the record of
// the run's first leaf and writes the run's particles relative to the run's origin (float64 subtraction, then the
// cast: close pairs keep their separation wherever the run sits in the box).
__global__ void merge_leaf_runs(const uint8_t* __restrict__ nchild, const int32_t* __restrict__ first_subnode,
                                const uint32_t* __restrict__ start, const uint32_t* __restrict__ count,
                                const int32_t* __restrict__ next_branch, const double* __restrict__ nmass,
                                const double* __restrict__ ncom, const double* __restrict__ spos,
                                const double* __restrict__ smass, int64_t nn, NodeRec* __restrict__ rec,
                                float4* __restrict__ src32) {
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
            NodeRec r = rec[j];
            r.com[0] = ox; r.com[1] = oy; r.com[2] = oz;
            r.kind = (int32_t)total;
            r.next_branch = next_branch[e - 1];
            r.nleaf = nl;
            rec[j] = r;
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
__global__ void update_gates(const double* __restrict__ hmax, double csep, int64_t nn, NodeRec* __restrict__ rec) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const double ch = __dmul_rn(csep, fmax(hmax[i], 0.0));
    rec[i].gate2 = __dmul_rn(ch, ch);
}
// fp32 walk records from the float64 moments (layout: multipole.cuh, m2p_fast)
__global__ void pack_walk_moments(const double* __restrict__ mom, int64_t nn, int order, int K, int rec,
                                  float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const double* m = mom + i * K;
    float* o = out + i * rec;
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

// Builds the topology from sorted keys. Returns false if `max_level` was reached with an over-full node.
bool build_topology(pnbx_tree_impl& t, cudaStream_t s, const DevBuf<double>& root4, const uint64_t* skhi,
                    const uint64_t* sklo, int max_level, StageTimer& tm) {
    const int64_t n = t.n;
    Bfs b;
    b.s = s;
    b.reserve(std::max<int64_t>(1024, n / 2 + 16));
    PNBX_LAUNCH(init_root, 1, 1, 0, s, b.start.p, b.count.p, b.parent.p, b.rank.p, b.depth.p, b.center.p, b.half.p,
                b.path_hi.p, b.path_lo.p, root4.get(), (uint32_t)n);
    b.size = 1;
    std::vector<int64_t> level_off{0, 1};
    int64_t level_begin = 0, level_size = 1;
    int depth = 0;
    const uint32_t cap = (uint32_t)std::min<int64_t>(t.leaf_capacity, UINT32_MAX);
    bool overflow = false;
    tm.begin("octree.emit_levels");
    while (level_size > 0) {
        if (depth == max_level) {
            // key width exhausted: is any node left that would have to split?
            std::vector<uint32_t> hc((size_t)level_size);
            PNBX_CUDA(cudaMemcpyAsync(hc.data(), b.count.p + level_begin, (size_t)level_size * 4, cudaMemcpyDeviceToHost, s));
            PNBX_CUDA(cudaStreamSynchronize(s));
            for (uint32_t c : hc) if (c > cap) { overflow = true; break; }
            PNBX_CUDA(cudaMemsetAsync(b.nchild.p + level_begin, 0, (size_t)level_size, s));
            PNBX_CUDA(cudaMemsetAsync(b.child_base.p + level_begin, 0xff, (size_t)level_size * 4, s));
            break;
        }
        const int child_level = depth + 1;
        DevBuf<uint32_t> bounds((size_t)level_size * 9, s);
        DevBuf<int32_t> nc((size_t)level_size + 1, s), off((size_t)level_size + 1, s);
        PNBX_CUDA(cudaMemsetAsync(nc.p, 0, ((size_t)level_size + 1) * 4, s));
        PNBX_LAUNCH(split_level, nblk(level_size * 9), 256, 0, s, b.start.p, b.count.p, level_begin, level_size,
                    child_level, cap, skhi, sklo, bounds.p, b.nchild.p, nc.p);
        PNBX_LAUNCH(count_children, nblk(level_size), 256, 0, s, b.count.p, level_begin, level_size, cap, bounds.p,
                    b.nchild.p, nc.p);
        exclusive_sum(nc.p, off.p, level_size + 1, s);
        int32_t total = 0;
        PNBX_CUDA(cudaMemcpyAsync(&total, off.p + level_size, 4, cudaMemcpyDeviceToHost, s));
        PNBX_CUDA(cudaStreamSynchronize(s));
        const int64_t next_begin = level_begin + level_size;
        if (total > 0) b.reserve(next_begin + total);
        PNBX_LAUNCH(emit_children, nblk(level_size), 256, 0, s, b.start.p, b.count.p, b.parent.p, b.child_base.p,
                    b.rank.p, b.depth.p, b.nchild.p, b.center.p, b.half.p, b.path_hi.p, b.path_lo.p, level_begin,
                    level_size, next_begin, off.p, bounds.p, child_level);
        if (total == 0) break;
        b.size = next_begin + total;
        level_begin = next_begin;
        level_size = total;
        level_off.push_back(b.size);
        ++depth;
    }
    tm.end();
    PNBX_CUDA(cudaGetLastError());
    if (overflow) return false;

    // ---- reference numbering
    tm.begin("octree.renumber_links");
    const int64_t nn = b.size;
    DevBuf<uint8_t> flag((size_t)nn, s);
    DevBuf<uint64_t> key((size_t)nn, s);
    PNBX_LAUNCH(internal_flags_keys, nblk(nn), 256, 0, s, b.nchild.p, b.start.p, b.depth.p, nn, flag.p, key.p);
    DevBuf<int32_t> int_idx((size_t)nn, s), n_sel(1, s);
    {
        size_t bytes = 0;
        cub::CountingInputIterator<int32_t> it(0);
        PNBX_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes, it, flag.p, int_idx.p, n_sel.p, (int)nn, s));
        DevBuf<uint8_t> tmp(bytes, s);
        PNBX_CUDA(cub::DeviceSelect::Flagged(tmp.get(), bytes, it, flag.p, int_idx.p, n_sel.p, (int)nn, s));
        ++launch_counter();
    }
    int32_t ni = 0;
    PNBX_CUDA(cudaMemcpyAsync(&ni, n_sel.p, 4, cudaMemcpyDeviceToHost, s));
    PNBX_CUDA(cudaStreamSynchronize(s));
    DevBuf<int32_t> first_child_ref((size_t)nn, s), ref((size_t)nn, s), nb_bfs((size_t)nn, s);
    PNBX_CUDA(cudaMemsetAsync(first_child_ref.p, 0xff, (size_t)nn * 4, s));
    if (ni > 0) {
        DevBuf<uint64_t> ikey((size_t)ni, s), ikey_s((size_t)ni, s);
        DevBuf<uint32_t> iidx_s((size_t)ni, s);
        PNBX_LAUNCH(gather_u32<uint64_t>, nblk(ni), 256, 0, s, key.p, (const uint32_t*)int_idx.p, (int64_t)ni, ikey.p);
        sort_pairs<uint64_t>(ikey.p, ikey_s.p, (const uint32_t*)int_idx.p, iidx_s.p, ni, 8 + bits_for((uint64_t)t.n), s);
        DevBuf<int32_t> ncs((size_t)ni, s), scan((size_t)ni, s);
        PNBX_LAUNCH(gather_nchild, nblk(ni), 256, 0, s, b.nchild.p, (const int32_t*)iidx_s.p, (int64_t)ni, ncs.p);
        exclusive_sum(ncs.p, scan.p, ni, s);
        PNBX_LAUNCH(scatter_first_child, nblk(ni), 256, 0, s, (const int32_t*)iidx_s.p, scan.p, (int64_t)ni,
                    first_child_ref.p);
    }
    PNBX_LAUNCH(assign_ref_ids, nblk(nn), 256, 0, s, b.parent.p, b.rank.p, first_child_ref.p, nn, ref.p);
    for (size_t d = 0; d + 1 < level_off.size(); ++d) {
        const int64_t lb = level_off[d], ls = level_off[d + 1] - lb;
        PNBX_LAUNCH(links_level, nblk(ls), 256, 0, s, b.parent.p, b.rank.p, b.nchild.p, ref.p, lb, ls, nb_bfs.p);
    }
    t.nn = nn;
    t.depth = (int)level_off.size() - 2;
    t.level_off = level_off;
    t.center.alloc((size_t)3 * nn, s); t.half.alloc((size_t)nn, s); t.node_depth.alloc((size_t)nn, s);
    t.node_start.alloc((size_t)nn, s); t.node_count.alloc((size_t)nn, s);
    t.first_subnode.alloc((size_t)nn, s); t.next_branch.alloc((size_t)nn, s);
    t.path_hi.alloc((size_t)nn, s); t.path_lo.alloc((size_t)nn, s);
    t.level_ids.alloc((size_t)nn, s);
    DevBuf<uint8_t> nchild_ref((size_t)nn, s);
    FinalArrays f{t.center.p, t.half.p, t.node_depth.p, t.node_start.p, t.node_count.p, t.first_subnode.p,
                  t.next_branch.p, t.path_hi.p, t.path_lo.p, nchild_ref.p, t.level_ids.p};
    PNBX_LAUNCH(scatter_nodes, nblk(nn), 256, 0, s, b.start.p, b.count.p, b.parent.p, b.depth.p, b.nchild.p, b.center.p,
                b.half.p, b.path_hi.p, b.path_lo.p, ref.p, first_child_ref.p, nb_bfs.p, nn, f);
    t.n_leaves = nn - ni;
    tm.end();

    // ---- ascending original index inside every leaf: stable sort of particles by leaf ordinal
    tm.begin("octree.leaf_order");
    if (n > 0) {
        DevBuf<uint32_t> lflag((size_t)n, s), lord_sorted((size_t)n, s), lord_orig((size_t)n, s), lord_out((size_t)n, s);
        DevBuf<uint32_t> iota((size_t)n, s), perm2((size_t)n, s);
        PNBX_CUDA(cudaMemsetAsync(lflag.p, 0, (size_t)n * 4, s));
        PNBX_LAUNCH(mark_leaf_starts, nblk(nn), 256, 0, s, b.start.p, b.nchild.p, nn, lflag.p);
        inclusive_sum_u32(lflag.p, lord_sorted.p, n, s);
        PNBX_LAUNCH(leaf_ordinal_to_particles, nblk(n), 256, 0, s, lord_sorted.p, t.perm.p, n, lord_orig.p);
        PNBX_LAUNCH(iota_u32, nblk(n), 256, 0, s, iota.p, n);
        sort_pairs<uint32_t>(lord_orig.p, lord_out.p, iota.p, perm2.p, n, bits_for((uint64_t)t.n_leaves + 1), s);
        t.perm = std::move(perm2);
    }
    tm.end();
    PNBX_CUDA(cudaGetLastError());
    t.node_nchild = std::move(nchild_ref);  // child counts in reference numbering, for the payload sweeps
    return true;
}

void gather_sorted_sources(pnbx_tree_impl& t, cudaStream_t s) {
    const int64_t n = t.n;
    if (n == 0) return;
    if (!t.spos.p) t.spos.alloc((size_t)3 * n, s);
    if (t.has_mass && !t.smass.p) t.smass.alloc((size_t)n, s);
    PNBX_LAUNCH(gather_sources, nblk(n), 256, 0, s, t.pos.p, t.has_mass ? t.mass.p : nullptr, t.perm.p, n, t.spos.p,
                t.has_mass ? t.smass.p : nullptr);
}
void gather_sorted_soft(pnbx_tree_impl& t, cudaStream_t s) {
    const int64_t n = t.n;
    if (!t.has_h || n == 0) return;
    if (!t.sh.p) { t.sh.alloc((size_t)n, s); t.sh32.alloc((size_t)n + 4, s); }
    PNBX_LAUNCH(gather_soft, nblk(n), 256, 0, s, t.h.p, t.perm.p, n, t.sh.p, t.sh32.p);
}

// tree.rs:968-1012
void build_mass_payload(pnbx_tree_impl& t, cudaStream_t s, StageTimer& tm) {
    const int64_t nn = t.nn;
    tm.begin("octree.build_mass_payload");
    gather_sorted_sources(t, s);  // masses may have been replaced
    t.nmass.alloc((size_t)nn, s);
    t.ncom.alloc((size_t)3 * nn, s);
    t.has_hmax = t.has_h;
    if (t.has_hmax) t.hmax.alloc((size_t)nn, s); else t.hmax.release();
    t.n_moments = mp::stored_coeffs(t.order);
    t.moments.alloc((size_t)nn * t.n_moments, s);
    PayloadArgs a;
    a.start = t.node_start.p; a.pcount = t.node_count.p; a.nchild = t.node_nchild.p; a.first_subnode = t.first_subnode.p;
    a.spos = t.spos.p; a.smass = t.has_mass ? t.smass.p : nullptr; a.sh = t.has_h ? t.sh.p : nullptr;
    a.nmass = t.nmass.p; a.ncom = t.ncom.p; a.hmax = t.has_hmax ? t.hmax.p : nullptr; a.moments = t.moments.p;
    a.order = t.order; a.ncoef = t.n_moments;
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
    a.ids = nullptr;
    a.count = nn;
    launch(true);  // every leaf of every level at once
    for (int d = (int)t.level_off.size() - 3; d >= 0; --d) {  // the deepest level holds leaves only
        a.ids = t.level_ids.p + t.level_off[d];
        a.count = t.level_off[d + 1] - t.level_off[d];
        launch(false);
    }
    t.rec.alloc((size_t)nn, s);
    PNBX_LAUNCH(build_walk_records, nblk(nn), 256, 0, s, t.nmass.p, t.ncom.p, t.half.p, t.has_hmax ? t.hmax.p : nullptr,
                t.kernel == PNBX_KERNEL_SPLINE ? 1.0 : 2.8, t.node_nchild.p, t.node_start.p, t.node_count.p,
                t.first_subnode.p, t.next_branch.p, nn, t.rec.p);
    if (t.n > 0 && !t.src32.p) t.src32.alloc((size_t)t.n, s);
    PNBX_LAUNCH(merge_leaf_runs, nblk(nn), 256, 0, s, t.node_nchild.p, t.first_subnode.p, t.node_start.p, t.node_count.p,
                t.next_branch.p, t.nmass.p, t.ncom.p, t.spos.p, t.has_mass ? t.smass.p : nullptr, nn, t.rec.p,
                t.n > 0 ? t.src32.p : nullptr);
    t.rec32 = mp::fast_rec_floats(t.order);
    t.moments32.alloc((size_t)nn * t.rec32, s);
    PNBX_LAUNCH(pack_walk_moments, nblk(nn), 256, 0, s, t.moments.p, nn, t.order, t.n_moments, t.rec32, t.moments32.p);
    PNBX_CUDA(cudaGetLastError());
    t.has_payload = true;
    tm.end();
}

Exec tree_exec(const pnbx_tree_impl& t) {
    pnbx_opts o{t.device, PNBX_MEM_HOST, 0, 0, nullptr};
    return make_exec(&o);
}

}  // namespace
}  // namespace pnbx

using namespace pnbx;

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
        Exec ex = make_exec(opts);
        cudaStream_t s = ex.stream;
        StageTimer tm(s);
        auto t = std::make_unique<pnbx_tree_impl>();
        t->device = ex.device;
        t->n = n;
        t->leaf_capacity = std::max<int64_t>(leaf_capacity, 1);  // tree.rs:701
        t->order_raw = multipole_order;
        t->order = std::min(multipole_order, 5);                 // tree.rs:1020
        t->kernel = kernel;
        t->has_mass = mass != nullptr;
        t->has_h = h != nullptr;

        tm.begin("octree.copy_in");
        auto copy_in = [&](DevBuf<double>& dst, const double* src, size_t cnt) {
            dst.alloc(std::max<size_t>(cnt, 1), s);
            if (!cnt) return;
            if (ex.device_ptrs) PNBX_CUDA(cudaMemcpyAsync(dst.p, src, cnt * sizeof(double), cudaMemcpyDeviceToDevice, s));
            else copy_h2d(dst.p, src, cnt * sizeof(double), ex);
        };
        copy_in(t->pos, pos, (size_t)3 * n);
        if (mass) copy_in(t->mass, mass, (size_t)n);
        if (h) copy_in(t->h, h, (size_t)n);
        tm.end();

        DevBuf<double> bb(6, s), root4(4, s);
        if (n > 0) {
            tm.begin("octree.bbox");
            launch_bbox(t->pos.p, n, bb.p, s);
            PNBX_LAUNCH(root_from_bbox, 1, 1, 0, s, bb.p, root4.p);
            double r4[4];
            PNBX_CUDA(cudaMemcpyAsync(r4, root4.p, sizeof(r4), cudaMemcpyDeviceToHost, s));
            PNBX_CUDA(cudaStreamSynchronize(s));
            for (int i = 0; i < 3; ++i) t->root_center[i] = r4[i];
            t->root_half = r4[3];
            tm.end();
        } else {
            // empty point set: the reference gets a NaN-centred, half = 1e-6 root leaf of mass 0
            double r4[4] = {0.0, 0.0, 0.0, 1e-6};
            PNBX_CUDA(cudaMemcpyAsync(root4.p, r4, sizeof(r4), cudaMemcpyHostToDevice, s));
            t->root_half = 1e-6;
        }

        // keys + sort; first with the 21-level word only, the second word only if a level-21 cell is over-full
        t->key_hi.alloc((size_t)std::max<int64_t>(n, 1), s);
        t->key_lo.alloc((size_t)std::max<int64_t>(n, 1), s);
        t->perm.alloc((size_t)std::max<int64_t>(n, 1), s);
        bool ok = false;
        for (int attempt = 0; attempt < 2 && !ok; ++attempt) {
            const bool use_lo = attempt == 1;
            const int levels = use_lo ? KEY_LEVELS : KEY_LEVELS_HI;
            DevBuf<uint64_t> skhi((size_t)std::max<int64_t>(n, 1), s), sklo;
            if (n > 0) {
                tm.begin("octree.keys_sort");
                PNBX_LAUNCH(path_keys, nblk(n), 256, 0, s, t->pos.p, n, root4.p, levels, t->key_hi.p, t->key_lo.p);
                DevBuf<uint32_t> iota((size_t)n, s);
                PNBX_LAUNCH(iota_u32, nblk(n), 256, 0, s, iota.p, n);
                if (!use_lo) {
                    sort_pairs<uint64_t>(t->key_hi.p, skhi.p, iota.p, t->perm.p, n, 63, s);
                } else {
                    DevBuf<uint64_t> tmpk((size_t)n, s), hi1((size_t)n, s);
                    DevBuf<uint32_t> idx1((size_t)n, s);
                    sort_pairs<uint64_t>(t->key_lo.p, tmpk.p, iota.p, idx1.p, n, 63, s);
                    PNBX_LAUNCH(gather_u32<uint64_t>, nblk(n), 256, 0, s, t->key_hi.p, idx1.p, n, hi1.p);
                    sort_pairs<uint64_t>(hi1.p, skhi.p, idx1.p, t->perm.p, n, 63, s);
                    sklo.alloc((size_t)n, s);
                    PNBX_LAUNCH(gather_u32<uint64_t>, nblk(n), 256, 0, s, t->key_lo.p, t->perm.p, n, sklo.p);
                }
                tm.end();
            }
            ok = build_topology(*t, s, root4, skhi.p, sklo.p, levels, tm);
        }
        if (!ok)
            throw ArgError{PNBX_ERR_DEPTH,
                           "octree deeper than 42 levels (more than leaf_capacity coincident or nearly coincident "
                           "points); the reference would recurse without bound here"};
        gather_sorted_soft(*t, s);
        if (t->has_mass) build_mass_payload(*t, s, tm);  // gravity.rs:210-220
        else gather_sorted_sources(*t, s);
        PNBX_CUDA(cudaStreamSynchronize(s));
        if (ex.own_stream) cudaStreamDestroy(ex.stream);
        *out = reinterpret_cast<pnbx_tree*>(t.release());
    });
}

extern "C" int pnbx_tree_build_mass(pnbx_tree* tp, const double* mass) {
    return guarded([&] {
        if (!tp) throw ArgError{PNBX_ERR_ARG, "tree is NULL"};
        auto& t = *reinterpret_cast<pnbx_tree_impl*>(tp);
        Exec ex = tree_exec(t);
        StageTimer tm(ex.stream);
        if (mass) {
            t.mass.alloc((size_t)std::max<int64_t>(t.n, 1), ex.stream);
            if (t.n) PNBX_CUDA(cudaMemcpyAsync(t.mass.p, mass, (size_t)t.n * sizeof(double), cudaMemcpyHostToDevice, ex.stream));
            t.has_mass = true;
        }
        build_mass_payload(t, ex.stream, tm);
        finish_exec(ex);
    });
}

extern "C" int pnbx_tree_set_softenings(pnbx_tree* tp, const double* h) {
    return guarded([&] {
        if (!tp) throw ArgError{PNBX_ERR_ARG, "tree is NULL"};
        auto& t = *reinterpret_cast<pnbx_tree_impl*>(tp);
        Exec ex = tree_exec(t);
        if (h) {
            t.h.alloc((size_t)std::max<int64_t>(t.n, 1), ex.stream);
            if (t.n) PNBX_CUDA(cudaMemcpyAsync(t.h.p, h, (size_t)t.n * sizeof(double), cudaMemcpyHostToDevice, ex.stream));
            t.has_h = true;
            gather_sorted_soft(t, ex.stream);
        } else {
            t.has_h = false;
        }
        // hmax is deliberately left as built (tree.rs:777-782 does not touch it)
        finish_exec(ex);
    });
}

extern "C" int pnbx_tree_set_kernel(pnbx_tree* tp, int kernel) {
    return guarded([&] {
        if (!tp) throw ArgError{PNBX_ERR_ARG, "tree is NULL"};
        if (kernel != PNBX_KERNEL_PLUMMER && kernel != PNBX_KERNEL_SPLINE)
            throw ArgError{PNBX_ERR_ARG, "kernel must be 0 (Plummer) or 1 (CubicSplineW2)"};
        auto& t = *reinterpret_cast<pnbx_tree_impl*>(tp);
        if (t.kernel != kernel && t.has_payload && t.has_hmax) {  // the gate factor c depends on the kernel (kernel.rs:20-28)
            Exec ex = tree_exec(t);
            PNBX_LAUNCH(update_gates, nblk(t.nn), 256, 0, ex.stream, t.hmax.p, kernel == PNBX_KERNEL_SPLINE ? 1.0 : 2.8, t.nn,
                        t.rec.p);
            PNBX_CUDA(cudaGetLastError());
            finish_exec(ex);
        }
        t.kernel = kernel;
    });
}

extern "C" void pnbx_tree_destroy(pnbx_tree* tp) {
    if (!tp) return;
    auto* t = reinterpret_cast<pnbx_tree_impl*>(tp);
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(t->device);
    delete t;
    cudaSetDevice(cur);
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
        Exec ex = tree_exec(t);
        cudaStream_t s = ex.stream;
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
        finish_exec(ex);
    });
}

extern "C" int pnbx_tree_dump_payload(const pnbx_tree* tp, double* mass, double* com, double* hmax, double* moments) {
    return guarded([&] {
        if (!tp) throw ArgError{PNBX_ERR_ARG, "tree is NULL"};
        const auto& t = *reinterpret_cast<const pnbx_tree_impl*>(tp);
        if (!t.has_payload) throw ArgError{PNBX_ERR_STATE, "mass payload not built; call build_mass() first"};
        Exec ex = tree_exec(t);
        const size_t nn = (size_t)t.nn;
        d2h(mass, t.nmass.p, nn, ex.stream);
        d2h(com, t.ncom.p, 3 * nn, ex.stream);
        if (t.has_hmax) d2h(hmax, t.hmax.p, nn, ex.stream);
        d2h(moments, t.moments.p, nn * t.n_moments, ex.stream);
        PNBX_CUDA(cudaStreamSynchronize(ex.stream));
        finish_exec(ex);
    });
}

extern "C" int pnbx_tree_dump_keys(const pnbx_tree* tp, uint64_t* key_hi, uint64_t* key_lo) {
    return guarded([&] {
        if (!tp) throw ArgError{PNBX_ERR_ARG, "tree is NULL"};
        const auto& t = *reinterpret_cast<const pnbx_tree_impl*>(tp);
        Exec ex = tree_exec(t);
        d2h(key_hi, t.key_hi.p, (size_t)t.n, ex.stream);
        d2h(key_lo, t.key_lo.p, (size_t)t.n, ex.stream);
        PNBX_CUDA(cudaStreamSynchronize(ex.stream));
        finish_exec(ex);
    });
}
