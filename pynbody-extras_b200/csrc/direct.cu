// direct.cu — K1: direct-summation potential / acceleration on B200 (sm_100a).
//
// Reproduces the eight solvers of the reference's direct.rs:115-658 (Newtonian and
// Plummer / cubic-spline softened, self and at-points) behind pnbx_direct().
//
// Design: the pairwise sum is FP32-FMA-pipe bound, not a contraction, so no
// tensor cores. Sources are packed once to float4 (x,y,z,m) relative to the bounding-box
// centre (float64 subtraction before the cast), streamed through shared memory by TMA bulk
// copies (cp.async.bulk + mbarrier, 3 stages) and broadcast to a register-blocked loop of
// TPT targets per thread. FP32 partial sums are folded into float64 accumulators once per
// 512-source tile, so no accumulator ever absorbs more than 512 terms (fp32 accuracy).
// The source range is split over blockIdx.y so the grid holds >= ~20 waves of work items; split
// partials are combined in a fixed order (deterministic, no float atomics).
#include <cfloat>
#include <mutex>

#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "spline.cuh"

namespace pnbx {
namespace {

#ifndef PNBX_DT
#define PNBX_DT 256
#endif
#ifndef PNBX_TPT
#define PNBX_TPT 4
#endif
#ifndef PNBX_MINB
#define PNBX_MINB 2
#endif
#ifndef PNBX_UNROLL
#define PNBX_UNROLL 4
#endif
constexpr int DT = PNBX_DT;    // threads per block
constexpr int TPT = PNBX_TPT;  // targets per thread
constexpr int UNROLL_F2 = PNBX_UNROLL;  // pair records per unrolled step of the packed loop
constexpr int STAGES = 3;   // smem pipeline depth
constexpr int TILE32 = 512; // sources per stage, fp32 path (8 KB of float4)
constexpr int TILE64 = 256; // sources per stage, fp64 verification path (8 KB of double4)

enum Soft { SOFT_NEWTON = 0, SOFT_PLUMMER_CONST = 1, SOFT_PLUMMER_PAIR = 2, SOFT_SPLINE = 3 };

// Which kernel variant a call needs depends on the DATA (are the softenings / masses constant? is one negative?).
// That is decided on the device (classify_direct) and every candidate variant is launched with a gate: the ones that
// were not chosen return at once. No host round trip, so device-pointer calls stay purely stream-ordered.
enum Variant { V_F2_CONSTM = 0, V_F2 = 1, V_F2H = 2, V_SCALAR_CONST = 3, V_SCALAR_PAIR = 4, V_F2H_CONSTM = 5 };
struct DirectPlan {
    int variant;
    float eps2_f, mass_f;
    double eps2_d, mass_d;
};

// The plan is read by the kernels from CONSTANT memory: a stream-ordered device-to-device copy puts it into one of
// PLAN_SLOTS slots of c_plans (slot index = kernel parameter), so eps^2 and the gate arrive through the constant bank in
// uniform registers exactly like kernel parameters would (as a value loaded from global memory eps^2 occupied a vector
// register and made every FFMA2 of the r^2 chain a three-vector-operand instruction: measured 2 % slower).
constexpr int PLAN_SLOTS = 64;
__constant__ DirectPlan c_plans[PLAN_SLOTS];

template <class T>
struct alignas(sizeof(T) * 4) Vec4 {
    T x, y, z, w;
};

// ------------------------------------------------------------------ mbarrier / TMA bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (UBLKCP in SASS).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float rsqrt_fast(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // one MUFU.RSQ
    return y;
}
__device__ __forceinline__ double rsqrt_fast(double x) { return 1.0 / sqrt(x); }
__device__ __forceinline__ float tmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double tmax(double a, double b) { return fmax(a, b); }
template <class T>
__device__ __forceinline__ T tiny();
template <>
__device__ __forceinline__ float tiny<float>() { return FLT_MIN; }
template <>
__device__ __forceinline__ double tiny<double>() { return DBL_MIN; }  // R2_TINY, direct.rs:7

// Springel W2 kernel and derivative (kernel.rs:84-128), evaluated only for r < h.
template <class T>
__device__ __forceinline__ T w2_inner(T u) {
    T u2 = u * u;
    if (u < T(0.5)) return T(16.0 / 3.0) * u2 + u2 * u2 * (T(32.0 / 5.0) * u - T(48.0 / 5.0)) - T(14.0 / 5.0);
    return T(1.0 / 15.0) / u + u2 * (T(32.0 / 3.0) + u * (T(-16.0) + u * (T(48.0 / 5.0) - T(32.0 / 15.0) * u))) -
           T(16.0 / 5.0);
}
template <class T>
__device__ __forceinline__ T w2p_inner(T u) {
    T u2 = u * u;
    if (u < T(0.5)) return u * (T(32.0 / 3.0) + u2 * (T(32.0) * u - T(192.0 / 5.0)));
    return T(-1.0 / 15.0) / u2 + u * (T(64.0 / 3.0) + u * (T(-48.0) + u * (T(192.0 / 5.0) - T(32.0 / 3.0) * u)));
}

// One target-source interaction. s = (dx-able source xyz, mass); hj = source softening.
// e = per-target constant: eps^2 (+tiny) for NEWTON / PLUMMER_CONST, h_i for the pair modes.
template <int WANT, int SOFT, bool CHECK, class T>
__device__ __forceinline__ void interact(T xi, T yi, T zi, T e, const Vec4<T>& s, T hj, bool is_self, T& ax, T& ay,
                                         T& az, T& pot) {
    T dx = s.x - xi, dy = s.y - yi, dz = s.z - zi;
    T m = s.w;
    if (SOFT == SOFT_NEWTON || SOFT == SOFT_PLUMMER_CONST || SOFT == SOFT_PLUMMER_PAIR) {
        T r2;
        if (SOFT == SOFT_PLUMMER_PAIR) {
            T h = tmax(e, hj);  // direct.rs:402,426,475,506 (self) / :560,577,623,646 (points: e = 0)
            r2 = fma(h, h, tiny<T>());
        } else {
            r2 = e;
        }
        r2 = fma(dx, dx, r2);
        r2 = fma(dy, dy, r2);
        r2 = fma(dz, dz, r2);
        if (CHECK && is_self) {  // own particle (index match): contributes exactly nothing
            r2 = T(1);
            m = T(0);
        }
        T rinv = rsqrt_fast(r2);
        if (WANT == PNBX_WANT_POT) pot = fma(-m, rinv, pot);
        if (WANT & PNBX_WANT_ACC) {
            T mr = m * rinv;
            if (WANT & PNBX_WANT_POT) pot -= mr;
            T g = mr * (rinv * rinv);
            ax = fma(dx, g, ax);
            ay = fma(dy, g, ay);
            az = fma(dz, g, az);
        }
    } else {  // SOFT_SPLINE
        T h = tmax(e, hj);
        T r2 = fma(dx, dx, tiny<T>());
        r2 = fma(dy, dy, r2);
        r2 = fma(dz, dz, r2);
        if (CHECK && is_self) {
            r2 = T(1);
            m = T(0);
        }
        T rinv = rsqrt_fast(r2);
        T k = -rinv;                   // potential per unit mass
        T g = rinv * rinv * rinv;      // accel factor
        if (h > T(0) && r2 < h * h) {  // kernel.rs:46-54, 72-80
            T hinv = T(1) / h;
            T u = (r2 * rinv) * hinv;
            if (WANT & PNBX_WANT_POT) k = w2_inner(u) * hinv;
            if (WANT & PNBX_WANT_ACC) g = w2p_inner(u) * (hinv * hinv) * rinv;
        }
        if (WANT & PNBX_WANT_POT) pot = fma(m, k, pot);
        if (WANT & PNBX_WANT_ACC) {
            T mg = m * g;
            ax = fma(dx, mg, ax);
            ay = fma(dy, mg, ay);
            az = fma(dz, mg, az);
        }
    }
}

// Cubic-spline fast path, pass 1: the pair as Newtonian unless it lies inside the softening (r < h), in which case
// it contributes nothing here and `inside` is raised; pass 2 (spline_inside) adds the W2 terms of those rare pairs.
template <int WANT, class T>
__device__ __forceinline__ void spline_outside(T xi, T yi, T zi, T hi, const Vec4<T>& s, T hj, bool& inside, T& ax,
                                               T& ay, T& az, T& pot) {
    const T dx = s.x - xi, dy = s.y - yi, dz = s.z - zi;
    const T h = tmax(hi, hj);  // both clamped at 0 when packed, so "h > 0" is implied by r2 < h*h (r2 >= tiny > 0)
    T r2 = fma(dx, dx, tiny<T>());
    r2 = fma(dy, dy, r2);
    r2 = fma(dz, dz, r2);
    const bool in = fma(-h, h, r2) < T(0);  // r < h: kernel.rs:46-54, 72-80
    inside |= in;
    const T m = in ? T(0) : s.w;
    const T rinv = rsqrt_fast(r2);
    const T mr = m * rinv;
    if (WANT & PNBX_WANT_POT) pot -= mr;
    if (WANT & PNBX_WANT_ACC) {
        const T g = mr * (rinv * rinv);
        ax = fma(dx, g, ax);
        ay = fma(dy, g, ay);
        az = fma(dz, g, az);
    }
}
template <int WANT, class T>
__device__ __forceinline__ void spline_inside(T xi, T yi, T zi, T hi, const Vec4<T>& s, T hj, T& ax, T& ay, T& az,
                                              T& pot) {
    const T dx = s.x - xi, dy = s.y - yi, dz = s.z - zi;
    const T h = tmax(hi, hj);
    T r2 = fma(dx, dx, tiny<T>());
    r2 = fma(dy, dy, r2);
    r2 = fma(dz, dz, r2);
    if (!(fma(-h, h, r2) < T(0))) return;  // the same predicate as pass 1, bit for bit
    const T rinv = rsqrt_fast(r2);
    const T hinv = T(1) / h;
    const T u = (r2 * rinv) * hinv;
    if (WANT & PNBX_WANT_POT) pot = fma(s.w, w2_inner(u) * hinv, pot);
    if (WANT & PNBX_WANT_ACC) {
        const T mg = s.w * (w2p_inner(u) * (hinv * hinv) * rinv);
        ax = fma(dx, mg, ax);
        ay = fma(dy, mg, ay);
        az = fma(dz, mg, az);
    }
}

template <int WANT, int SOFT, class T, int TILE>
__global__ void __launch_bounds__(DT, PNBX_MINB)
direct_kernel(const Vec4<T>* __restrict__ src, const T* __restrict__ src_h, int64_t n_src,
              const Vec4<T>* __restrict__ tgt, const T* __restrict__ tgt_h, int64_t m, int64_t self_base,
              int plan_slot, int my_variant, int tiles_per_split, double* __restrict__ out_pot,
              double* __restrict__ out_acc) {
    constexpr bool PAIR_H = (SOFT == SOFT_PLUMMER_PAIR || SOFT == SOFT_SPLINE);
    const DirectPlan& plan = c_plans[plan_slot];
    if (plan.variant != my_variant) return;  // not the variant this call's data needs (whole grid, before any barrier)
    const T eps2_const = sizeof(T) == 4 ? (T)plan.eps2_f : (T)plan.eps2_d;
    __shared__ Vec4<T> s_src[STAGES][TILE];
    __shared__ alignas(16) T s_h[PAIR_H ? STAGES : 1][PAIR_H ? TILE : 4];
    __shared__ alignas(8) uint64_t s_full[STAGES];

    const int tid = threadIdx.x;
    const int64_t n_tiles = (n_src + TILE - 1) / TILE;
    const int64_t tile_begin = (int64_t)blockIdx.y * tiles_per_split;
    const int64_t tile_end = tile_begin + tiles_per_split < n_tiles ? tile_begin + tiles_per_split : n_tiles;
    const int64_t tgt_base = (int64_t)blockIdx.x * (DT * TPT);

    // ---- targets in registers
    T xi[TPT], yi[TPT], zi[TPT], ei[TPT];
    int64_t gi[TPT];
#pragma unroll
    for (int k = 0; k < TPT; ++k) {
        int64_t i = tgt_base + k * DT + tid;
        if (i > m - 1) i = m - 1;
        Vec4<T> t = tgt[i];
        xi[k] = t.x; yi[k] = t.y; zi[k] = t.z;
        if (PAIR_H) ei[k] = tgt_h ? tgt_h[i] : T(0);
        else ei[k] = eps2_const;
        gi[k] = self_base >= 0 ? self_base + i : -1;
    }
    double Ax[TPT], Ay[TPT], Az[TPT], P[TPT];
#pragma unroll
    for (int k = 0; k < TPT; ++k) Ax[k] = Ay[k] = Az[k] = P[k] = 0.0;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&s_full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int64_t tile) {  // called by thread 0 only
        int st = (int)((tile - tile_begin) % STAGES);
        int64_t j0 = tile * TILE;
        int cnt = (int)(n_src - j0 < TILE ? n_src - j0 : TILE);
        uint32_t bytes = (uint32_t)cnt * sizeof(Vec4<T>);
        uint32_t hbytes = PAIR_H ? (uint32_t)((cnt * sizeof(T) + 15) / 16 * 16) : 0u;  // src_h is padded
        mbar_expect_tx(&s_full[st], bytes + hbytes);
        bulk_g2s(&s_src[st][0], src + j0, bytes, &s_full[st]);
        if (PAIR_H) bulk_g2s(&s_h[st][0], src_h + j0, hbytes, &s_full[st]);
    };
    if (tid == 0)
        for (int s = 0; s < STAGES - 1; ++s)
            if (tile_begin + s < tile_end) issue(tile_begin + s);

    const int64_t blk_lo = self_base >= 0 ? self_base + tgt_base : INT64_MAX;
    const int64_t blk_hi = self_base >= 0 ? blk_lo + DT * TPT : INT64_MIN;

    for (int64_t tile = tile_begin; tile < tile_end; ++tile) {
        const int it = (int)(tile - tile_begin);
        const int st = it % STAGES;
        if (tid == 0 && tile + STAGES - 1 < tile_end) issue(tile + STAGES - 1);
        mbar_wait(&s_full[st], (uint32_t)((it / STAGES) & 1));

        const int64_t j0 = tile * TILE;
        const int cnt = (int)(n_src - j0 < TILE ? n_src - j0 : TILE);
        T ax[TPT], ay[TPT], az[TPT], p[TPT];
#pragma unroll
        for (int k = 0; k < TPT; ++k) ax[k] = ay[k] = az[k] = p[k] = T(0);

        const bool diag = (j0 < blk_hi) && (j0 + cnt > blk_lo);  // tile holds some of this block's own particles
        if (!diag && cnt == TILE && SOFT == SOFT_SPLINE) {
            // two passes: branch-free Newtonian sweep, then (only if some lane of the warp saw a pair with r < h,
            // which is rare) a second sweep of the tile that adds the W2-kernel terms of exactly those pairs
            bool inside = false;
#pragma unroll 8
            for (int j = 0; j < TILE; ++j) {
                const Vec4<T> s = s_src[st][j];
                const T hj = s_h[st][j];
#pragma unroll
                for (int k = 0; k < TPT; ++k) spline_outside<WANT, T>(xi[k], yi[k], zi[k], ei[k], s, hj, inside, ax[k], ay[k], az[k], p[k]);
            }
            if (__any_sync(0xffffffffu, inside)) {
                if (inside) {
                    for (int j = 0; j < TILE; ++j) {
                        const Vec4<T> s = s_src[st][j];
                        const T hj = s_h[st][j];
#pragma unroll
                        for (int k = 0; k < TPT; ++k) spline_inside<WANT, T>(xi[k], yi[k], zi[k], ei[k], s, hj, ax[k], ay[k], az[k], p[k]);
                    }
                }
            }
        } else if (!diag && cnt == TILE) {
#pragma unroll 8
            for (int j = 0; j < TILE; ++j) {
                Vec4<T> s = s_src[st][j];
                T hj = PAIR_H ? s_h[st][j] : T(0);
#pragma unroll
                for (int k = 0; k < TPT; ++k) interact<WANT, SOFT, false, T>(xi[k], yi[k], zi[k], ei[k], s, hj, false, ax[k], ay[k], az[k], p[k]);
            }
        } else {
            // ragged tail and/or self-skip by index (direct.rs:166, 297, 421, 501)
            for (int j = 0; j < cnt; ++j) {
                Vec4<T> s = s_src[st][j];
                T hj = PAIR_H ? s_h[st][j] : T(0);
                const int64_t gj = j0 + j;
#pragma unroll
                for (int k = 0; k < TPT; ++k)
                    interact<WANT, SOFT, true, T>(xi[k], yi[k], zi[k], ei[k], s, hj, gj == gi[k], ax[k], ay[k], az[k], p[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < TPT; ++k) {  // fold the tile partials into float64
            if (WANT & PNBX_WANT_ACC) { Ax[k] += (double)ax[k]; Ay[k] += (double)ay[k]; Az[k] += (double)az[k]; }
            if (WANT & PNBX_WANT_POT) P[k] += (double)p[k];
        }
        __syncthreads();  // everyone is done with stage `st` before it is refilled
    }

    const int64_t split_off = (int64_t)blockIdx.y * m;
#pragma unroll
    for (int k = 0; k < TPT; ++k) {
        int64_t i = tgt_base + k * DT + tid;
        if (i < m) {
            if (WANT & PNBX_WANT_POT) out_pot[split_off + i] = P[k];
            if (WANT & PNBX_WANT_ACC) {
                double* a = out_acc + 3 * (split_off + i);
                a[0] = Ax[k]; a[1] = Ay[k]; a[2] = Az[k];
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Packed-FP32x2 variant (Blackwell FFMA2 / FADD2 / FMUL2, `*.f32x2` PTX) for the constant-softening
// modes. Two consecutive SOURCES share one instruction: the tile is stored as pair records
// {x0,x1,y0,y1,z0,z1,m0,m1} (32 B), the target coordinates are duplicated once into both halves.
// Per pair of interactions: 3 sub2 + 3 fma2 + 2 MUFU.RSQ + 3 mul2 + 3 fma2 (+1 fma2 for phi) =
// 14 issue slots for 24 FP32-pipe lane-cycles, so the kernel is bound by the FMA pipe itself
// (12 lane-ops per interaction) instead of by instruction issue (13 slots) — DESIGN.md §K1.
struct alignas(32) Pair8 {
    float x0, x1, y0, y1, z0, z1, m0, m1;
};
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_pack(float a, float b) {
    f2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void f2_unpack(f2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t f2_sub(f2_t a, f2_t b) {
    f2_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) {
    f2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b) {
    f2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) {
    f2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

constexpr int TILEP = TILE32 / 2;  // pair records per stage

// CONSTM: all source masses are equal (detected on the device; the usual case for a single particle family) —
// the mass is factored out of the loop (one multiply less per interaction: 11 lane-ops) and applied once at the end.
template <int WANT, bool CONSTM>
__global__ void __launch_bounds__(DT, PNBX_MINB)
direct_kernel_f2(const Pair8* __restrict__ src, int64_t n_src, const Vec4<float>* __restrict__ tgt, int64_t m,
                 int64_t self_base, int plan_slot, int tiles_per_split,
                 double* __restrict__ out_pot, double* __restrict__ out_acc) {
    __shared__ Pair8 s_src[STAGES][TILEP];
    const DirectPlan& plan = c_plans[plan_slot];
    if (plan.variant != (CONSTM ? V_F2_CONSTM : V_F2)) return;
    const float eps2_const = plan.eps2_f;
    __shared__ alignas(8) uint64_t s_full[STAGES];
    const int tid = threadIdx.x;
    const int64_t n_pairs = (n_src + 1) / 2;  // the odd tail is padded by pack_pairs (zero mass, far away)
    const int64_t n_tiles = (n_pairs + TILEP - 1) / TILEP;
    const int64_t tile_begin = (int64_t)blockIdx.y * tiles_per_split;
    const int64_t tile_end = tile_begin + tiles_per_split < n_tiles ? tile_begin + tiles_per_split : n_tiles;
    const int64_t tgt_base = (int64_t)blockIdx.x * (DT * TPT);

    f2_t xi[TPT], yi[TPT], zi[TPT];
    float xs[TPT], ys[TPT], zs[TPT];
    int64_t gi[TPT];
#pragma unroll
    for (int k = 0; k < TPT; ++k) {
        int64_t i = tgt_base + k * DT + tid;
        if (i > m - 1) i = m - 1;
        Vec4<float> t = tgt[i];
        xs[k] = t.x; ys[k] = t.y; zs[k] = t.z;
        xi[k] = f2_pack(t.x, t.x); yi[k] = f2_pack(t.y, t.y); zi[k] = f2_pack(t.z, t.z);
        gi[k] = self_base >= 0 ? self_base + i : -1;
    }
    const f2_t e2 = f2_pack(eps2_const, eps2_const);
    const f2_t one2 = f2_pack(1.f, 1.f);
    double Ax[TPT], Ay[TPT], Az[TPT], P[TPT];
#pragma unroll
    for (int k = 0; k < TPT; ++k) Ax[k] = Ay[k] = Az[k] = P[k] = 0.0;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&s_full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int64_t tile) {
        int st = (int)((tile - tile_begin) % STAGES);
        int64_t q0 = tile * TILEP;
        int cnt = (int)(n_pairs - q0 < TILEP ? n_pairs - q0 : TILEP);
        uint32_t bytes = (uint32_t)cnt * sizeof(Pair8);
        mbar_expect_tx(&s_full[st], bytes);
        bulk_g2s(&s_src[st][0], src + q0, bytes, &s_full[st]);
    };
    if (tid == 0)
        for (int s = 0; s < STAGES - 1; ++s)
            if (tile_begin + s < tile_end) issue(tile_begin + s);

    const int64_t blk_lo = self_base >= 0 ? self_base + tgt_base : INT64_MAX;
    const int64_t blk_hi = self_base >= 0 ? blk_lo + DT * TPT : INT64_MIN;

    for (int64_t tile = tile_begin; tile < tile_end; ++tile) {
        const int it = (int)(tile - tile_begin);
        const int st = it % STAGES;
        if (tid == 0 && tile + STAGES - 1 < tile_end) issue(tile + STAGES - 1);
        mbar_wait(&s_full[st], (uint32_t)((it / STAGES) & 1));
        const int64_t q0 = tile * TILEP;
        const int cnt = (int)(n_pairs - q0 < TILEP ? n_pairs - q0 : TILEP);
        const int64_t j0 = 2 * q0;
        const bool diag = (j0 < blk_hi) && (j0 + 2 * cnt > blk_lo);
        const bool has_pad = CONSTM && (n_src & 1) && tile == n_tiles - 1;  // the zero-mass pad needs its mass
        if (!diag && cnt == TILEP && !has_pad) {
            f2_t ax[TPT], ay[TPT], az[TPT], p[TPT];
#pragma unroll
            for (int k = 0; k < TPT; ++k) ax[k] = ay[k] = az[k] = p[k] = 0ull;  // {+0.f, +0.f}
#pragma unroll UNROLL_F2
            for (int q = 0; q < TILEP; ++q) {
                const ulonglong4 v = *reinterpret_cast<const ulonglong4*>(&s_src[st][q]);  // {x0x1, y0y1, z0z1, m0m1}
#pragma unroll
                for (int k = 0; k < TPT; ++k) {
                    const f2_t dx = f2_sub(v.x, xi[k]), dy = f2_sub(v.y, yi[k]), dz = f2_sub(v.z, zi[k]);
                    f2_t r2 = f2_fma(dx, dx, e2);
                    r2 = f2_fma(dy, dy, r2);
                    r2 = f2_fma(dz, dz, r2);
                    float ra, rb;
                    f2_unpack(r2, ra, rb);
                    const f2_t rinv = f2_pack(rsqrt_fast(ra), rsqrt_fast(rb));
                    if (WANT & PNBX_WANT_POT) p[k] = f2_fma(CONSTM ? one2 : v.w, rinv, p[k]);
                    if (WANT & PNBX_WANT_ACC) {
                        const f2_t rr = f2_mul(rinv, rinv);
                        const f2_t g = CONSTM ? f2_mul(rr, rinv) : f2_mul(f2_mul(v.w, rinv), rr);
                        ax[k] = f2_fma(dx, g, ax[k]);
                        ay[k] = f2_fma(dy, g, ay[k]);
                        az[k] = f2_fma(dz, g, az[k]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < TPT; ++k) {
                float a, b;
                if (WANT & PNBX_WANT_ACC) {
                    f2_unpack(ax[k], a, b); Ax[k] += (double)a + (double)b;
                    f2_unpack(ay[k], a, b); Ay[k] += (double)a + (double)b;
                    f2_unpack(az[k], a, b); Az[k] += (double)a + (double)b;
                }
                if (WANT & PNBX_WANT_POT) { f2_unpack(p[k], a, b); P[k] -= (double)a + (double)b; }
            }
        } else {
            // ragged tail and/or self-skip by index: scalar loop over the same pair records
            float ax[TPT], ay[TPT], az[TPT], p[TPT];
#pragma unroll
            for (int k = 0; k < TPT; ++k) ax[k] = ay[k] = az[k] = p[k] = 0.f;
            const int nsrc_tile = (int)((n_src - j0) < 2 * cnt ? (n_src - j0) : 2 * cnt);
            for (int j = 0; j < nsrc_tile; ++j) {
                const Pair8& pr = s_src[st][j >> 1];
                Vec4<float> sj;
                if (j & 1) { sj.x = pr.x1; sj.y = pr.y1; sj.z = pr.z1; sj.w = CONSTM ? 1.f : pr.m1; }
                else { sj.x = pr.x0; sj.y = pr.y0; sj.z = pr.z0; sj.w = CONSTM ? 1.f : pr.m0; }
                const int64_t gj = j0 + j;
#pragma unroll
                for (int k = 0; k < TPT; ++k)
                    interact<WANT, SOFT_PLUMMER_CONST, true, float>(xs[k], ys[k], zs[k], eps2_const, sj, 0.f, gj == gi[k],
                                                                    ax[k], ay[k], az[k], p[k]);
            }
#pragma unroll
            for (int k = 0; k < TPT; ++k) {
                if (WANT & PNBX_WANT_ACC) { Ax[k] += (double)ax[k]; Ay[k] += (double)ay[k]; Az[k] += (double)az[k]; }
                if (WANT & PNBX_WANT_POT) P[k] += (double)p[k];
            }
        }
        __syncthreads();
    }
    const int64_t split_off = (int64_t)blockIdx.y * m;
    const double ms = CONSTM ? plan.mass_d : 1.0;
#pragma unroll
    for (int k = 0; k < TPT; ++k) {
        int64_t i = tgt_base + k * DT + tid;
        if (i < m) {
            if (WANT & PNBX_WANT_POT) out_pot[split_off + i] = P[k] * ms;
            if (WANT & PNBX_WANT_ACC) {
                double* a = out_acc + 3 * (split_off + i);
                a[0] = Ax[k] * ms; a[1] = Ay[k] * ms; a[2] = Az[k] * ms;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Packed-FP32x2 variant for PER-PAIR softening h = max(h_i, h_j) (direct.rs:370-658), valid when every softening that
// takes part is >= 0 so that h^2 = max(h_i^2, h_j^2): the sources carry their squared softening next to the pair
// records (a second bulk copy per tile), the targets theirs in a register (0 for at-points targets: max(h_j, 0)).
//   HMODE 1, Plummer:  r2 = |d|^2 + max(h_i^2, h_j^2): 2 FMNMX more than the constant-softening loop per pair record.
//   HMODE 2, spline:   pass 1 = Newtonian sweep that zeroes the mass of pairs with r < h and flags them, pass 2 (only
//                      when some lane of the warp flagged one) adds their W2 terms, as in the scalar kernel.
// One scalar pair (diagonal / ragged tiles, and the spline's pass 2 with ONLY_INSIDE): same r2 chain as the packed loop.
template <int WANT, int HMODE, bool ONLY_INSIDE>
__device__ __forceinline__ void pair_h2_scalar(float xi, float yi, float zi, float th2, float sx, float sy, float sz,
                                               float sm, float hs2, bool is_self, float& ax, float& ay, float& az,
                                               float& pot) {
    const float dx = sx - xi, dy = sy - yi, dz = sz - zi;
    const float h2 = fmaxf(hs2, th2);
    if (HMODE == 1) {
        float r2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, fmaxf(h2, FLT_MIN))));
        float m = sm;
        if (is_self) { r2 = 1.f; m = 0.f; }
        const float rinv = rsqrt_fast(r2);
        const float mr = m * rinv;
        if (WANT & PNBX_WANT_POT) pot -= mr;
        if (WANT & PNBX_WANT_ACC) {
            const float g = mr * (rinv * rinv);
            ax = fmaf(dx, g, ax); ay = fmaf(dy, g, ay); az = fmaf(dz, g, az);
        }
    } else {
        const float r2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, FLT_MIN)));
        const bool in = r2 < h2;  // the packed pass-1 predicate, bit for bit
        if (ONLY_INSIDE && !in) return;
        if (is_self) return;
        const float rinv = rsqrt_fast(r2);
        float kpot = -rinv, kacc = rinv * rinv * rinv;
        if (in) {
            const float hinv = rsqrt_fast(h2);
            w2_terms_f32(r2, rinv, h2 * hinv, hinv, kpot, kacc);
        }
        if (WANT & PNBX_WANT_POT) pot = fmaf(sm, kpot, pot);
        if (WANT & PNBX_WANT_ACC) {
            const float g = sm * kacc;
            ax = fmaf(dx, g, ax); ay = fmaf(dy, g, ay); az = fmaf(dz, g, az);
        }
    }
}

// Extents of one source tile (TILEP pair records): squared-softening range and bounding box. A target block compares
// them with its own extents once per tile and picks a cheaper inner loop when the whole tile allows it:
//   Plummer, tile h_max <= block h_min: max(h_i, h_j) = h_i for every pair -> the softening is a per-target constant
//   Plummer, tile h_min >= block h_max: max(h_i, h_j) = h_j -> the addend comes straight from the source record
//   spline, boxes further apart than the largest softening radius of either side: every pair is Newtonian
// Each pair gets bit for bit the value the general loop would give it. Whole-array self calls sort the particles so
// that nearly all tiles qualify (run_direct: by softening for Plummer, along a Morton curve for the spline).
struct alignas(16) TileMeta {
    float h2min, h2max, lox, loy, loz, hix, hiy, hiz;
};
enum Regime { R_GENERAL = 0, R_TARGET_H = 1, R_SOURCE_H = 2, R_FAR = 3 };

// The packed pass over one full tile. REGIME is a compile-time copy of the loop for each case above.
template <int WANT, int HMODE, bool CONSTM, int REGIME>
__device__ __forceinline__ void f2h_tile(const Pair8* __restrict__ t_src, const float2* __restrict__ t_h2,
                                         const f2_t (&xi)[TPT], const f2_t (&yi)[TPT], const f2_t (&zi)[TPT],
                                         const float (&th2f)[TPT], f2_t (&ax)[TPT], f2_t (&ay)[TPT], f2_t (&az)[TPT],
                                         f2_t (&p)[TPT], unsigned (&inside)[TPT]) {
    const f2_t tiny2 = f2_pack(FLT_MIN, FLT_MIN);
#pragma unroll UNROLL_F2
    for (int q = 0; q < TILEP; ++q) {
        const ulonglong4 v = *reinterpret_cast<const ulonglong4*>(&t_src[q]);  // {x0x1, y0y1, z0z1, m0m1}
        float2 hh = make_float2(0.f, 0.f);
        if (REGIME == R_GENERAL || REGIME == R_SOURCE_H) hh = t_h2[q];
        const unsigned gbit = 1u << (q >> 3);
#pragma unroll
        for (int k = 0; k < TPT; ++k) {
            const f2_t dx = f2_sub(v.x, xi[k]), dy = f2_sub(v.y, yi[k]), dz = f2_sub(v.z, zi[k]);
            // th2f[k] = max(th2[k], FLT_MIN) for Plummer (the + R2_TINY folded into the softening), th2[k] for
            // the spline: one max per source, h^2 = max(h_s^2, h_t^2)
            float h2a = 0.f, h2b = 0.f;
            f2_t add = tiny2;
            if (REGIME == R_GENERAL) {
                h2a = fmaxf(hh.x, th2f[k]); h2b = fmaxf(hh.y, th2f[k]);
                if (HMODE == 1) add = f2_pack(h2a, h2b);
            } else if (REGIME == R_TARGET_H) {
                add = f2_pack(th2f[k], th2f[k]);
            } else if (REGIME == R_SOURCE_H) {
                add = f2_pack(hh.x, hh.y);
            }
            f2_t r2 = f2_fma(dx, dx, add);
            r2 = f2_fma(dy, dy, r2);
            r2 = f2_fma(dz, dz, r2);
            float ra, rb;
            f2_unpack(r2, ra, rb);
            float ia_r = rsqrt_fast(ra), ib_r = rsqrt_fast(rb);
            f2_t mm = v.w;
            if (HMODE == 2 && REGIME == R_GENERAL) {
                const bool ia = ra < h2a, ib = rb < h2b;  // r < h: left to pass 2 (no add-then-subtract)
                inside[k] |= (ia | ib) ? gbit : 0u;
                if (CONSTM) {  // zero 1/r: the pair then adds nothing to either sum
                    ia_r = ia ? 0.f : ia_r;
                    ib_r = ib ? 0.f : ib_r;
                } else {
                    float m0, m1;
                    f2_unpack(v.w, m0, m1);
                    mm = f2_pack(ia ? 0.f : m0, ib ? 0.f : m1);
                }
            }
            const f2_t rinv = f2_pack(ia_r, ib_r);
            if (WANT & PNBX_WANT_POT) p[k] = CONSTM ? f2_add(p[k], rinv) : f2_fma(mm, rinv, p[k]);
            if (WANT & PNBX_WANT_ACC) {
                const f2_t rr = f2_mul(rinv, rinv);
                const f2_t g = CONSTM ? f2_mul(rr, rinv) : f2_mul(f2_mul(mm, rinv), rr);
                ax[k] = f2_fma(dx, g, ax[k]);
                ay[k] = f2_fma(dy, g, ay[k]);
                az[k] = f2_fma(dz, g, az[k]);
            }
        }
    }
}

template <int WANT, int HMODE, bool CONSTM>
__global__ void __launch_bounds__(DT, PNBX_MINB)
direct_kernel_f2h(const Pair8* __restrict__ src, const float* __restrict__ src_h2, const TileMeta* __restrict__ tile_meta,
                  int64_t n_src, const Vec4<float>* __restrict__ tgt, const float* __restrict__ tgt_h2, int64_t m,
                  int64_t self_base, int plan_slot, int tiles_per_split, double* __restrict__ out_pot,
                  double* __restrict__ out_acc) {
    __shared__ Pair8 s_src[STAGES][TILEP];
    // CONSTM: all source masses equal (decided on the device like in direct_kernel_f2): the mass leaves the loops and is
    // applied once at the end; the spline's "pair inside its softening radius" zeroes 1/r instead of the mass
    if (c_plans[plan_slot].variant != (CONSTM ? V_F2H_CONSTM : V_F2H)) return;
    __shared__ alignas(16) float2 s_h2[STAGES][TILEP];
    __shared__ alignas(8) uint64_t s_full[STAGES];
    const int tid = threadIdx.x;
    const int64_t n_pairs = (n_src + 1) / 2;  // the odd tail is padded (zero mass, far away, h = 0)
    const int64_t n_tiles = (n_pairs + TILEP - 1) / TILEP;
    const int64_t tile_begin = (int64_t)blockIdx.y * tiles_per_split;
    const int64_t tile_end = tile_begin + tiles_per_split < n_tiles ? tile_begin + tiles_per_split : n_tiles;
    const int64_t tgt_base = (int64_t)blockIdx.x * (DT * TPT);

    f2_t xi[TPT], yi[TPT], zi[TPT];
    // squared target softening; Plummer folds the + R2_TINY in: max(h_s^2, h_t^2, FLT_MIN) = max(h_s^2, th2f)
    float th2f[TPT];
#pragma unroll
    for (int k = 0; k < TPT; ++k) {
        int64_t i = tgt_base + k * DT + tid;
        if (i > m - 1) i = m - 1;
        Vec4<float> t = tgt[i];
        xi[k] = f2_pack(t.x, t.x); yi[k] = f2_pack(t.y, t.y); zi[k] = f2_pack(t.z, t.z);
        const float t2 = tgt_h2 ? tgt_h2[i] : 0.f;
        th2f[k] = HMODE == 1 ? fmaxf(t2, FLT_MIN) : t2;
    }
    // global source index of this thread's k-th target: g0 + k*DT (lanes past the end are clamped duplicates whose
    // results are never stored, so their index does not matter)
    const int64_t g0 = self_base >= 0 ? self_base + tgt_base + tid : INT64_MIN / 2;
    auto lo = [](f2_t v) { float a, b; f2_unpack(v, a, b); return a; };
    double Ax[TPT], Ay[TPT], Az[TPT], P[TPT];
#pragma unroll
    for (int k = 0; k < TPT; ++k) Ax[k] = Ay[k] = Az[k] = P[k] = 0.0;

    // extents of this block's targets: {h2 min, -h2 max, x min, -x max, y min, -y max, z min, -z max}, all as minima
    __shared__ float s_ext[8][DT / 32];
    {
        float ext[8];
        ext[0] = th2f[0]; ext[1] = -th2f[0];
        ext[2] = lo(xi[0]); ext[3] = -ext[2]; ext[4] = lo(yi[0]); ext[5] = -ext[4]; ext[6] = lo(zi[0]); ext[7] = -ext[6];
#pragma unroll
        for (int k = 1; k < TPT; ++k) {
            ext[0] = fminf(ext[0], th2f[k]); ext[1] = fminf(ext[1], -th2f[k]);
            ext[2] = fminf(ext[2], lo(xi[k])); ext[3] = fminf(ext[3], -lo(xi[k]));
            ext[4] = fminf(ext[4], lo(yi[k])); ext[5] = fminf(ext[5], -lo(yi[k]));
            ext[6] = fminf(ext[6], lo(zi[k])); ext[7] = fminf(ext[7], -lo(zi[k]));
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ext[r] = fminf(ext[r], __shfl_xor_sync(0xffffffffu, ext[r], o));
            if ((tid & 31) == 0) s_ext[r][tid >> 5] = ext[r];
        }
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&s_full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // the block's extents stay in shared memory (read once per tile): no registers held across the packed loops
    __shared__ float s_blk[8];
    if (tid < 8) {
        float v = s_ext[tid][0];
#pragma unroll
        for (int w = 1; w < DT / 32; ++w) v = fminf(v, s_ext[tid][w]);
        s_blk[tid] = v;
    }
    __syncthreads();
    auto issue = [&](int64_t tile) {
        int st = (int)((tile - tile_begin) % STAGES);
        int64_t q0 = tile * TILEP;
        int cnt = (int)(n_pairs - q0 < TILEP ? n_pairs - q0 : TILEP);
        uint32_t bytes = (uint32_t)cnt * sizeof(Pair8);
        uint32_t hbytes = (uint32_t)((cnt * sizeof(float2) + 15) / 16 * 16);  // src_h2 is padded
        mbar_expect_tx(&s_full[st], bytes + hbytes);
        bulk_g2s(&s_src[st][0], src + q0, bytes, &s_full[st]);
        bulk_g2s(&s_h2[st][0], src_h2 + 2 * q0, hbytes, &s_full[st]);
    };
    if (tid == 0)
        for (int s = 0; s < STAGES - 1; ++s)
            if (tile_begin + s < tile_end) issue(tile_begin + s);

    const int64_t blk_lo = self_base >= 0 ? self_base + tgt_base : INT64_MAX;
    const int64_t blk_hi = self_base >= 0 ? blk_lo + DT * TPT : INT64_MIN;

    for (int64_t tile = tile_begin; tile < tile_end; ++tile) {
        const int it = (int)(tile - tile_begin);
        const int st = it % STAGES;
        if (tid == 0 && tile + STAGES - 1 < tile_end) issue(tile + STAGES - 1);
        mbar_wait(&s_full[st], (uint32_t)((it / STAGES) & 1));
        const int64_t q0 = tile * TILEP;
        const int cnt = (int)(n_pairs - q0 < TILEP ? n_pairs - q0 : TILEP);
        const int64_t j0 = 2 * q0;
        const bool diag = (j0 < blk_hi) && (j0 + 2 * cnt > blk_lo);
        const bool has_pad = CONSTM && (n_src & 1) && tile == n_tiles - 1;  // the zero-mass pad needs its mass
        if (!diag && cnt == TILEP && !has_pad) {
            f2_t ax[TPT], ay[TPT], az[TPT], p[TPT];
#pragma unroll
            for (int k = 0; k < TPT; ++k) ax[k] = ay[k] = az[k] = p[k] = 0ull;  // {+0.f, +0.f}
            // spline: bit g of `inside` = some pair of records [8g, 8g+8) lies inside its softening radius for one of
            // this lane's targets; pass 2 revisits only those groups (TILEP / 8 = 32 groups: one 32-bit mask per lane)
            unsigned inside[TPT];  // one group mask per target: pass 2 touches only (group, target) pairs that need it
#pragma unroll
            for (int k = 0; k < TPT; ++k) inside[k] = 0u;
            static_assert(TILEP == 256, "the group mask of the spline pass assumes 32 groups of 8 pair records");
            int regime = R_GENERAL;
            if (tile_meta) {
                const float4 ma = __ldg(reinterpret_cast<const float4*>(tile_meta + tile));
                const float4 mb = __ldg(reinterpret_cast<const float4*>(tile_meta + tile) + 1);
                const float blk_h2min = s_blk[0], blk_h2max = -s_blk[1];
                if (HMODE == 1) {
                    if (ma.y <= blk_h2min) regime = R_TARGET_H;
                    else if (ma.x >= blk_h2max) regime = R_SOURCE_H;
                } else {
                    // gap between the two boxes, component-wise: {lo - blk_hi, blk_lo - hi}
                    const float gx = fmaxf(0.f, fmaxf(ma.z + s_blk[3], s_blk[2] - mb.y));   // tile.lox - blk.hix, blk.lox - tile.hix
                    const float gy = fmaxf(0.f, fmaxf(ma.w + s_blk[5], s_blk[4] - mb.z));
                    const float gz = fmaxf(0.f, fmaxf(mb.x + s_blk[7], s_blk[6] - mb.w));
                    const float gap2 = gx * gx + gy * gy + gz * gz;
                    // margin: the pair distances are rounded fp32 sums (relative error ~1e-6), the test must imply
                    // r2 >= h2 for the rounded values of every pair
                    if (gap2 > fmaxf(ma.y, blk_h2max) * 1.0001f + FLT_MIN) regime = R_FAR;
                }
            }
            if (HMODE == 1 && regime == R_TARGET_H)
                f2h_tile<WANT, HMODE, CONSTM, R_TARGET_H>(s_src[st], s_h2[st], xi, yi, zi, th2f, ax, ay, az, p, inside);
            else if (HMODE == 1 && regime == R_SOURCE_H)
                f2h_tile<WANT, HMODE, CONSTM, R_SOURCE_H>(s_src[st], s_h2[st], xi, yi, zi, th2f, ax, ay, az, p, inside);
            else if (HMODE == 2 && regime == R_FAR)
                f2h_tile<WANT, HMODE, CONSTM, R_FAR>(s_src[st], s_h2[st], xi, yi, zi, th2f, ax, ay, az, p, inside);
            else
                f2h_tile<WANT, HMODE, CONSTM, R_GENERAL>(s_src[st], s_h2[st], xi, yi, zi, th2f, ax, ay, az, p, inside);
            float sax[TPT], say[TPT], saz[TPT], sp[TPT];
#pragma unroll
            for (int k = 0; k < TPT; ++k) {
                float a, b;
                sax[k] = say[k] = saz[k] = sp[k] = 0.f;
                if (WANT & PNBX_WANT_ACC) {
                    f2_unpack(ax[k], a, b); sax[k] = a + b;
                    f2_unpack(ay[k], a, b); say[k] = a + b;
                    f2_unpack(az[k], a, b); saz[k] = a + b;
                }
                if (WANT & PNBX_WANT_POT) { f2_unpack(p[k], a, b); sp[k] = -(a + b); }
            }
#ifndef PNBX_EXPERIMENT_NO_PASS2
            if (HMODE == 2) {
                // pass 2: W2 terms of the pairs with r < h, only in the flagged groups of 8 pair records (each lane walks
                // its own set bits; the predicate inside pair_h2_scalar is pass 1's, bit for bit)
#pragma unroll
                for (int k = 0; k < TPT; ++k) {
                    unsigned todo = inside[k];
                    while (todo) {
                        const int g = __ffs((int)todo) - 1;
                        todo &= todo - 1u;
#pragma unroll 2
                        for (int q = 8 * g; q < 8 * g + 8; ++q) {
                            const Pair8 pr = s_src[st][q];
                            const float2 hh = s_h2[st][q];
                            pair_h2_scalar<WANT, 2, true>(lo(xi[k]), lo(yi[k]), lo(zi[k]), th2f[k], pr.x0, pr.y0, pr.z0,
                                                          CONSTM ? 1.f : pr.m0, hh.x, false, sax[k], say[k], saz[k], sp[k]);
                            pair_h2_scalar<WANT, 2, true>(lo(xi[k]), lo(yi[k]), lo(zi[k]), th2f[k], pr.x1, pr.y1, pr.z1,
                                                          CONSTM ? 1.f : pr.m1, hh.y, false, sax[k], say[k], saz[k], sp[k]);
                        }
                    }
                }
            }
#endif
#pragma unroll
            for (int k = 0; k < TPT; ++k) {
                if (WANT & PNBX_WANT_ACC) { Ax[k] += (double)sax[k]; Ay[k] += (double)say[k]; Az[k] += (double)saz[k]; }
                if (WANT & PNBX_WANT_POT) P[k] += (double)sp[k];
            }
        } else {
            // ragged tail and/or self-skip by index: scalar loop over the same pair records
            float ax[TPT], ay[TPT], az[TPT], p[TPT];
#pragma unroll
            for (int k = 0; k < TPT; ++k) ax[k] = ay[k] = az[k] = p[k] = 0.f;
            const int nsrc_tile = (int)((n_src - j0) < 2 * cnt ? (n_src - j0) : 2 * cnt);
            for (int j = 0; j < nsrc_tile; ++j) {
                const Pair8& pr = s_src[st][j >> 1];
                const float2 hh = s_h2[st][j >> 1];
                const bool odd = j & 1;
                const float sx = odd ? pr.x1 : pr.x0, sy = odd ? pr.y1 : pr.y0, sz = odd ? pr.z1 : pr.z0;
                const float sm = CONSTM ? 1.f : (odd ? pr.m1 : pr.m0), hs2 = odd ? hh.y : hh.x;
                const int64_t gj = j0 + j;
#pragma unroll
                for (int k = 0; k < TPT; ++k)
                    pair_h2_scalar<WANT, HMODE, false>(lo(xi[k]), lo(yi[k]), lo(zi[k]), th2f[k], sx, sy, sz, sm, hs2, gj == g0 + k * DT, ax[k],
                                                       ay[k], az[k], p[k]);
            }
#pragma unroll
            for (int k = 0; k < TPT; ++k) {
                if (WANT & PNBX_WANT_ACC) { Ax[k] += (double)ax[k]; Ay[k] += (double)ay[k]; Az[k] += (double)az[k]; }
                if (WANT & PNBX_WANT_POT) P[k] += (double)p[k];
            }
        }
        __syncthreads();
    }
    const int64_t split_off = (int64_t)blockIdx.y * m;
    const double ms = CONSTM ? c_plans[plan_slot].mass_d : 1.0;
#pragma unroll
    for (int k = 0; k < TPT; ++k) {
        int64_t i = tgt_base + k * DT + tid;
        if (i < m) {
            if (WANT & PNBX_WANT_POT) out_pot[split_off + i] = P[k] * ms;
            if (WANT & PNBX_WANT_ACC) {
                double* a = out_acc + 3 * (split_off + i);
                a[0] = Ax[k] * ms; a[1] = Ay[k] * ms; a[2] = Az[k] * ms;
            }
        }
    }
}

// Extents of every source tile for direct_kernel_f2h (one block per tile, one thread per pair record).
__global__ void tile_extents(const Pair8* __restrict__ src, const float* __restrict__ h2, int64_t n_pairs,
                             TileMeta* __restrict__ out) {
    __shared__ float s_red[8][TILEP / 32];
    const int64_t q = (int64_t)blockIdx.x * TILEP + threadIdx.x;
    float v[8] = {INFINITY, INFINITY, INFINITY, INFINITY, INFINITY, INFINITY, INFINITY, INFINITY};
    if (q < n_pairs) {
        const Pair8 pr = src[q];
        const float ha = h2[2 * q], hb = h2[2 * q + 1];
        v[0] = fminf(ha, hb); v[1] = -fmaxf(ha, hb);
        v[2] = fminf(pr.x0, pr.x1); v[3] = -fmaxf(pr.x0, pr.x1);
        v[4] = fminf(pr.y0, pr.y1); v[5] = -fmaxf(pr.y0, pr.y1);
        v[6] = fminf(pr.z0, pr.z1); v[7] = -fmaxf(pr.z0, pr.z1);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[r] = fminf(v[r], __shfl_xor_sync(0xffffffffu, v[r], o));
        if ((threadIdx.x & 31) == 0) s_red[r][threadIdx.x >> 5] = v[r];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
            for (int w = 0; w < TILEP / 32; ++w) v[r] = fminf(v[r], s_red[r][w]);
        TileMeta t;
        t.h2min = v[0]; t.h2max = -v[1];
        t.lox = v[2]; t.hix = -v[3]; t.loy = v[4]; t.hiy = -v[5]; t.loz = v[6]; t.hiz = -v[7];
        out[blockIdx.x] = t;
    }
}

// Sort keys of the whole-array self calls with per-particle softenings (run_direct): Plummer — the clamped softening
// (non-negative floats order like their bit patterns); spline — a 30-bit Morton code inside the bounding box.
__device__ __forceinline__ uint32_t spread10(uint32_t x) {
    x &= 0x3ffu;
    x = (x | (x << 16)) & 0x030000ffu;
    x = (x | (x << 8)) & 0x0300f00fu;
    x = (x | (x << 4)) & 0x030c30c3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}
// Particles outside [first_lo, first_hi) get the top key bit (free in both keys): a self call on a target shard sorts
// its own particles to the front, so that target k is source k again (self_base = 0) and both parts are sorted.
__global__ void direct_sort_keys(const double* __restrict__ pos, const double* __restrict__ h, int64_t n,
                                 const double* __restrict__ bbox6, int by_position, int64_t first_lo, int64_t first_hi,
                                 uint32_t* __restrict__ key, uint32_t* __restrict__ idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    idx[i] = (uint32_t)i;
    const uint32_t behind = (i < first_lo || i >= first_hi) ? 0x80000000u : 0u;
    if (!by_position) {
        key[i] = __float_as_uint((float)fmax(h[i], 0.0)) | behind;  // non-negative floats: sign bit free
        return;
    }
    uint32_t c[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double lo = bbox6[d], ext = bbox6[3 + d] - lo;
        double u = ext > 0.0 ? (pos[3 * i + d] - lo) / ext : 0.0;
        u = fmin(fmax(u, 0.0), 1.0);
        c[d] = min(1023u, (uint32_t)(u * 1024.0));
    }
    key[i] = spread10(c[0]) | (spread10(c[1]) << 1) | (spread10(c[2]) << 2) | behind;  // 30-bit code
}
__global__ void gather_sorted(const uint32_t* __restrict__ perm, int64_t n, const double* __restrict__ pos,
                              const double* __restrict__ mass, const double* __restrict__ h, double* __restrict__ pos_s,
                              double* __restrict__ mass_s, double* __restrict__ h_s) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t i = perm[k];
    pos_s[3 * k] = pos[3 * i]; pos_s[3 * k + 1] = pos[3 * i + 1]; pos_s[3 * k + 2] = pos[3 * i + 2];
    if (mass) mass_s[k] = mass[i];
    if (h) h_s[k] = h[i];
}
// results of the sorted sweep back to the caller's order (also sums the source splits, in split order)
__global__ void scatter_results(const uint32_t* __restrict__ perm, int64_t m, int splits, const double* __restrict__ pot_s,
                                const double* __restrict__ acc_s, int64_t first, double* __restrict__ pot, double* __restrict__ acc) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const int64_t i = (int64_t)perm[k] - first;  // slot in the caller's output: original index - tgt_begin
    if (pot) {
        double v = 0.0;
        for (int sidx = 0; sidx < splits; ++sidx) v += pot_s[(int64_t)sidx * m + k];
        pot[i] = v;
    }
    if (acc) {
        double x = 0.0, y = 0.0, z = 0.0;
        for (int sidx = 0; sidx < splits; ++sidx) {
            const double* a = acc_s + 3 * ((int64_t)sidx * m + k);
            x += a[0]; y += a[1]; z += a[2];
        }
        acc[3 * i] = x; acc[3 * i + 1] = y; acc[3 * i + 2] = z;
    }
}

// squared clamped softenings for direct_kernel_f2h: out[i] = max(h_i, 0)^2, zero padding
__global__ void pack_h2(const double* __restrict__ in, int64_t n, int64_t n_padded, float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_padded) return;
    const float h = i < n ? (float)fmax(in[i], 0.0) : 0.f;
    out[i] = h * h;
}

// float64 (n,3) positions + mass -> pair records relative to the bbox centre; an odd tail gets a
// zero-mass partner far away (its r^2 is huge but finite: contributes exactly 0, never NaN).
__global__ void pack_pairs(const double* __restrict__ pos, const double* __restrict__ mass, int64_t n,
                           const double* __restrict__ bbox6, Pair8* __restrict__ out) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * q >= n) return;
    const double cx = (bbox6[0] + bbox6[3]) * 0.5, cy = (bbox6[1] + bbox6[4]) * 0.5, cz = (bbox6[2] + bbox6[5]) * 0.5;
    const int64_t i0 = 2 * q, i1 = 2 * q + 1;
    Pair8 v;
    v.x0 = (float)(pos[3 * i0] - cx); v.y0 = (float)(pos[3 * i0 + 1] - cy); v.z0 = (float)(pos[3 * i0 + 2] - cz);
    v.m0 = mass ? (float)mass[i0] : 1.f;
    if (i1 < n) {
        v.x1 = (float)(pos[3 * i1] - cx); v.y1 = (float)(pos[3 * i1 + 1] - cy); v.z1 = (float)(pos[3 * i1 + 2] - cz);
        v.m1 = mass ? (float)mass[i1] : 1.f;
    } else {
        v.x1 = 1e18f; v.y1 = 1e18f; v.z1 = 1e18f; v.m1 = 0.f;
    }
    out[q] = v;
}

// Sum the per-split partials in split order (fixed order => run-to-run deterministic).
__global__ void reduce_splits(const double* __restrict__ part, int64_t count, int splits, double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s += part[(int64_t)k * count + i];
    out[i] = s;
}

// float64 (n,3) positions [+ mass] -> Vec4<T>(x-cx, y-cy, z-cz, m|1|w0). centre read from device bbox.
template <class T>
__global__ void pack_points(const double* __restrict__ pos, const double* __restrict__ mass, int64_t n,
                            const double* __restrict__ bbox6, T w_default, Vec4<T>* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double cx = (bbox6[0] + bbox6[3]) * 0.5, cy = (bbox6[1] + bbox6[4]) * 0.5, cz = (bbox6[2] + bbox6[5]) * 0.5;
    Vec4<T> v;
    v.x = (T)(pos[3 * i + 0] - cx);
    v.y = (T)(pos[3 * i + 1] - cy);
    v.z = (T)(pos[3 * i + 2] - cz);
    v.w = mass ? (T)mass[i] : w_default;
    out[i] = v;
}
// clamp0: the cubic spline treats h <= 0 as Newtonian (kernel.rs:46-48, 72-74), and max(h_i, h_j) <= 0 iff both are,
// so clamping every h at 0 up front leaves all spline results unchanged and saves a compare per pair.
template <class T>
__global__ void pack_scalar(const double* __restrict__ in, int64_t n, int64_t n_padded, bool clamp0, T* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_padded) return;
    double v = i < n ? in[i] : 0.0;
    if (clamp0) v = fmax(v, 0.0);
    out[i] = (T)v;
}
// min / max of the softening array, to detect the (common) constant-softening case.
__global__ void minmax_scalar_blocks(const double* __restrict__ in, int64_t n, double* __restrict__ part) {
    __shared__ double smn[256], smx[256];
    double mn = INFINITY, mx = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        double v = in[i];
        mn = fmin(mn, v);
        mx = fmax(mx, v);
    }
    smn[threadIdx.x] = mn;
    smx[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            smn[threadIdx.x] = fmin(smn[threadIdx.x], smn[threadIdx.x + o]);
            smx[threadIdx.x] = fmax(smx[threadIdx.x], smx[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = smn[0]; part[2 * blockIdx.x + 1] = smx[0]; }
}
// Final min/max of the softenings and masses, then the variant decision (mirrors the case analysis of direct.rs /
// kernel.rs; see run_direct). One thread.
struct ClassifyArgs {
    const double* hpart; const double* mpart; int nparts;  // partial min/max pairs (nullable)
    int kernel, self, packed, allow_f2h, allow_constm;
};
__global__ void classify_direct(ClassifyArgs c, DirectPlan* __restrict__ plan) {
    double hmn = 0.0, hmx = 0.0, mmn = 1.0, mmx = 1.0;
    if (c.hpart) {
        hmn = INFINITY; hmx = -INFINITY;
        for (int i = 0; i < c.nparts; ++i) { hmn = fmin(hmn, c.hpart[2 * i]); hmx = fmax(hmx, c.hpart[2 * i + 1]); }
    }
    if (c.mpart) {
        mmn = INFINITY; mmx = -INFINITY;
        for (int i = 0; i < c.nparts; ++i) { mmn = fmin(mmn, c.mpart[2 * i]); mmx = fmax(mmx, c.mpart[2 * i + 1]); }
    }
    DirectPlan p;
    double he2 = 0.0;   // constant Plummer softening squared
    bool pair = false;  // per-pair softening h = max(h_i, h_j) needed
    const bool constant = hmn == hmx;
    if (c.kernel == PNBX_KERNEL_PLUMMER && c.hpart) {
        // constant h: self max(h,h) = h; points max(h,0). Negative constant: |h| in self mode (h*h), Newtonian at
        // points (direct.rs:560) — both equal he^2 below.
        if (constant) { const double he = c.self ? hmn : fmax(hmn, 0.0); he2 = he * he; }
        else pair = true;
    } else if (c.kernel == PNBX_KERNEL_SPLINE && c.hpart) {
        pair = !(constant && hmn <= 0.0);  // spline with h <= 0 is Newtonian (kernel.rs:46-48)
    }
    p.eps2_f = (float)he2 + FLT_MIN;  // + the reference's R2_TINY
    p.eps2_d = he2 + DBL_MIN;
    p.mass_d = mmn;
    p.mass_f = (float)mmn;
    if (!pair) {
        p.variant = !c.packed ? V_SCALAR_CONST : (c.allow_constm && mmn == mmx) ? V_F2_CONSTM : V_F2;
    } else {
        // packed per-pair path: needs h^2 = max(h_i^2, h_j^2), i.e. every softening that takes part >= 0. The spline
        // clamps at 0 anyway (h <= 0 is Newtonian), at-points Plummer uses max(h_j, 0) (direct.rs:560); Plummer in
        // self mode with a negative softening (max(h_i, h_j) of signed values) stays on the scalar kernel.
        const bool f2h_ok = c.packed && c.allow_f2h && (c.kernel == PNBX_KERNEL_SPLINE || !c.self || hmn >= 0.0);
        p.variant = !f2h_ok ? V_SCALAR_PAIR : (c.allow_constm && mmn == mmx) ? V_F2H_CONSTM : V_F2H;
    }
    *plan = p;
}

template <int WANT, int SOFT, class T, int TILE>
void launch_direct_t(const Vec4<T>* src, const T* src_h, int64_t n, const Vec4<T>* tgt, const T* tgt_h, int64_t m,
                     int64_t self_base, int plan, int variant, int splits, int tiles_per_split, double* pot,
                     double* acc, cudaStream_t s) {
    dim3 grid((unsigned)ceil_div(m, DT * TPT), (unsigned)splits);
    PNBX_LAUNCH((direct_kernel<WANT, SOFT, T, TILE>), grid, DT, 0, s, src, src_h, n, tgt, tgt_h, m, self_base, plan, variant,
                tiles_per_split, pot, acc);
}

template <class T, int TILE>
void launch_direct(int want, int soft, const Vec4<T>* src, const T* src_h, int64_t n, const Vec4<T>* tgt,
                   const T* tgt_h, int64_t m, int64_t self_base, int plan, int variant, int splits,
                   int tiles_per_split, double* pot, double* acc, cudaStream_t s) {
#define PNBX_CASE(W, S)                                                                                          \
    if (want == W && soft == S) {                                                                                \
        launch_direct_t<W, S, T, TILE>(src, src_h, n, tgt, tgt_h, m, self_base, plan, variant, splits,          \
                                       tiles_per_split, pot, acc, s);                                            \
        return;                                                                                                  \
    }
    PNBX_CASE(1, SOFT_PLUMMER_CONST) PNBX_CASE(2, SOFT_PLUMMER_CONST) PNBX_CASE(3, SOFT_PLUMMER_CONST)
    PNBX_CASE(1, SOFT_PLUMMER_PAIR) PNBX_CASE(2, SOFT_PLUMMER_PAIR) PNBX_CASE(3, SOFT_PLUMMER_PAIR)
    PNBX_CASE(1, SOFT_SPLINE) PNBX_CASE(2, SOFT_SPLINE) PNBX_CASE(3, SOFT_SPLINE)
#undef PNBX_CASE
    throw ArgError{PNBX_ERR_ARG, "internal: no direct kernel variant"};
}

// Constant-memory plan slots, per device: a slot is reused only after the kernels of its previous user have run
// (event), so any number of calls may be queued on any streams.
struct PlanSlots {
    std::mutex mu;
    unsigned next = 0;
    cudaEvent_t done[PLAN_SLOTS] = {};
};
PlanSlots g_plan_slots[KernelEvents::MAXDEV];

int acquire_plan_slot(int device, cudaStream_t s, const DirectPlan* d_plan) {
    PlanSlots& ps = g_plan_slots[device < KernelEvents::MAXDEV ? device : 0];
    int slot;
    {
        std::lock_guard<std::mutex> lock(ps.mu);
        slot = (int)(ps.next++ % PLAN_SLOTS);
        if (!ps.done[slot]) PNBX_CUDA(cudaEventCreateWithFlags(&ps.done[slot], cudaEventDisableTiming));
        else PNBX_CUDA(cudaStreamWaitEvent(s, ps.done[slot], 0));
    }
    PNBX_CUDA(cudaMemcpyToSymbolAsync(c_plans, d_plan, sizeof(DirectPlan), (size_t)slot * sizeof(DirectPlan),
                                      cudaMemcpyDeviceToDevice, s));
    return slot;
}
void release_plan_slot(int device, cudaStream_t s, int slot) {
    PlanSlots& ps = g_plan_slots[device < KernelEvents::MAXDEV ? device : 0];
    std::lock_guard<std::mutex> lock(ps.mu);
    PNBX_CUDA(cudaEventRecord(ps.done[slot], s));
}

template <class T, int TILE>
void run_direct(const Exec& ex, const double* d_pos, const double* d_mass, const double* d_h, int64_t n,
                const double* d_tgt, int64_t m, int64_t tgt_begin, int kernel, int want, double* d_pot,
                double* d_acc, StageTimer& tm) {
    cudaStream_t s = ex.stream;
    const bool self = d_tgt == nullptr;
    constexpr int NPART = 296;

    tm.begin("direct.pack");
    DevBuf<double> bbox(6, s);
    launch_bbox(d_pos, n, bbox.get(), s);

    // ---- variant decision on the device
    const bool soft_kernel = kernel == PNBX_KERNEL_PLUMMER || kernel == PNBX_KERNEL_SPLINE;
    const bool have_h = soft_kernel && d_h != nullptr;
    const bool packed = sizeof(T) == 4 && !getenv("PNBX_DIRECT_SCALAR");  // packed FFMA2 kernels (fp32 only)
    const bool allow_f2h = !getenv("PNBX_DIRECT_NO_F2H");
    const bool allow_constm = !getenv("PNBX_DIRECT_NO_CONSTM");
    DevBuf<double> hpart, mpart;
    DevBuf<DirectPlan> plan(1, s);
    if (have_h) {
        hpart.alloc(2 * NPART, s);
        PNBX_LAUNCH(minmax_scalar_blocks, NPART, 256, 0, s, d_h, n, hpart.get());
    }
    const bool check_mass = packed && allow_constm && d_mass != nullptr;  // unit masses are constant (direct.rs:121-128)
    if (check_mass) {
        mpart.alloc(2 * NPART, s);
        PNBX_LAUNCH(minmax_scalar_blocks, NPART, 256, 0, s, d_mass, n, mpart.get());
    }
    ClassifyArgs ca{hpart.get(), mpart.get(), NPART, kernel, self ? 1 : 0, packed ? 1 : 0, allow_f2h ? 1 : 0,
                    allow_constm ? 1 : 0};
    PNBX_LAUNCH(classify_direct, 1, 1, 0, s, ca, plan.get());
    const int slot = acquire_plan_slot(ex.device, s, plan.get());
    // which variants can the decision come out as? (everything the host can rule out is not even launched)
    const bool may_const = true;                                   // constant / absent softening
    const bool may_pair = have_h;                                  // per-pair softening
    const bool may_f2_constm = packed && may_const && allow_constm;
    const bool may_f2 = packed && may_const && (d_mass != nullptr || !allow_constm);
    const bool may_f2h = packed && may_pair && allow_f2h;
    const bool may_scalar_const = !packed;
    const bool may_scalar_pair = may_pair && (!packed || !allow_f2h || (kernel == PNBX_KERNEL_PLUMMER && self));

    // ---- per-particle softenings: sort the particles so that whole (target block, source tile) combinations resolve
    // max(h_i, h_j) / "is any pair inside its softening radius" at once (TileMeta).
    //   self calls: by softening for Plummer, along a Morton curve for the spline; the target shard's own particles
    //     (all of them in a whole-array call) are sorted to the front, so target k is source k and self_base = 0;
    //   spline at points: the order of the sources is free, so they are Morton-sorted, and so are the points (inside
    //     the sources' bounding box) — Plummer at points needs no order (every tile takes its softening from the
    //     source records).
    // The sweep runs on the sorted copies and scatter_results puts the sums back in the caller's order. The sort is
    // stable: constant softenings keep the caller's order under the Plummer key.
    const char* sort_env = getenv("PNBX_DIRECT_SORT_MIN");
    const int64_t sort_min = sort_env ? atoll(sort_env) : 65536;
    const bool can_sort = sizeof(T) == 4 && may_f2h && n >= sort_min && sort_min >= 0;
    const bool sorted = can_sort && self;                                        // sources and targets, one permutation
    const bool sorted_pts = can_sort && !self && kernel == PNBX_KERNEL_SPLINE && m < ((int64_t)1 << 31);  // sources; points with their own
    DevBuf<uint32_t> perm;  // sorted target position -> caller's target index
    DevBuf<double> pos_s, mass_s, h_s, tgt_s;
    auto sort_perm = [&](const double* p3, const double* hh, int64_t cnt, bool by_position, int64_t first_lo,
                         int64_t first_hi, DevBuf<uint32_t>& out) {
        DevBuf<uint32_t> key((size_t)cnt, s), key_s((size_t)cnt, s), idx((size_t)cnt, s);
        out.alloc((size_t)cnt, s);
        PNBX_LAUNCH(direct_sort_keys, (unsigned)ceil_div(cnt, 256), 256, 0, s, p3, hh, cnt, bbox.get(), by_position ? 1 : 0,
                    first_lo, first_hi, key.get(), idx.get());
        size_t bytes = 0;
        PNBX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, key.get(), key_s.get(), idx.get(), out.get(), (int)cnt, 0, 32, s));
        DevBuf<uint8_t> tmp(bytes, s);
        PNBX_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(), bytes, key.get(), key_s.get(), idx.get(), out.get(), (int)cnt, 0, 32, s));
        ++launch_counter();
    };
    if (sorted || sorted_pts) {
        DevBuf<uint32_t> sperm_own;
        DevBuf<uint32_t>& sperm = sorted ? perm : sperm_own;
        if (sorted) sort_perm(d_pos, d_h, n, kernel == PNBX_KERNEL_SPLINE, tgt_begin, tgt_begin + m, sperm);
        else sort_perm(d_pos, d_h, n, true, 0, n, sperm);
        pos_s.alloc((size_t)3 * n, s);
        if (d_mass) mass_s.alloc((size_t)n, s);
        h_s.alloc((size_t)n, s);
        PNBX_LAUNCH(gather_sorted, (unsigned)ceil_div(n, 256), 256, 0, s, sperm.get(), n, d_pos, d_mass, d_h, pos_s.get(),
                    mass_s.get(), h_s.get());
        d_pos = pos_s.get();
        if (d_mass) d_mass = mass_s.get();
        d_h = h_s.get();
        if (sorted_pts) {
            sort_perm(d_tgt, nullptr, m, true, 0, m, perm);
            tgt_s.alloc((size_t)3 * m, s);
            PNBX_LAUNCH(gather_sorted, (unsigned)ceil_div(m, 256), 256, 0, s, perm.get(), m, d_tgt, nullptr, nullptr, tgt_s.get(),
                        nullptr, nullptr);
            d_tgt = tgt_s.get();
        }
    }
    const bool scatter_back = sorted || sorted_pts;
    const int64_t tb = sorted ? 0 : tgt_begin;  // where the targets start in the (sorted) source arrays

    // ---- packing
    DevBuf<Vec4<T>> src4;
    DevBuf<Pair8> srcp;
    DevBuf<float> srch2;
    DevBuf<TileMeta> tmeta;
    DevBuf<T> srch;
    if (packed) {
        srcp.alloc((size_t)(n + 1) / 2, s);
        PNBX_LAUNCH(pack_pairs, (unsigned)ceil_div((n + 1) / 2, 256), 256, 0, s, d_pos, d_mass, n, bbox.get(), srcp.get());
        if (may_f2h) {
            const int64_t np = ((n + 1) / 2) * 2 + 4;  // whole pair records + room for the 16-byte rounding of the bulk copy
            srch2.alloc((size_t)np, s);
            PNBX_LAUNCH(pack_h2, (unsigned)ceil_div(np, 256), 256, 0, s, d_h, n, np, srch2.get());
            if (!getenv("PNBX_DIRECT_NO_REGIMES")) {
                const int64_t n_pairs = (n + 1) / 2;
                tmeta.alloc((size_t)ceil_div(n_pairs, TILEP), s);
                PNBX_LAUNCH(tile_extents, (unsigned)ceil_div(n_pairs, TILEP), TILEP, 0, s, srcp.get(), srch2.get(), n_pairs, tmeta.get());
            }
        }
    }
    if (may_scalar_const || may_scalar_pair) {
        src4.alloc((size_t)n, s);
        PNBX_LAUNCH(pack_points<T>, (unsigned)ceil_div(n, 256), 256, 0, s, d_pos, d_mass, n, bbox.get(), T(1), src4.get());
        if (may_scalar_pair) {
            int64_t np = ceil_div(n, 4) * 4 + 4;  // padded to 16 B granules for the bulk copy
            srch.alloc((size_t)np, s);
            PNBX_LAUNCH(pack_scalar<T>, (unsigned)ceil_div(np, 256), 256, 0, s, d_h, n, np, kernel == PNBX_KERNEL_SPLINE, srch.get());
        }
    }
    DevBuf<Vec4<T>> tgt4((size_t)m, s);
    {
        const double* tp = self ? d_pos + 3 * tb : d_tgt;
        PNBX_LAUNCH(pack_points<T>, (unsigned)ceil_div(m, 256), 256, 0, s, tp, nullptr, m, bbox.get(), T(0), tgt4.get());
    }
    PNBX_CUDA(cudaGetLastError());
    tm.end();

    // work decomposition: target blocks x source splits, >= ~20 waves of 2 CTAs/SM when possible
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ex.device);
    const int64_t n_tb = ceil_div(m, DT * TPT);
    const int64_t n_tiles = ceil_div(n, TILE);
    int64_t splits = ceil_div((int64_t)sms * 2 * 20, n_tb);
    splits = std::max<int64_t>(1, std::min<int64_t>(splits, std::min<int64_t>(64, ceil_div(n_tiles, 8))));
    int tiles_per_split = (int)ceil_div(n_tiles, splits);
    splits = ceil_div(n_tiles, tiles_per_split);

    DevBuf<double> part_pot, part_acc;
    double* kp = d_pot;
    double* ka = d_acc;
    if (splits > 1 || scatter_back) {
        if (want & PNBX_WANT_POT) { part_pot.alloc((size_t)(splits * m), s); kp = part_pot.get(); }
        if (want & PNBX_WANT_ACC) { part_acc.alloc((size_t)(splits * m * 3), s); ka = part_acc.get(); }
    }
    tm.begin("direct.kernel");
    kernel_events().begin(s);
    const int64_t sb = self ? tb : -1;
    dim3 grid((unsigned)n_tb, (unsigned)splits);
    if constexpr (sizeof(T) == 4) {
        const Pair8* sp = srcp.get();
        const Vec4<float>* tp = reinterpret_cast<const Vec4<float>*>(tgt4.get());
        const int pl = slot;
#define PNBX_F2(W)                                                                                                     \
    if (want == W) {                                                                                                   \
        if (may_f2_constm) PNBX_LAUNCH((direct_kernel_f2<W, true>), grid, DT, 0, s, sp, n, tp, m, sb, pl, tiles_per_split, kp, ka);  \
        if (may_f2) PNBX_LAUNCH((direct_kernel_f2<W, false>), grid, DT, 0, s, sp, n, tp, m, sb, pl, tiles_per_split, kp, ka);        \
        if (may_f2h) {                                                                                                 \
            const float* th2 = self ? srch2.get() + tb : nullptr; /* at-points targets have no softening */           \
            const bool cm = may_f2_constm, gm = may_f2; /* equal / general masses: same host knowledge as for f2 */    \
            const TileMeta* tmp_ = tmeta.get();                                                                        \
            if (kernel == PNBX_KERNEL_PLUMMER) {                                                                       \
                if (cm) PNBX_LAUNCH((direct_kernel_f2h<W, 1, true>), grid, DT, 0, s, sp, srch2.get(), tmp_, n, tp, th2, m, sb, pl, tiles_per_split, kp, ka); \
                if (gm) PNBX_LAUNCH((direct_kernel_f2h<W, 1, false>), grid, DT, 0, s, sp, srch2.get(), tmp_, n, tp, th2, m, sb, pl, tiles_per_split, kp, ka); \
            } else {                                                                                                   \
                if (cm) PNBX_LAUNCH((direct_kernel_f2h<W, 2, true>), grid, DT, 0, s, sp, srch2.get(), tmp_, n, tp, th2, m, sb, pl, tiles_per_split, kp, ka); \
                if (gm) PNBX_LAUNCH((direct_kernel_f2h<W, 2, false>), grid, DT, 0, s, sp, srch2.get(), tmp_, n, tp, th2, m, sb, pl, tiles_per_split, kp, ka); \
            }                                                                                                          \
        }                                                                                                              \
    }
        PNBX_F2(1) PNBX_F2(2) PNBX_F2(3)
#undef PNBX_F2
    }
    if (may_scalar_const)
        launch_direct<T, TILE>(want, SOFT_PLUMMER_CONST, src4.get(), nullptr, n, tgt4.get(), nullptr, m, sb, slot,
                               V_SCALAR_CONST, (int)splits, tiles_per_split, kp, ka, s);
    if (may_scalar_pair)
        launch_direct<T, TILE>(want, kernel == PNBX_KERNEL_SPLINE ? SOFT_SPLINE : SOFT_PLUMMER_PAIR, src4.get(), srch.get(), n,
                               tgt4.get(), self ? srch.get() + tb : nullptr, m, sb, slot, V_SCALAR_PAIR,
                               (int)splits, tiles_per_split, kp, ka, s);
    release_plan_slot(ex.device, s, slot);
    kernel_events().end(s);
    PNBX_CUDA(cudaGetLastError());
    if (scatter_back) {
        PNBX_LAUNCH(scatter_results, (unsigned)ceil_div(m, 256), 256, 0, s, perm.get(), m, (int)splits,
                    (want & PNBX_WANT_POT) ? kp : nullptr, (want & PNBX_WANT_ACC) ? ka : nullptr, sorted ? tgt_begin : 0,
                    (want & PNBX_WANT_POT) ? d_pot : nullptr, (want & PNBX_WANT_ACC) ? d_acc : nullptr);
        PNBX_CUDA(cudaGetLastError());
    } else if (splits > 1) {
        if (want & PNBX_WANT_POT)
            PNBX_LAUNCH(reduce_splits, (unsigned)ceil_div(m, 256), 256, 0, s, kp, m, (int)splits, d_pot);
        if (want & PNBX_WANT_ACC)
            PNBX_LAUNCH(reduce_splits, (unsigned)ceil_div(3 * m, 256), 256, 0, s, ka, 3 * m, (int)splits, d_acc);
        PNBX_CUDA(cudaGetLastError());
    }
    tm.end();
}

}  // namespace

// Direct sum on ONE device with device-resident float64 inputs: the body of pnbx_direct, also run per device by the
// multi-device path (multi.cu). Stream-ordered on ex.stream, no host synchronisation.
void direct_on_device(const Exec& ex, const double* d_pos, const double* d_mass, const double* d_h, int64_t n,
                      const double* d_tgt, int64_t m, int64_t tgt_begin, int kernel, int want, double* d_pot,
                      double* d_acc, StageTimer& tm) {
    if (ex.f64) run_direct<double, TILE64>(ex, d_pos, d_mass, d_h, n, d_tgt, m, tgt_begin, kernel, want, d_pot, d_acc, tm);
    else run_direct<float, TILE32>(ex, d_pos, d_mass, d_h, n, d_tgt, m, tgt_begin, kernel, want, d_pot, d_acc, tm);
}

}  // namespace pnbx

extern "C" int pnbx_direct(const double* src_pos, const double* src_mass, const double* src_h, int64_t n,
                           const double* tgt_pos, int64_t m, int64_t tgt_begin, int kernel, int want,
                           double* out_pot, double* out_acc, const pnbx_opts* opts) {
    using namespace pnbx;
    return guarded([&] {
        if (n < 0 || m < 0) throw ArgError{PNBX_ERR_ARG, "negative size"};
        if (n > 0 && !src_pos) throw ArgError{PNBX_ERR_ARG, "positions must be (N,3) float64 array"};
        if (kernel != PNBX_KERNEL_NONE && kernel != PNBX_KERNEL_PLUMMER && kernel != PNBX_KERNEL_SPLINE)
            throw ArgError{PNBX_ERR_ARG, "kernel must be 0 (Plummer) or 1 (CubicSplineW2)"};
        if (kernel == PNBX_KERNEL_NONE && src_h)  // gravity.rs:480-484
            throw ArgError{PNBX_ERR_ARG, "softenings require an explicit kernel; pass kernel=0/1 (or omit softenings)"};
        if (!(want & (PNBX_WANT_POT | PNBX_WANT_ACC)) || (want & ~3)) throw ArgError{PNBX_ERR_ARG, "bad `want` mask"};
        if ((want & PNBX_WANT_POT) && !out_pot && m > 0) throw ArgError{PNBX_ERR_ARG, "out_pot is NULL"};
        if ((want & PNBX_WANT_ACC) && !out_acc && m > 0) throw ArgError{PNBX_ERR_ARG, "out_acc is NULL"};
        const bool self = tgt_pos == nullptr;
        if (self && (tgt_begin < 0 || tgt_begin + m > n)) throw ArgError{PNBX_ERR_ARG, "target shard outside [0, N)"};
        if (n >= (int64_t)1 << 31 || m >= (int64_t)1 << 40) throw ArgError{PNBX_ERR_ARG, "problem too large"};

        // PNBX_DEVICES: whole-array host call without an explicit device -> all listed GPUs (multi.cu)
        if (n > 0 && m > 0 && tgt_begin == 0 && (!self || m == n) &&
            (!opts || (opts->mem_space == PNBX_MEM_HOST && opts->device < 0)) &&
            multi_direct(src_pos, src_mass, src_h, n, tgt_pos, m, kernel, want, out_pot, out_acc, opts))
            return;
        Exec ex = make_exec(opts);
        StageTimer tm(ex.stream);
        if (m == 0) { finish_exec(ex); return; }

        OutArray<double> o_pot, o_acc;
        if (want & PNBX_WANT_POT) o_pot.bind(out_pot, (size_t)m, ex);
        if (want & PNBX_WANT_ACC) o_acc.bind(out_acc, (size_t)3 * m, ex);
        if (n == 0) {  // direct.rs:195-197: zeros
            if (o_pot.d) PNBX_CUDA(cudaMemsetAsync(o_pot.d, 0, (size_t)m * sizeof(double), ex.stream));
            if (o_acc.d) PNBX_CUDA(cudaMemsetAsync(o_acc.d, 0, (size_t)3 * m * sizeof(double), ex.stream));
        } else {
            tm.begin("direct.h2d");
            InArray<double> i_pos, i_mass, i_h, i_tgt;
            i_pos.bind(src_pos, (size_t)3 * n, ex);
            i_mass.bind(src_mass, (size_t)n, ex);
            i_h.bind(src_h, (size_t)n, ex);
            if (!self) i_tgt.bind(tgt_pos, (size_t)3 * m, ex);
            tm.end();
            direct_on_device(ex, i_pos.d, i_mass.d, i_h.d, n, i_tgt.d, m, tgt_begin, kernel, want, o_pot.d, o_acc.d, tm);
        }
        tm.begin("direct.d2h");
        o_pot.finish(ex);
        o_acc.finish(ex);
        tm.end();
        finish_exec(ex);
    });
}
