// multipole.cuh — Cartesian multipoles to order 5 on the device.
//
// Conventions follow the reference (crates/gravity/src/multipole.rs): coefficient
// M_lmn = sum m x^l y^m z^n / (l! m! n!) about the node's centre of mass, stored in the
// reference's field order (multipole.rs:11-74) so dumps compare 1:1 with the oracle.
//   P2M  : multipole.rs:82-170      (float64, same operation order -> bit-equal payloads)
//   M2M  : multipole.rs:1536-1595   (float64, same loop and accumulation order)
//   D    : derivatives of 1/r, multipole.rs:591-856, 1216-1349 (templated: f32 walk / f64 check)
//   M2P  : multipole.rs:858-1025, 1352-1528 (potential has no dipole term; acceleration at
//          order p uses moments through order p-1 — SURVEY F9)
#pragma once
#include <cfloat>
#include <cstdint>
#include <type_traits>

namespace pnbx {
namespace mp {

// index of coefficient (l,m,n) in reference field order
enum : int {
    I000, I100, I010, I001, I200, I020, I002, I110, I101, I011, I300, I030, I003, I210, I201, I120, I102, I021, I012,
    I111, I400, I040, I004, I310, I301, I130, I103, I031, I013, I220, I202, I022, I211, I121, I112, I500, I050, I005,
    I410, I401, I140, I104, I041, I014, I320, I302, I230, I203, I032, I023, I221, I212, I122, I311, I131, I113, NCOEF
};

// coefficients kept per node for a given multipole_order (MultipoleMoments::from_full, multipole.rs:270-279)
__host__ __device__ constexpr int stored_coeffs(int order) {
    return order <= 1 ? 1 : order == 2 ? 10 : order == 3 ? 20 : order == 4 ? 35 : 56;
}

struct Lmn {
    int8_t l, m, n;
};
// exponent triples in field order
__device__ constexpr Lmn kLmn[NCOEF] = {
    {0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {2, 0, 0}, {0, 2, 0}, {0, 0, 2}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1},
    {3, 0, 0}, {0, 3, 0}, {0, 0, 3}, {2, 1, 0}, {2, 0, 1}, {1, 2, 0}, {1, 0, 2}, {0, 2, 1}, {0, 1, 2}, {1, 1, 1},
    {4, 0, 0}, {0, 4, 0}, {0, 0, 4}, {3, 1, 0}, {3, 0, 1}, {1, 3, 0}, {1, 0, 3}, {0, 3, 1}, {0, 1, 3}, {2, 2, 0},
    {2, 0, 2}, {0, 2, 2}, {2, 1, 1}, {1, 2, 1}, {1, 1, 2}, {5, 0, 0}, {0, 5, 0}, {0, 0, 5}, {4, 1, 0}, {4, 0, 1},
    {1, 4, 0}, {1, 0, 4}, {0, 4, 1}, {0, 1, 4}, {3, 2, 0}, {3, 0, 2}, {2, 3, 0}, {2, 0, 3}, {0, 3, 2}, {0, 2, 3},
    {2, 2, 1}, {2, 1, 2}, {1, 2, 2}, {3, 1, 1}, {1, 3, 1}, {1, 1, 3}};

// ---- P2M ---------------------------------------------------------------------------------------
// Each coefficient adds  ((c * mass) * f1) * f2 ...  with the factor sequence the reference
// writes (e.g. m120 += 0.5*mass*y*y*x). A factor is a coordinate (power 1) or a powi() value.
// Encoded as up to 5 tokens, token = axis*8 + power, 0 = end.
struct P2MTerm {
    double c;
    uint8_t tok[5];
};
#define PX(p) (uint8_t)(0 * 8 + (p))
#define PY(p) (uint8_t)(1 * 8 + (p))
#define PZ(p) (uint8_t)(2 * 8 + (p))
__device__ constexpr P2MTerm kP2M[NCOEF] = {
    {1.0, {0, 0, 0, 0, 0}},  // m000 += mass
    {1.0, {PX(1)}}, {1.0, {PY(1)}}, {1.0, {PZ(1)}},
    {0.5, {PX(1), PX(1)}}, {0.5, {PY(1), PY(1)}}, {0.5, {PZ(1), PZ(1)}},
    {1.0, {PX(1), PY(1)}}, {1.0, {PX(1), PZ(1)}}, {1.0, {PY(1), PZ(1)}},
    {1.0 / 6.0, {PX(3)}}, {1.0 / 6.0, {PY(3)}}, {1.0 / 6.0, {PZ(3)}},
    {0.5, {PX(1), PX(1), PY(1)}}, {0.5, {PX(1), PX(1), PZ(1)}}, {0.5, {PY(1), PY(1), PX(1)}},
    {0.5, {PX(1), PZ(1), PZ(1)}}, {0.5, {PY(1), PY(1), PZ(1)}}, {0.5, {PY(1), PZ(1), PZ(1)}},
    {1.0, {PX(1), PY(1), PZ(1)}},
    {1.0 / 24.0, {PX(4)}}, {1.0 / 24.0, {PY(4)}}, {1.0 / 24.0, {PZ(4)}},
    {1.0 / 6.0, {PX(3), PY(1)}}, {1.0 / 6.0, {PX(3), PZ(1)}}, {1.0 / 6.0, {PY(3), PX(1)}},
    {1.0 / 6.0, {PX(1), PZ(3)}}, {1.0 / 6.0, {PY(3), PZ(1)}}, {1.0 / 6.0, {PY(1), PZ(3)}},
    {0.25, {PX(1), PX(1), PY(1), PY(1)}}, {0.25, {PX(1), PX(1), PZ(1), PZ(1)}}, {0.25, {PY(1), PY(1), PZ(1), PZ(1)}},
    {0.5, {PX(1), PX(1), PY(1), PZ(1)}}, {0.5, {PY(1), PY(1), PX(1), PZ(1)}}, {0.5, {PZ(1), PZ(1), PX(1), PY(1)}},
    {1.0 / 120.0, {PX(5)}}, {1.0 / 120.0, {PY(5)}}, {1.0 / 120.0, {PZ(5)}},
    {1.0 / 24.0, {PX(4), PY(1)}}, {1.0 / 24.0, {PX(4), PZ(1)}}, {1.0 / 24.0, {PY(4), PX(1)}},
    {1.0 / 24.0, {PZ(4), PX(1)}}, {1.0 / 24.0, {PY(4), PZ(1)}}, {1.0 / 24.0, {PZ(4), PY(1)}},
    {1.0 / 12.0, {PX(3), PY(2)}}, {1.0 / 12.0, {PX(3), PZ(2)}}, {1.0 / 12.0, {PX(2), PY(3)}},
    {1.0 / 12.0, {PX(2), PZ(3)}}, {1.0 / 12.0, {PY(3), PZ(2)}}, {1.0 / 12.0, {PY(2), PZ(3)}},
    {0.25, {PX(1), PX(1), PY(1), PY(1), PZ(1)}}, {0.25, {PX(1), PX(1), PZ(1), PZ(1), PY(1)}},
    {0.25, {PY(1), PY(1), PZ(1), PZ(1), PX(1)}},
    {1.0 / 6.0, {PX(3), PY(1), PZ(1)}}, {1.0 / 6.0, {PY(3), PX(1), PZ(1)}}, {1.0 / 6.0, {PZ(3), PX(1), PY(1)}}};
#undef PX
#undef PY
#undef PZ

// ---- P2M / M2M, order known at compile time -------------------------------------------------------
// f64::powi is compiler-rt's __powidf2 (square-and-multiply): x^2 = x*x, x^3 = x*(x*x), x^4 = (x*x)*(x*x),
// x^5 = x*((x*x)*(x*x)) — the products below are written in exactly that order.
// M2M: out[lmn] = sum_{i<=l, j<=m, k<=n} (-1)^|d| shift^d / d! * child[ijk], d = (l-i, m-j, n-k), accumulated in the
// reference's loop order (multipole.rs:1544-1592); the caller's `acc` += out (add_assign, multipole.rs:173-230).
// With ORDER a template parameter every loop below unrolls, the tables
// fold to constants, moments live in registers and the only true divisions left are by 6, 12, 24, ...
// (division by 1, 2, 4 is exact, so it is issued as a multiplication). Same operation order, same bits.
template <int NC>
__device__ __forceinline__ void p2m_accumulate_ct(double (&mom)[NC], double mass, double x, double y, double z) {
    double pw[3][6];
    const double v[3] = {x, y, z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        pw[a][1] = v[a];
        pw[a][2] = __dmul_rn(v[a], v[a]);
        pw[a][3] = __dmul_rn(v[a], pw[a][2]);
        pw[a][4] = __dmul_rn(pw[a][2], pw[a][2]);
        pw[a][5] = __dmul_rn(v[a], pw[a][4]);
    }
    mom[0] = __dadd_rn(mom[0], mass);
#pragma unroll
    for (int i = 1; i < NC; ++i) {
        double t = __dmul_rn(kP2M[i].c, mass);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int tok = kP2M[i].tok[k];
            if (tok != 0) t = __dmul_rn(t, pw[tok >> 3][tok & 7]);
        }
        mom[i] = __dadd_rn(mom[i], t);
    }
}

__host__ __device__ constexpr int lmn_index_ct(int l, int m, int n) {
    const int o = l + m + n;
    const int lo = o == 0 ? 0 : o == 1 ? 1 : o == 2 ? 4 : o == 3 ? 10 : o == 4 ? 20 : 35;
    const int hi = o == 0 ? 1 : o == 1 ? 4 : o == 2 ? 10 : o == 3 ? 20 : o == 4 ? 35 : 56;
    for (int i = lo; i < hi; ++i)
        if (kLmn[i].l == l && kLmn[i].m == m) return i;
    return -1;
}
__host__ __device__ constexpr bool is_pow2_small(int d) { return d == 1 || d == 2 || d == 4 || d == 8 || d == 16; }

// compile-time loop: f(std::integral_constant<int, B>{}), ..., f(std::integral_constant<int, E-1>{})
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// The loops run at COMPILE time (static_for over integral constants): every (target, source) pair is its own
// instantiation, so exponents, signs and factorial denominators are constants, the moments stay in registers and only
// the divisions by 6, 12, 24, ... remain as divisions. A plain `#pragma unroll` nest is too large for the compiler to
// fold (it kept run-time loops, local-memory arrays and ~45 generic divisions per child at order 3).
template <int ORDER, int NC>
__device__ __forceinline__ void m2m_accumulate_ct(double (&acc)[NC], const double* __restrict__ child,
                                                  const double shift[3]) {
    double spw[3][6];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        spw[a][0] = 1.0;
        spw[a][1] = shift[a];
        spw[a][2] = __dmul_rn(shift[a], shift[a]);
        spw[a][3] = __dmul_rn(shift[a], spw[a][2]);
        spw[a][4] = __dmul_rn(spw[a][2], spw[a][2]);
        spw[a][5] = __dmul_rn(shift[a], spw[a][4]);
    }
    double ch[NC];
#pragma unroll
    for (int t = 0; t < NC; ++t) ch[t] = child[t];
    static_for<0, NC>([&](auto tc) {
        constexpr int t = decltype(tc)::value;
        constexpr int l = kLmn[t].l, m = kLmn[t].m, n = kLmn[t].n;
        if constexpr (l + m + n <= ORDER) {
            double sum = 0.0;
            static_for<0, l + 1>([&](auto ic) {
                static_for<0, m + 1>([&](auto jc) {
                    static_for<0, n + 1>([&](auto kc) {
                        constexpr int i = decltype(ic)::value, j = decltype(jc)::value, k = decltype(kc)::value;
                        constexpr int dl = l - i, dm = m - j, dn = n - k;
                        constexpr int fact[6] = {1, 1, 2, 6, 24, 120};
                        constexpr int den = fact[dl] * fact[dm] * fact[dn];
                        const double base = ch[lmn_index_ct(i, j, k)];
                        // pow = sx*sy*sz (1.0 factors are exact identities), coeff = sign*pow/(dl! dm! dn!)
                        double pw = (dl + dm + dn == 0) ? 1.0 : __dmul_rn(__dmul_rn(spw[0][dl], spw[1][dm]), spw[2][dn]);
                        if ((dl + dm + dn) & 1) pw = -pw;
                        const double coeff = is_pow2_small(den) ? __dmul_rn(pw, 1.0 / den) : __ddiv_rn(pw, (double)den);
                        const double term = __dmul_rn(coeff, base);
                        sum = (base == 0.0) ? sum : __dadd_rn(sum, term);  // the reference skips zero moments
                    });
                });
            });
            acc[t] = __dadd_rn(acc[t], sum);
        }
    });
}

// ---- derivatives of 1/r and M2P ----------------------------------------------------------------
template <class T>
__device__ __forceinline__ T inv_sqrt(T x);
template <>
__device__ __forceinline__ float inv_sqrt<float>(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <>
__device__ __forceinline__ double inv_sqrt<double>(double x) { return 1.0 / sqrt(x); }

// D[lmn] = d^(l+m+n)/dx^l dy^m dz^n (1/r) up to total order DORD, r = |(dx,dy,dz)|, tiny added to r^2.
// Built from the radial factors dt_k = (-1)^(k-1) (2k-3)!! / r^k times powers of the unit vector.
template <int DORD, class T>
__device__ __forceinline__ void derivatives(T dx, T dy, T dz, T tiny, T* D) {
    const T r2 = dx * dx + dy * dy + dz * dz + tiny;
    const T ri = inv_sqrt<T>(r2);
    const T x = dx * ri, y = dy * ri, z = dz * ri;
    const T t1 = ri;
    D[I000] = t1;
    if (DORD < 1) return;
    T t2 = -t1 * ri;
    D[I100] = t2 * x;
    D[I010] = t2 * y;
    D[I001] = t2 * z;
    if (DORD < 2) return;
    T t3 = T(-3) * t2 * ri;
    const T x2 = x * x, y2 = y * y, z2 = z * z;
    t2 *= ri;
    D[I200] = t3 * x2 + t2;
    D[I020] = t3 * y2 + t2;
    D[I002] = t3 * z2 + t2;
    D[I110] = t3 * x * y;
    D[I101] = t3 * x * z;
    D[I011] = t3 * y * z;
    if (DORD < 3) return;
    T t4 = T(-5) * t3 * ri;
    const T x3 = x2 * x, y3 = y2 * y, z3 = z2 * z;
    t3 *= ri;
    D[I300] = t4 * x3 + T(3) * t3 * x;
    D[I030] = t4 * y3 + T(3) * t3 * y;
    D[I003] = t4 * z3 + T(3) * t3 * z;
    D[I210] = t4 * x2 * y + t3 * y;
    D[I201] = t4 * x2 * z + t3 * z;
    D[I120] = t4 * y2 * x + t3 * x;
    D[I102] = t4 * z2 * x + t3 * x;
    D[I021] = t4 * y2 * z + t3 * z;
    D[I012] = t4 * z2 * y + t3 * y;
    D[I111] = t4 * x * y * z;
    if (DORD < 4) return;
    T t5 = T(-7) * t4 * ri;
    const T x4 = x3 * x, y4 = y3 * y, z4 = z3 * z;
    t3 *= ri;
    t4 *= ri;
    D[I400] = t5 * x4 + T(6) * t4 * x2 + T(3) * t3;
    D[I040] = t5 * y4 + T(6) * t4 * y2 + T(3) * t3;
    D[I004] = t5 * z4 + T(6) * t4 * z2 + T(3) * t3;
    D[I310] = t5 * x3 * y + T(3) * t4 * x * y;
    D[I301] = t5 * x3 * z + T(3) * t4 * x * z;
    D[I130] = t5 * y3 * x + T(3) * t4 * y * x;
    D[I103] = t5 * z3 * x + T(3) * t4 * x * z;
    D[I031] = t5 * y3 * z + T(3) * t4 * z * y;
    D[I013] = t5 * z3 * y + T(3) * t4 * z * y;
    D[I220] = t5 * x2 * y2 + t4 * (x2 + y2) + t3;
    D[I202] = t5 * x2 * z2 + t4 * (x2 + z2) + t3;
    D[I022] = t5 * y2 * z2 + t4 * (y2 + z2) + t3;
    D[I211] = t5 * x2 * y * z + t4 * y * z;
    D[I121] = t5 * y2 * x * z + t4 * x * z;
    D[I112] = t5 * z2 * x * y + t4 * x * y;
    if (DORD < 5) return;
    const T t6 = T(-9) * t5 * ri;
    const T x5 = x4 * x, y5 = y4 * y, z5 = z4 * z;
    t4 *= ri;
    t5 *= ri;
    D[I500] = t6 * x5 + T(10) * t5 * x3 + T(15) * t4 * x;
    D[I050] = t6 * y5 + T(10) * t5 * y3 + T(15) * t4 * y;
    D[I005] = t6 * z5 + T(10) * t5 * z3 + T(15) * t4 * z;
    D[I410] = t6 * x4 * y + T(6) * t5 * x2 * y + T(3) * t4 * y;
    D[I401] = t6 * x4 * z + T(6) * t5 * x2 * z + T(3) * t4 * z;
    D[I140] = t6 * y4 * x + T(6) * t5 * y2 * x + T(3) * t4 * x;
    D[I041] = t6 * y4 * z + T(6) * t5 * y2 * z + T(3) * t4 * z;
    D[I104] = t6 * z4 * x + T(6) * t5 * z2 * x + T(3) * t4 * x;
    D[I014] = t6 * z4 * y + T(6) * t5 * z2 * y + T(3) * t4 * y;
    D[I320] = t6 * x3 * y2 + t5 * x3 + T(3) * t5 * x * y2 + T(3) * t4 * x;
    D[I302] = t6 * x3 * z2 + t5 * x3 + T(3) * t5 * x * z2 + T(3) * t4 * x;
    D[I230] = t6 * y3 * x2 + t5 * y3 + T(3) * t5 * y * x2 + T(3) * t4 * y;
    D[I032] = t6 * y3 * z2 + t5 * y3 + T(3) * t5 * y * z2 + T(3) * t4 * y;
    D[I203] = t6 * z3 * x2 + t5 * z3 + T(3) * t5 * z * x2 + T(3) * t4 * z;
    D[I023] = t6 * z3 * y2 + t5 * z3 + T(3) * t5 * z * y2 + T(3) * t4 * z;
    D[I311] = t6 * x3 * y * z + T(3) * t5 * x * y * z;
    D[I131] = t6 * y3 * x * z + T(3) * t5 * x * y * z;
    D[I113] = t6 * z3 * x * y + T(3) * t5 * x * y * z;
    D[I122] = t6 * x * y2 * z2 + t5 * x * y2 + t5 * x * z2 + t4 * x;
    D[I212] = t6 * y * x2 * z2 + t5 * y * x2 + t5 * y * z2 + t4 * y;
    D[I221] = t6 * z * x2 * y2 + t5 * z * x2 + t5 * z * y2 + t4 * z;
}

// phi = - sum_{order != 1} M_lmn D_lmn  through `ORDER` (no dipole term, multipole.rs:866-917)
template <int ORDER, class T>
__device__ __forceinline__ T m2p_potential(const T* M, const T* D) {
    T phi = -M[I000] * D[I000];
    if (ORDER >= 2) {
#pragma unroll
        for (int i = I200; i <= I011; ++i) phi -= M[i] * D[i];
    }
    if (ORDER >= 3) {
#pragma unroll
        for (int i = I300; i <= I111; ++i) phi -= M[i] * D[i];
    }
    if (ORDER >= 4) {
#pragma unroll
        for (int i = I400; i <= I112; ++i) phi -= M[i] * D[i];
    }
    if (ORDER >= 5) {
#pragma unroll
        for (int i = I500; i <= I113; ++i) phi -= M[i] * D[i];
    }
    return phi;
}

// a_x = - sum M_lmn D_(l+1)mn etc. with moments through order ORDER-1 (multipole.rs:927-1025, 1408-1528)
template <int ORDER, class T>
__device__ __forceinline__ void m2p_accel(const T* M, const T* D, T& ax, T& ay, T& az) {
    ax = -M[I000] * D[I100];
    ay = -M[I000] * D[I010];
    az = -M[I000] * D[I001];
    if (ORDER >= 2) {
        ax -= M[I100] * D[I200] + M[I010] * D[I110] + M[I001] * D[I101];
        ay -= M[I100] * D[I110] + M[I010] * D[I020] + M[I001] * D[I011];
        az -= M[I100] * D[I101] + M[I010] * D[I011] + M[I001] * D[I002];
    }
    if (ORDER >= 3) {
        ax -= M[I200] * D[I300] + M[I020] * D[I120] + M[I002] * D[I102] + M[I110] * D[I210] + M[I101] * D[I201] + M[I011] * D[I111];
        ay -= M[I200] * D[I210] + M[I020] * D[I030] + M[I002] * D[I012] + M[I110] * D[I120] + M[I101] * D[I111] + M[I011] * D[I021];
        az -= M[I200] * D[I201] + M[I020] * D[I021] + M[I002] * D[I003] + M[I110] * D[I111] + M[I101] * D[I102] + M[I011] * D[I012];
    }
    if (ORDER >= 4) {
        ax -= M[I003] * D[I103] + M[I012] * D[I112] + M[I021] * D[I121] + M[I030] * D[I130] + M[I102] * D[I202] +
              M[I111] * D[I211] + M[I120] * D[I220] + M[I201] * D[I301] + M[I210] * D[I310] + M[I300] * D[I400];
        ay -= M[I003] * D[I013] + M[I012] * D[I022] + M[I021] * D[I031] + M[I030] * D[I040] + M[I102] * D[I112] +
              M[I111] * D[I121] + M[I120] * D[I130] + M[I201] * D[I211] + M[I210] * D[I220] + M[I300] * D[I310];
        az -= M[I003] * D[I004] + M[I012] * D[I013] + M[I021] * D[I022] + M[I030] * D[I031] + M[I102] * D[I103] +
              M[I111] * D[I112] + M[I120] * D[I121] + M[I201] * D[I202] + M[I210] * D[I211] + M[I300] * D[I301];
    }
    if (ORDER >= 5) {
        ax -= M[I004] * D[I104] + M[I013] * D[I113] + M[I022] * D[I122] + M[I031] * D[I131] + M[I040] * D[I140] +
              M[I103] * D[I203] + M[I112] * D[I212] + M[I121] * D[I221] + M[I130] * D[I230] + M[I202] * D[I302] +
              M[I211] * D[I311] + M[I220] * D[I320] + M[I301] * D[I401] + M[I310] * D[I410] + M[I400] * D[I500];
        ay -= M[I004] * D[I014] + M[I013] * D[I023] + M[I022] * D[I032] + M[I031] * D[I041] + M[I040] * D[I050] +
              M[I103] * D[I113] + M[I112] * D[I122] + M[I121] * D[I131] + M[I130] * D[I140] + M[I202] * D[I212] +
              M[I211] * D[I221] + M[I220] * D[I230] + M[I301] * D[I311] + M[I310] * D[I320] + M[I400] * D[I410];
        az -= M[I004] * D[I005] + M[I013] * D[I014] + M[I022] * D[I023] + M[I031] * D[I032] + M[I040] * D[I041] +
              M[I103] * D[I104] + M[I112] * D[I113] + M[I121] * D[I122] + M[I130] * D[I131] + M[I202] * D[I203] +
              M[I211] * D[I212] + M[I220] * D[I221] + M[I301] * D[I302] + M[I310] * D[I311] + M[I400] * D[I401];
    }
}

// ---- fp32 fast M2P for orders <= 3 ---------------------------------------------------------------
// Algebraically the same sums as m2p_potential / m2p_accel (D_ij = t3 u_i u_j + t2' d_ij,
// D_ijk = t4 u_i u_j u_k + t3' (d_ij u_k + d_ik u_j + d_jk u_i), u = d/r), contracted with the moments
// BEFORE expanding the tensors, which needs ~3x fewer instructions:
//   phi = -M/r - (3 s - trS)/r^3 + (15 C(u) - 3 w.u)/r^4
//   a   =  M d/r^3 + ((15 s - 3 trS) u - 6 S u)/r^4          (dipole about the COM is rounding noise: dropped)
// with S the symmetric second-moment matrix (S_xx = m200, S_xy = m110/2, ...), s = u.S.u,
// C(u) = sum_{l+m+n=3} M_lmn u^lmn, w_x = 3 m300 + m120 + m102 (cyclic).
// Because |u| = 1 the trace and the w terms fold into the forms themselves: with the traceless T = 3 (S - trS/3 I)
// and the cubic C'(u) = 15 C(u) - 3 (w.u)(u.u) (coefficients precomputed per node in float64, pack_walk_moments)
//   phi = -M/r - (u.T.u)/r^3 + C'(u)/r^4          a = M d/r^3 + 2 (2.5 (u.T.u) u - T u)/r^4
// Per-node record (float): [0] M, [1..6] T (xx,yy,zz,xy,xz,yz), [7] pad, [8..17] C' (x3 y3 z3 x2y | x2z xy2 xz2 y2z |
// yz2 xyz), [18,19] pad.  Floats per node: 1 (order<=1), 8 (order 2), 20 (order 3). Orders 4, 5: see m2p_fast45.
__host__ __device__ constexpr int fast_rec_floats(int order) {
    return order <= 1 ? 1 : order == 2 ? 8 : order == 3 ? 20 : order == 4 ? 48 : 84;
}

template <int ORDER, int WANT>
__device__ __forceinline__ void m2p_fast(const float* __restrict__ rec, float dx, float dy, float dz, float& pot,
                                         float& ax, float& ay, float& az) {
    const float r2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, FLT_MIN)));
    const float ri = inv_sqrt<float>(r2);
    const float ri2 = ri * ri, ri3 = ri2 * ri;
    if (ORDER <= 1) {
        const float M = __ldg(rec);
        if (WANT & 1) pot = -M * ri;
        if (WANT & 2) { const float g = M * ri3; ax = g * dx; ay = g * dy; az = g * dz; }
        return;
    }
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(rec));
    const float M = r0.x;
    const bool need_quad = (WANT & 1) || ORDER >= 3;
    const float ux = dx * ri, uy = dy * ri, uz = dz * ri;
    float qx = 0.f, qy = 0.f, qz = 0.f, st = 0.f;
    if (need_quad) {
        const float4 r1 = __ldg(reinterpret_cast<const float4*>(rec) + 1);
        const float Txx = r0.y, Tyy = r0.z, Tzz = r0.w, Txy = r1.x, Txz = r1.y, Tyz = r1.z;
        qx = fmaf(Txz, uz, fmaf(Txy, uy, Txx * ux));  // T u
        qy = fmaf(Tyz, uz, fmaf(Tyy, uy, Txy * ux));
        qz = fmaf(Tzz, uz, fmaf(Tyz, uy, Txz * ux));
        st = fmaf(qz, uz, fmaf(qy, uy, qx * ux));     // u.T.u = 3 s - trS
    }
    if (WANT & 1) {
        float phi = -M * ri;
        phi = fmaf(-ri3, st, phi);
        if (ORDER >= 3) {
            const float4 r2v = __ldg(reinterpret_cast<const float4*>(rec) + 2);
            const float4 r3v = __ldg(reinterpret_cast<const float4*>(rec) + 3);
            const float4 r4v = __ldg(reinterpret_cast<const float4*>(rec) + 4);
            const float c300 = r2v.x, c030 = r2v.y, c003 = r2v.z, c210 = r2v.w;
            const float c201 = r3v.x, c120 = r3v.y, c102 = r3v.z, c021 = r3v.w;
            const float c012 = r4v.x, c111 = r4v.y;
            const float cx = fmaf(c201, uz, fmaf(c210, uy, c300 * ux));
            const float cy = fmaf(c021, uz, fmaf(c120, ux, c030 * uy));
            const float cz = fmaf(c012, uy, fmaf(c102, ux, c003 * uz));
            float C = (ux * ux) * cx;
            C = fmaf(uy * uy, cy, C);
            C = fmaf(uz * uz, cz, C);
            C = fmaf(c111 * ux, uy * uz, C);
            phi = fmaf(ri2 * ri2, C, phi);
        }
        pot = phi;
    }
    if (WANT & 2) {
        const float g = M * ri3;
        ax = g * dx; ay = g * dy; az = g * dz;
        if (ORDER >= 3) {
            const float ri4x2 = (ri2 + ri2) * ri2;
            const float c1 = 2.5f * st;
            ax = fmaf(ri4x2, fmaf(c1, ux, -qx), ax);
            ay = fmaf(ri4x2, fmaf(c1, uy, -qy), ay);
            az = fmaf(ri4x2, fmaf(c1, uz, -qz), az);
        }
    }
}

// ---- fp32 fast M2P for orders 4 and 5 ------------------------------------------------------------
// For a homogeneous polynomial P_p(d) = sum_{|n|=p} M_n d^n the tensor contraction collapses to
//   sum_n M_n D_n(d) = sum_k c_{p,k} (Lap^k P_p)(d) / r^(2p+1-2k),   c_{p,k} = (-1)^(p-k) (2p-2k-1)!! / (2^k k!),
// (checked against the reference's d_lmn for p = 2, 3, 4), so with u = d/r:
//   Psi_2 = (3 S(u,u) - trS) / r^3                        Psi_3 = (-15 C(u) + 9 v.u) / r^4
//   Psi_4 = (105 Q(u) - 7.5 LapQ(u) + 0.375 Lap2Q) / r^5   Psi_5 = (-945 R(u) + 52.5 LapR(u) - 1.875 Lap2R.u) / r^6
//   phi = -(M/r + Psi_2 + ... + Psi_order)                  a = -grad_d (M/r + Psi_2 + ... + Psi_(order-1))
// The Laplacians' coefficients are precomputed per node (pack_walk_moments). Record (floats): [0] M, [1..6] 6S (xx,yy,zz,
// xy,xz,yz), [7] 3 trS, [8..17] octupole (field order), [18..20] 9 v = 3 w, [21..23] pad,
// [24..38] Q (m400..m112, field order), [39..44] LapQ as a quadratic form (xx,yy,zz,xy,xz,yz), [45] Lap2Q, [46,47] pad,
// [48..68] R (m500..m113), [69..78] LapR as a cubic (x3,y3,z3,x2y,x2z,xy2,xz2,y2z,yz2,xyz), [79..81] Lap2R, [82,83] pad.
template <int ORDER, int WANT>
__device__ __forceinline__ void m2p_fast45(const float* __restrict__ rec, float dx, float dy, float dz, float& pot,
                                           float& ax, float& ay, float& az) {
    static_assert(ORDER == 4 || ORDER == 5, "orders 4 and 5 only");
    float r[ORDER == 4 ? 48 : 84];
    constexpr int NV = (ORDER == 4 ? 48 : 84) / 4;
    constexpr int NEED = (WANT & 1) ? NV : (ORDER == 4 ? 6 : 12);  // acc-only needs the records of order-1 moments
#pragma unroll
    for (int v = 0; v < NEED; ++v) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(rec) + v);
        r[4 * v] = q.x; r[4 * v + 1] = q.y; r[4 * v + 2] = q.z; r[4 * v + 3] = q.w;
    }
    const float r2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, FLT_MIN)));
    const float ri = inv_sqrt<float>(r2);
    const float ri2 = ri * ri, ri3 = ri2 * ri, ri4 = ri2 * ri2, ri5 = ri4 * ri, ri6 = ri4 * ri2;
    const float a = dx * ri, b = dy * ri, c = dz * ri;
    const float a2 = a * a, b2 = b * b, c2 = c * c, ab = a * b, ac = a * c, bc = b * c;
    const float M = r[0];
    // quadrupole: q6 = 6 S u, s6 = 6 S(u,u), tr3 = 3 trS
    const float qx = fmaf(r[5], c, fmaf(r[4], b, r[1] * a));
    const float qy = fmaf(r[6], c, fmaf(r[2], b, r[4] * a));
    const float qz = fmaf(r[3], c, fmaf(r[6], b, r[5] * a));
    const float s6 = fmaf(qz, c, fmaf(qy, b, qx * a));
    const float tr3 = r[7];
    // octupole: gradient of the cubic form C(u) (C = u.gradC / 3), w3 = 9 v
    const float m300 = r[8], m030 = r[9], m003 = r[10], m210 = r[11], m201 = r[12], m120 = r[13], m102 = r[14],
                m021 = r[15], m012 = r[16], m111 = r[17];
    float gCx = 0.f, gCy = 0.f, gCz = 0.f, C3;
    if (WANT & 2) {  // the gradient is needed anyway: C = u.gradC / 3 (Euler)
        gCx = fmaf(3.f * m300, a2, fmaf(2.f * m210, ab, fmaf(2.f * m201, ac, fmaf(m120, b2, fmaf(m102, c2, m111 * bc)))));
        gCy = fmaf(3.f * m030, b2, fmaf(2.f * m120, ab, fmaf(2.f * m021, bc, fmaf(m210, a2, fmaf(m012, c2, m111 * ac)))));
        gCz = fmaf(3.f * m003, c2, fmaf(2.f * m102, ac, fmaf(2.f * m012, bc, fmaf(m201, a2, fmaf(m021, b2, m111 * ab)))));
        C3 = fmaf(gCz, c, fmaf(gCy, b, gCx * a));                         // 3 C(u)
    } else {         // potential only: the cubic form directly
        const float cx = fmaf(m201, c, fmaf(m210, b, m300 * a));
        const float cy = fmaf(m021, c, fmaf(m120, a, m030 * b));
        const float cz = fmaf(m012, b, fmaf(m102, a, m003 * c));
        C3 = 3.f * fmaf(m111 * a, bc, fmaf(c2, cz, fmaf(b2, cy, a2 * cx)));
    }
    const float wu3 = fmaf(r[20], c, fmaf(r[19], b, r[18] * a));          // 9 v.u
    float phi = 0.f;
    if (WANT & 1) {
        phi = -M * ri;
        phi = fmaf(-ri3, fmaf(0.5f, s6, -(1.f / 3.f) * tr3), phi);        // -Psi_2
        phi = fmaf(ri4, fmaf(5.f, C3, -wu3), phi);                        // -Psi_3 = (15 C - 9 v.u)/r^4
    }
    if (WANT & 2) {
        const float g = M * ri3;
        ax = g * dx; ay = g * dy; az = g * dz;
        const float c1 = fmaf(2.5f, s6, -tr3);                            // -grad Psi_2
        ax = fmaf(ri4, fmaf(c1, a, -qx), ax);
        ay = fmaf(ri4, fmaf(c1, b, -qy), ay);
        az = fmaf(ri4, fmaf(c1, c, -qz), az);
        // -grad Psi_3 = [15 gradC - 105 C u - 9 v + 45 (v.u) u] / r^5
        const float c3 = fmaf(-35.f, C3, 5.f * wu3);                      // -105 C + 45 v.u
        ax = fmaf(ri5, fmaf(15.f, gCx, fmaf(c3, a, -r[18])), ax);
        ay = fmaf(ri5, fmaf(15.f, gCy, fmaf(c3, b, -r[19])), ay);
        az = fmaf(ri5, fmaf(15.f, gCz, fmaf(c3, c, -r[20])), az);
    }
    if ((WANT & 1) || ORDER == 5) {
        // hexadecapole: gradient of the quartic form Q(u) (Q = u.gradQ / 4), LapQ(u) and its gradient, Lap2Q
        const float m400 = r[24], m040 = r[25], m004 = r[26], m310 = r[27], m301 = r[28], m130 = r[29], m103 = r[30],
                    m031 = r[31], m013 = r[32], m220 = r[33], m202 = r[34], m022 = r[35], m211 = r[36], m121 = r[37],
                    m112 = r[38];
        const float a3 = a2 * a, b3 = b2 * b, c3p = c2 * c;
        float gQx = 0.f, gQy = 0.f, gQz = 0.f, Q4;
        if ((WANT & 2) && ORDER == 5) {  // gradient needed: Q = u.gradQ / 4 (Euler)
            gQx = 4.f * m400 * a3 + 3.f * (m310 * a2 * b + m301 * a2 * c) + m130 * b3 + m103 * c3p +
                  2.f * a * (m220 * b2 + m202 * c2 + m211 * bc) + bc * (m121 * b + m112 * c);
            gQy = 4.f * m040 * b3 + 3.f * (m130 * a * b2 + m031 * b2 * c) + m310 * a3 + m013 * c3p +
                  2.f * b * (m220 * a2 + m022 * c2 + m121 * ac) + ac * (m211 * a + m112 * c);
            gQz = 4.f * m004 * c3p + 3.f * (m103 * a * c2 + m013 * b * c2) + m301 * a3 + m031 * b3 +
                  2.f * c * (m202 * a2 + m022 * b2 + m112 * ab) + ab * (m211 * a + m121 * b);
            Q4 = fmaf(gQz, c, fmaf(gQy, b, gQx * a));                     // 4 Q(u)
        } else {                          // potential only: the quartic form directly
            const float q3 = fmaf(a3, fmaf(m301, c, fmaf(m310, b, m400 * a)),
                                  fmaf(b3, fmaf(m031, c, fmaf(m130, a, m040 * b)), c3p * fmaf(m013, b, fmaf(m103, a, m004 * c))));
            const float q2 = fmaf(a2, fmaf(m211, bc, fmaf(m202, c2, m220 * b2)), fmaf(b2, fmaf(m121, ac, m022 * c2), m112 * ab * c2));
            Q4 = 4.f * (q3 + q2);
        }
        const float Lxx = r[39], Lyy = r[40], Lzz = r[41], Lxy = r[42], Lxz = r[43], Lyz = r[44], L2 = r[45];
        const float gLx = fmaf(2.f * Lxx, a, fmaf(Lxy, b, Lxz * c));      // grad LapQ
        const float gLy = fmaf(2.f * Lyy, b, fmaf(Lxy, a, Lyz * c));
        const float gLz = fmaf(2.f * Lzz, c, fmaf(Lxz, a, Lyz * b));
        const float LQ2 = fmaf(gLz, c, fmaf(gLy, b, gLx * a));            // 2 LapQ(u)
        if (WANT & 1) phi = fmaf(-ri5, fmaf(26.25f, Q4, fmaf(-3.75f, LQ2, 0.375f * L2)), phi);  // -Psi_4
        if ((WANT & 2) && ORDER == 5) {
            // -grad Psi_4 = [-105 gradQ + 945 Q u + 7.5 gradLapQ - 52.5 LapQ u + 1.875 Lap2Q u] / r^6
            const float c4 = fmaf(236.25f, Q4, fmaf(-26.25f, LQ2, 1.875f * L2));
            ax = fmaf(ri6, fmaf(-105.f, gQx, fmaf(7.5f, gLx, c4 * a)), ax);
            ay = fmaf(ri6, fmaf(-105.f, gQy, fmaf(7.5f, gLy, c4 * b)), ay);
            az = fmaf(ri6, fmaf(-105.f, gQz, fmaf(7.5f, gLz, c4 * c)), az);
        }
    }
    if ((WANT & 1) && ORDER == 5) {
        // 32-pole: quintic form R(u), cubic LapR(u), linear Lap2R.u
        const float* m = r + 48;  // m500 m050 m005 m410 m401 m140 m104 m041 m014 m320 m302 m230 m203 m032 m023 m221 m212 m122 m311 m131 m113
        const float a3 = a2 * a, b3 = b2 * b, c3p = c2 * c, a4 = a2 * a2, b4 = b2 * b2, c4p = c2 * c2;
        float R = m[0] * a4 * a + m[1] * b4 * b + m[2] * c4p * c;
        R += a4 * (m[3] * b + m[4] * c) + b4 * (m[5] * a + m[7] * c) + c4p * (m[6] * a + m[8] * b);
        R += a3 * (m[9] * b2 + m[10] * c2 + m[18] * bc) + b3 * (m[11] * a2 + m[13] * c2 + m[19] * ac) +
             c3p * (m[12] * a2 + m[14] * b2 + m[20] * ab);
        R += m[15] * a2 * b2 * c + m[16] * a2 * b * c2 + m[17] * a * b2 * c2;
        const float* l = r + 69;  // x3 y3 z3 x2y x2z xy2 xz2 y2z yz2 xyz
        const float LR = l[0] * a3 + l[1] * b3 + l[2] * c3p + a2 * (l[3] * b + l[4] * c) + b2 * (l[5] * a + l[7] * c) +
                         c2 * (l[6] * a + l[8] * b) + l[9] * a * bc;
        const float L2R = fmaf(r[81], c, fmaf(r[80], b, r[79] * a));
        phi = fmaf(-ri6, fmaf(-945.f, R, fmaf(52.5f, LR, -1.875f * L2R)), phi);  // -Psi_5
    }
    if (WANT & 1) pot = phi;
}

}  // namespace mp
}  // namespace pnbx
