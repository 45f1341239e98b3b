// gravity_oracle.cpp — CPU restatement (C++17, float64, OpenMP over targets) of the
// reference's gravity hot path. TEST INFRASTRUCTURE ONLY: nothing under
// pynbody-extras_b200/ may include, link or call this file. Only tests/, the
// smoke() check and bench.py's cpu_baseline / --impl reference legs use it.
//
// PARITY UNPINNED: the reference (Rust crate `gravity`) cannot be built in this image
// (no cargo/rustc) and its tests hold no golden vectors. This file is
// pinned only by (a) the reference's own property tests restated in tests/ and
// (b) analytic known answers. Each function cites the reference lines it follows.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -fopenmp). -ffp-contract=off keeps
// the reference's rounding: Rust never contracts a*b+c, and uses an explicit fused
// mul_add only where cited (std::fma below).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// direct.rs:7, tree.rs:36, multipole.rs:4
constexpr double R2_TINY = std::numeric_limits<double>::min();  // f64::MIN_POSITIVE
constexpr double MIN_SOFTENING = 0.0;                            // tree.rs:37
constexpr int64_t NONE = -1;                                     // usize::MAX stand-in

enum Kernel { PLUMMER = 0, SPLINE = 1 };  // kernel.rs:3-12

// Rust f64::max / f64::min semantics (IEEE maxNum) == std::fmax.
inline double rmax(double a, double b) { return std::fmax(a, b); }

// ---------------------------------------------------------------- kernel.rs:20-28
inline double multipole_min_separation_factor(int kernel) { return kernel == PLUMMER ? 2.8 : 1.0; }

// ---------------------------------------------------------------- kernel.rs:84-106
inline double w2(double u) {
    if (u < 0.5) {
        double u2 = u * u;
        double u4 = u2 * u2;
        double u5 = u4 * u;
        return (16.0 / 3.0) * u2 - (48.0 / 5.0) * u4 + (32.0 / 5.0) * u5 - 14.0 / 5.0;
    } else if (u < 1.0) {
        double inv_u = 1.0 / u;
        double u2 = u * u;
        double u3 = u2 * u;
        double u4 = u2 * u2;
        double u5 = u4 * u;
        return (1.0 / 15.0) * inv_u + (32.0 / 3.0) * u2 - 16.0 * u3 + (48.0 / 5.0) * u4 -
               (32.0 / 15.0) * u5 - 16.0 / 5.0;
    }
    return -1.0 / u;
}

// ---------------------------------------------------------------- kernel.rs:108-128
inline double w2_prime(double u) {
    if (u < 0.5) {
        double u2 = u * u;
        double u3 = u2 * u;
        double u4 = u2 * u2;
        return (32.0 / 3.0) * u - (192.0 / 5.0) * u3 + 32.0 * u4;
    } else if (u < 1.0) {
        double u2 = u * u;
        double u3 = u2 * u;
        double u4 = u2 * u2;
        return -(1.0 / 15.0) * (1.0 / u2) + (64.0 / 3.0) * u - 48.0 * u2 + (192.0 / 5.0) * u3 -
               (32.0 / 3.0) * u4;
    }
    return 1.0 / (u * u);
}

// ---------------------------------------------------------------- kernel.rs:41-56
inline double kernel_potential_per_unit_mass(int kind, double r, double h) {
    if (r == 0.0) return 0.0;
    if (kind == PLUMMER) return -1.0 / std::sqrt(r * r + h * h);
    if (h <= 0.0) return -1.0 / r;
    double h_inv = 1.0 / h;
    double u = r * h_inv;
    return w2(u) * h_inv;
}

// ---------------------------------------------------------------- kernel.rs:62-82
inline double kernel_accel_factor(int kind, double r, double h) {
    if (r == 0.0) return 0.0;
    if (kind == PLUMMER) {
        double s2 = r * r + h * h;
        return 1.0 / (std::sqrt(s2) * s2);
    }
    if (h <= 0.0) return 1.0 / (r * r * r);
    double h_inv = 1.0 / h;
    double u = r * h_inv;
    return w2_prime(u) * (h_inv * h_inv) / r;
}

// =============================================================== direct.rs
// The reference has a serial symmetric-pair path for n < 512 and a per-target path
// otherwise; both are restated because their rounding differs.
struct Vec3 {
    double v[3];
};

inline double mass_of(const double* masses, int64_t j) { return masses ? masses[j] : 1.0; }  // direct.rs:121-128

// direct.rs:115-185
void direct_accelerations(const double* pos, const double* masses, int64_t n, double* acc) {
    std::fill(acc, acc + 3 * n, 0.0);
    if (n == 0) return;
    if (n < 512) {
        for (int64_t i = 0; i < n; ++i) {
            const double* pi = pos + 3 * i;
            double mi = mass_of(masses, i);
            for (int64_t j = i + 1; j < n; ++j) {
                const double* pj = pos + 3 * j;
                double dx = pj[0] - pi[0], dy = pj[1] - pi[1], dz = pj[2] - pi[2];
                double r2 = dx * dx + dy * dy + dz * dz;
                double s2 = r2 + R2_TINY;
                double invr3 = 1.0 / (std::sqrt(s2) * s2);
                double mj = mass_of(masses, j);
                acc[3 * i + 0] += mj * dx * invr3;
                acc[3 * i + 1] += mj * dy * invr3;
                acc[3 * i + 2] += mj * dz * invr3;
                acc[3 * j + 0] -= mi * dx * invr3;
                acc[3 * j + 1] -= mi * dy * invr3;
                acc[3 * j + 2] -= mi * dz * invr3;
            }
        }
        return;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const double* pi = pos + 3 * i;
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            const double* pj = pos + 3 * j;
            double dx = pj[0] - pi[0], dy = pj[1] - pi[1], dz = pj[2] - pi[2];
            double r2 = dx * dx + dy * dy + dz * dz;
            double s2 = r2 + R2_TINY;
            double invr3 = 1.0 / (std::sqrt(s2) * s2);
            double mj = mass_of(masses, j);
            ax += mj * dx * invr3;
            ay += mj * dy * invr3;
            az += mj * dz * invr3;
        }
        acc[3 * i + 0] = ax;
        acc[3 * i + 1] = ay;
        acc[3 * i + 2] = az;
    }
}

// direct.rs:187-251 (serial and parallel bodies are identical per target)
void direct_accelerations_at_points(const double* pos, const double* masses, int64_t n_src,
                                    const double* tgt, int64_t n_tgt, double* acc) {
    std::fill(acc, acc + 3 * n_tgt, 0.0);
    if (n_tgt == 0 || n_src == 0) return;
#pragma omp parallel for schedule(static) if (n_tgt >= 512)
    for (int64_t i = 0; i < n_tgt; ++i) {
        const double* pi = tgt + 3 * i;
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int64_t j = 0; j < n_src; ++j) {
            const double* pj = pos + 3 * j;
            double dx = pj[0] - pi[0], dy = pj[1] - pi[1], dz = pj[2] - pi[2];
            double r2 = dx * dx + dy * dy + dz * dz;
            double s2 = r2 + R2_TINY;
            double invr3 = 1.0 / (std::sqrt(s2) * s2);
            double m = mass_of(masses, j);
            ax += m * dx * invr3;
            ay += m * dy * invr3;
            az += m * dz * invr3;
        }
        acc[3 * i + 0] = ax;
        acc[3 * i + 1] = ay;
        acc[3 * i + 2] = az;
    }
}

// direct.rs:255-313
void direct_potentials(const double* pos, const double* masses, int64_t n, double* pot) {
    std::fill(pot, pot + n, 0.0);
    if (n == 0) return;
    if (n < 512) {
        for (int64_t i = 0; i < n; ++i) {
            const double* pi = pos + 3 * i;
            double mi = mass_of(masses, i);
            for (int64_t j = i + 1; j < n; ++j) {
                const double* pj = pos + 3 * j;
                double dx = pj[0] - pi[0], dy = pj[1] - pi[1], dz = pj[2] - pi[2];
                double r2 = dx * dx + dy * dy + dz * dz;
                double invr = 1.0 / std::sqrt(r2 + R2_TINY);
                double mj = mass_of(masses, j);
                double phi_pair = -invr;
                pot[i] += phi_pair * mj;
                pot[j] += phi_pair * mi;
            }
        }
        return;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const double* pi = pos + 3 * i;
        double phi = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            const double* pj = pos + 3 * j;
            double dx = pj[0] - pi[0], dy = pj[1] - pi[1], dz = pj[2] - pi[2];
            double r2 = dx * dx + dy * dy + dz * dz;
            double invr = 1.0 / std::sqrt(r2 + R2_TINY);
            double mj = mass_of(masses, j);
            phi += -mj * invr;
        }
        pot[i] = phi;
    }
}

// direct.rs:315-368
void direct_potentials_at_points(const double* pos, const double* masses, int64_t n_src,
                                 const double* tgt, int64_t n_tgt, double* pot) {
    std::fill(pot, pot + n_tgt, 0.0);
    if (n_tgt == 0 || n_src == 0) return;
#pragma omp parallel for schedule(static) if (n_tgt >= 512)
    for (int64_t i = 0; i < n_tgt; ++i) {
        const double* pi = tgt + 3 * i;
        double phi = 0.0;
        for (int64_t j = 0; j < n_src; ++j) {
            const double* pj = pos + 3 * j;
            double dx = pj[0] - pi[0], dy = pj[1] - pi[1], dz = pj[2] - pi[2];
            double r2 = dx * dx + dy * dy + dz * dz;
            double invr = 1.0 / std::sqrt(r2 + R2_TINY);
            double m = mass_of(masses, j);
            phi += -m * invr;
        }
        pot[i] = phi;
    }
}

inline double soft_of(const double* hs, int64_t j) { return hs ? hs[j] : 0.0; }  // direct.rs:396,401

// direct.rs:370-441
void direct_potentials_kernel(const double* pos, const double* masses, const double* hs, int64_t n,
                              int kernel, double* pot) {
    std::fill(pot, pot + n, 0.0);
    if (n == 0) return;
    if (n < 512) {
        for (int64_t i = 0; i < n; ++i) {
            const double* pi = pos + 3 * i;
            double mi = mass_of(masses, i);
            double hi = soft_of(hs, i);
            for (int64_t j = i + 1; j < n; ++j) {
                const double* pj = pos + 3 * j;
                double mj = mass_of(masses, j);
                double hj = soft_of(hs, j);
                double h = rmax(hi, hj);
                double dx = pj[0] - pi[0], dy = pj[1] - pi[1], dz = pj[2] - pi[2];
                double r2 = dx * dx + dy * dy + dz * dz;
                double r = std::sqrt(r2 + R2_TINY);
                double phi_pair = kernel_potential_per_unit_mass(kernel, r, h);
                pot[i] += mj * phi_pair;
                pot[j] += mi * phi_pair;
            }
        }
        return;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const double* pi = pos + 3 * i;
        double hi = soft_of(hs, i);
        double phi = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            const double* pj = pos + 3 * j;
            double hj = soft_of(hs, j);
            double h = rmax(hi, hj);
            double dx = pj[0] - pi[0], dy = pj[1] - pi[1], dz = pj[2] - pi[2];
            double r2 = dx * dx + dy * dy + dz * dz;
            double r = std::sqrt(r2 + R2_TINY);
            double phi_ij = kernel_potential_per_unit_mass(kernel, r, h);
            phi += mass_of(masses, j) * phi_ij;
        }
        pot[i] = phi;
    }
}

// direct.rs:443-524
void direct_accelerations_kernel(const double* pos, const double* masses, const double* hs,
                                 int64_t n, int kernel, double* acc) {
    std::fill(acc, acc + 3 * n, 0.0);
    if (n == 0) return;
    if (n < 512) {
        for (int64_t i = 0; i < n; ++i) {
            const double* pi = pos + 3 * i;
            double mi = mass_of(masses, i);
            double hi = soft_of(hs, i);
            for (int64_t j = i + 1; j < n; ++j) {
                const double* pj = pos + 3 * j;
                double mj = mass_of(masses, j);
                double hj = soft_of(hs, j);
                double h = rmax(hi, hj);
                double dx = pj[0] - pi[0], dy = pj[1] - pi[1], dz = pj[2] - pi[2];
                double r2 = dx * dx + dy * dy + dz * dz;
                double r = std::sqrt(r2 + R2_TINY);
                double g = kernel_accel_factor(kernel, r, h);
                acc[3 * i + 0] += mj * dx * g;
                acc[3 * i + 1] += mj * dy * g;
                acc[3 * i + 2] += mj * dz * g;
                acc[3 * j + 0] -= mi * dx * g;
                acc[3 * j + 1] -= mi * dy * g;
                acc[3 * j + 2] -= mi * dz * g;
            }
        }
        return;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const double* pi = pos + 3 * i;
        double hi = soft_of(hs, i);
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            const double* pj = pos + 3 * j;
            double hj = soft_of(hs, j);
            double h = rmax(hi, hj);
            double dx = pj[0] - pi[0], dy = pj[1] - pi[1], dz = pj[2] - pi[2];
            double r2 = dx * dx + dy * dy + dz * dz;
            double r = std::sqrt(r2 + R2_TINY);
            double g = kernel_accel_factor(kernel, r, h);
            double mj = mass_of(masses, j);
            ax += mj * dx * g;
            ay += mj * dy * g;
            az += mj * dz * g;
        }
        acc[3 * i + 0] = ax;
        acc[3 * i + 1] = ay;
        acc[3 * i + 2] = az;
    }
}

// direct.rs:526-585
void direct_potentials_kernel_at_points(const double* pos, const double* masses, const double* hs,
                                        int64_t n_src, const double* tgt, int64_t n_tgt, int kernel,
                                        double* pot) {
    std::fill(pot, pot + n_tgt, 0.0);
    if (n_tgt == 0 || n_src == 0) return;
#pragma omp parallel for schedule(static) if (n_tgt >= 512)
    for (int64_t i = 0; i < n_tgt; ++i) {
        const double* pi = tgt + 3 * i;
        double phi = 0.0;
        for (int64_t j = 0; j < n_src; ++j) {
            const double* pj = pos + 3 * j;
            double dx = pj[0] - pi[0], dy = pj[1] - pi[1], dz = pj[2] - pi[2];
            double r2 = dx * dx + dy * dy + dz * dz;
            double r = std::sqrt(r2 + R2_TINY);
            double hj = soft_of(hs, j);
            double h = rmax(hj, 0.0);
            phi += mass_of(masses, j) * kernel_potential_per_unit_mass(kernel, r, h);
        }
        pot[i] = phi;
    }
}

// direct.rs:587-658
void direct_accelerations_kernel_at_points(const double* pos, const double* masses,
                                           const double* hs, int64_t n_src, const double* tgt,
                                           int64_t n_tgt, int kernel, double* acc) {
    std::fill(acc, acc + 3 * n_tgt, 0.0);
    if (n_tgt == 0 || n_src == 0) return;
#pragma omp parallel for schedule(static) if (n_tgt >= 512)
    for (int64_t i = 0; i < n_tgt; ++i) {
        const double* pi = tgt + 3 * i;
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int64_t j = 0; j < n_src; ++j) {
            const double* pj = pos + 3 * j;
            double dx = pj[0] - pi[0], dy = pj[1] - pi[1], dz = pj[2] - pi[2];
            double r2 = dx * dx + dy * dy + dz * dz;
            double r = std::sqrt(r2 + R2_TINY);
            double hj = soft_of(hs, j);
            double h = rmax(hj, 0.0);
            double g = kernel_accel_factor(kernel, r, h);
            double mj = mass_of(masses, j);
            ax += mj * dx * g;
            ay += mj * dy * g;
            az += mj * dz * g;
        }
        acc[3 * i + 0] = ax;
        acc[3 * i + 1] = ay;
        acc[3 * i + 2] = az;
    }
}

// =============================================================== multipole.rs
// Field order of MultipoleMoment (multipole.rs:11-74); PotentialDerivatives uses the
// same order (multipole.rs:1157-1214).
#define PNBX_LMN_LIST(X)                                                                         \
    X(000) X(100) X(010) X(001) X(200) X(020) X(002) X(110) X(101) X(011) X(300) X(030) X(003)  \
    X(210) X(201) X(120) X(102) X(021) X(012) X(111) X(400) X(040) X(004) X(310) X(301) X(130)  \
    X(103) X(031) X(013) X(220) X(202) X(022) X(211) X(121) X(112) X(500) X(050) X(005) X(410)  \
    X(401) X(140) X(104) X(041) X(014) X(320) X(302) X(230) X(203) X(032) X(023) X(221) X(212)  \
    X(122) X(311) X(131) X(113)

struct Moment {
#define X(t) double m##t = 0.0;
    PNBX_LMN_LIST(X)
#undef X
};
struct Deriv {
#define X(t) double d##t = 0.0;
    PNBX_LMN_LIST(X)
#undef X
};
static_assert(sizeof(Moment) == 56 * sizeof(double), "56 coefficients");

// exponent triples in field order
constexpr int LMN[56][3] = {
#define X(t) {(0##t / 64) % 8, (0##t / 8) % 8, 0##t % 8},  // octal literal: digits are l,m,n
    PNBX_LMN_LIST(X)
#undef X
};

inline double* mfield(Moment& m, int idx) { return reinterpret_cast<double*>(&m) + idx; }
inline const double* mfield(const Moment& m, int idx) { return reinterpret_cast<const double*>(&m) + idx; }

// index of (l,m,n) in field order, -1 if l+m+n > 5  (get_moment/set_moment, multipole.rs:1031-1153)
struct LmnIndex {
    int idx[6][6][6];
    LmnIndex() {
        for (auto& a : idx) for (auto& b : a) for (auto& c : b) c = -1;
        for (int i = 0; i < 56; ++i) idx[LMN[i][0]][LMN[i][1]][LMN[i][2]] = i;
    }
};
const LmnIndex LMN_INDEX;

// Rust f64::powi lowers to compiler-rt __powidf2 (square-and-multiply).
inline double powi(double a, int b) {
    double r = 1.0;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return r;
}

// multipole.rs:82-170 (from_points_const<O>), :237-252
Moment moment_from_points(const double* positions, const double* masses, const int64_t* indices,
                          int64_t count, const double center[3], int order) {
    const int O = std::min(order, 5);
    Moment m;
    for (int64_t k = 0; k < count; ++k) {
        int64_t pi = indices[k];
        const double* p = positions + 3 * pi;
        double mass = masses ? masses[pi] : 1.0;
        double x = p[0] - center[0];
        double y = p[1] - center[1];
        double z = p[2] - center[2];
        m.m000 += mass;
        if (O >= 1) {
            m.m100 += mass * x;
            m.m010 += mass * y;
            m.m001 += mass * z;
        }
        if (O >= 2) {
            m.m200 += 0.5 * mass * x * x;
            m.m020 += 0.5 * mass * y * y;
            m.m002 += 0.5 * mass * z * z;
            m.m110 += mass * x * y;
            m.m101 += mass * x * z;
            m.m011 += mass * y * z;
        }
        if (O >= 3) {
            m.m300 += (1.0 / 6.0) * mass * powi(x, 3);
            m.m030 += (1.0 / 6.0) * mass * powi(y, 3);
            m.m003 += (1.0 / 6.0) * mass * powi(z, 3);
            m.m210 += 0.5 * mass * x * x * y;
            m.m201 += 0.5 * mass * x * x * z;
            m.m120 += 0.5 * mass * y * y * x;
            m.m102 += 0.5 * mass * x * z * z;
            m.m021 += 0.5 * mass * y * y * z;
            m.m012 += 0.5 * mass * y * z * z;
            m.m111 += mass * x * y * z;
        }
        if (O >= 4) {
            m.m400 += (1.0 / 24.0) * mass * powi(x, 4);
            m.m040 += (1.0 / 24.0) * mass * powi(y, 4);
            m.m004 += (1.0 / 24.0) * mass * powi(z, 4);
            m.m310 += (1.0 / 6.0) * mass * powi(x, 3) * y;
            m.m301 += (1.0 / 6.0) * mass * powi(x, 3) * z;
            m.m130 += (1.0 / 6.0) * mass * powi(y, 3) * x;
            m.m103 += (1.0 / 6.0) * mass * x * powi(z, 3);
            m.m031 += (1.0 / 6.0) * mass * powi(y, 3) * z;
            m.m013 += (1.0 / 6.0) * mass * y * powi(z, 3);
            m.m220 += 0.25 * mass * x * x * y * y;
            m.m202 += 0.25 * mass * x * x * z * z;
            m.m022 += 0.25 * mass * y * y * z * z;
            m.m211 += 0.5 * mass * x * x * y * z;
            m.m121 += 0.5 * mass * y * y * x * z;
            m.m112 += 0.5 * mass * z * z * x * y;
        }
        if (O >= 5) {
            m.m500 += (1.0 / 120.0) * mass * powi(x, 5);
            m.m050 += (1.0 / 120.0) * mass * powi(y, 5);
            m.m005 += (1.0 / 120.0) * mass * powi(z, 5);
            m.m410 += (1.0 / 24.0) * mass * powi(x, 4) * y;
            m.m401 += (1.0 / 24.0) * mass * powi(x, 4) * z;
            m.m140 += (1.0 / 24.0) * mass * powi(y, 4) * x;
            m.m104 += (1.0 / 24.0) * mass * powi(z, 4) * x;
            m.m041 += (1.0 / 24.0) * mass * powi(y, 4) * z;
            m.m014 += (1.0 / 24.0) * mass * powi(z, 4) * y;
            m.m320 += (1.0 / 12.0) * mass * powi(x, 3) * powi(y, 2);
            m.m302 += (1.0 / 12.0) * mass * powi(x, 3) * powi(z, 2);
            m.m230 += (1.0 / 12.0) * mass * powi(x, 2) * powi(y, 3);
            m.m203 += (1.0 / 12.0) * mass * powi(x, 2) * powi(z, 3);
            m.m032 += (1.0 / 12.0) * mass * powi(y, 3) * powi(z, 2);
            m.m023 += (1.0 / 12.0) * mass * powi(y, 2) * powi(z, 3);
            m.m221 += 0.25 * mass * x * x * y * y * z;
            m.m212 += 0.25 * mass * x * x * z * z * y;
            m.m122 += 0.25 * mass * y * y * z * z * x;
            m.m311 += (1.0 / 6.0) * mass * powi(x, 3) * y * z;
            m.m131 += (1.0 / 6.0) * mass * powi(y, 3) * x * z;
            m.m113 += (1.0 / 6.0) * mass * powi(z, 3) * x * y;
        }
    }
    return m;
}

// multipole.rs:173-230
inline void moment_add_assign(Moment& a, const Moment& b) {
    for (int i = 0; i < 56; ++i) *mfield(a, i) += *mfield(b, i);
}

// multipole.rs:1028, 1536-1595
constexpr double FACT[6] = {1.0, 1.0, 2.0, 6.0, 24.0, 120.0};
Moment translate_multipole(const Moment& child, const double shift[3], int order) {
    const int o = std::min(order, 5);
    Moment out;
    for (int l = 0; l <= o; ++l)
        for (int mm = 0; mm <= o; ++mm)
            for (int n = 0; n <= o; ++n) {
                if (l + mm + n > o) continue;
                double sum = 0.0;
                for (int i = 0; i <= l; ++i)
                    for (int j = 0; j <= mm; ++j)
                        for (int k = 0; k <= n; ++k) {
                            double base = *mfield(child, LMN_INDEX.idx[i][j][k]);
                            if (base == 0.0) continue;
                            int dl = l - i, dm = mm - j, dn = n - k;
                            double pw;
                            if (dl + dm + dn == 0) {
                                pw = 1.0;
                            } else {
                                double sx = dl > 0 ? powi(shift[0], dl) : 1.0;
                                double sy = dm > 0 ? powi(shift[1], dm) : 1.0;
                                double sz = dn > 0 ? powi(shift[2], dn) : 1.0;
                                pw = sx * sy * sz;
                            }
                            double sign = ((dl + dm + dn) % 2 == 0) ? 1.0 : -1.0;
                            double coeff = sign * pw / (FACT[dl] * FACT[dm] * FACT[dn]);
                            sum += coeff * base;
                        }
                *mfield(out, LMN_INDEX.idx[l][mm][n]) = sum;
            }
    return out;
}

// Number of stored coefficients after MultipoleMoments::from_full (multipole.rs:270-279,
// 300-377): order 0|1 -> 1, 2 -> 10, 3 -> 20, 4 -> 35, 5 -> 56.
inline int stored_coeffs(int order) {
    switch (std::min(order, 5)) {
        case 0: case 1: return 1;
        case 2: return 10;
        case 3: return 20;
        case 4: return 35;
        default: return 56;
    }
}
inline void compact_moment(Moment& m, int order) {  // drop what from_full drops
    int keep = stored_coeffs(order);
    for (int i = keep; i < 56; ++i) *mfield(m, i) = 0.0;
}

// multipole.rs:591-612 (order 1), 629-673 (2), 700-771 (3), 1216-1349 (generic, used for 4, 5)
Deriv derivatives(double dx, double dy, double dz, double eps2, int order) {
    const int max = std::min(order, 5);
    Deriv d;
    double r2 = dx * dx + dy * dy + dz * dz + eps2 + R2_TINY;
    double r = std::sqrt(r2);
    double r_inv = 1.0 / r;

    double dt_1 = r_inv;
    double dt_2 = -dt_1 * r_inv;
    double dt_3 = -3.0 * dt_2 * r_inv;
    double dt_4 = -5.0 * dt_3 * r_inv;
    double dt_5 = -7.0 * dt_4 * r_inv;
    double dt_6 = -9.0 * dt_5 * r_inv;

    double rx_r = dx * r_inv, ry_r = dy * r_inv, rz_r = dz * r_inv;
    double rx_r2 = rx_r * rx_r, ry_r2 = ry_r * ry_r, rz_r2 = rz_r * rz_r;
    double rx_r3 = rx_r2 * rx_r, ry_r3 = ry_r2 * ry_r, rz_r3 = rz_r2 * rz_r;
    double rx_r4 = rx_r3 * rx_r, ry_r4 = ry_r3 * ry_r, rz_r4 = rz_r3 * rz_r;
    double rx_r5 = rx_r4 * rx_r, ry_r5 = ry_r4 * ry_r, rz_r5 = rz_r4 * rz_r;

    d.d000 = dt_1;
    if (max == 0) return d;

    d.d100 = dt_2 * rx_r;
    d.d010 = dt_2 * ry_r;
    d.d001 = dt_2 * rz_r;
    if (max == 1) return d;

    dt_2 *= r_inv;
    d.d200 = dt_3 * rx_r2 + dt_2;
    d.d020 = dt_3 * ry_r2 + dt_2;
    d.d002 = dt_3 * rz_r2 + dt_2;
    d.d110 = dt_3 * rx_r * ry_r;
    d.d101 = dt_3 * rx_r * rz_r;
    d.d011 = dt_3 * ry_r * rz_r;
    if (max == 2) return d;

    dt_3 *= r_inv;
    d.d300 = dt_4 * rx_r3 + 3.0 * dt_3 * rx_r;
    d.d030 = dt_4 * ry_r3 + 3.0 * dt_3 * ry_r;
    d.d003 = dt_4 * rz_r3 + 3.0 * dt_3 * rz_r;
    d.d210 = dt_4 * rx_r2 * ry_r + dt_3 * ry_r;
    d.d201 = dt_4 * rx_r2 * rz_r + dt_3 * rz_r;
    d.d120 = dt_4 * ry_r2 * rx_r + dt_3 * rx_r;
    d.d102 = dt_4 * rz_r2 * rx_r + dt_3 * rx_r;
    d.d021 = dt_4 * ry_r2 * rz_r + dt_3 * rz_r;
    d.d012 = dt_4 * rz_r2 * ry_r + dt_3 * ry_r;
    d.d111 = dt_4 * rx_r * ry_r * rz_r;
    if (max == 3) return d;

    dt_3 *= r_inv;
    dt_4 *= r_inv;
    d.d400 = dt_5 * rx_r4 + 6.0 * dt_4 * rx_r2 + 3.0 * dt_3;
    d.d040 = dt_5 * ry_r4 + 6.0 * dt_4 * ry_r2 + 3.0 * dt_3;
    d.d004 = dt_5 * rz_r4 + 6.0 * dt_4 * rz_r2 + 3.0 * dt_3;
    d.d310 = dt_5 * rx_r3 * ry_r + 3.0 * dt_4 * rx_r * ry_r;
    d.d301 = dt_5 * rx_r3 * rz_r + 3.0 * dt_4 * rx_r * rz_r;
    d.d130 = dt_5 * ry_r3 * rx_r + 3.0 * dt_4 * ry_r * rx_r;
    d.d103 = dt_5 * rz_r3 * rx_r + 3.0 * dt_4 * rx_r * rz_r;
    d.d031 = dt_5 * ry_r3 * rz_r + 3.0 * dt_4 * rz_r * ry_r;
    d.d013 = dt_5 * rz_r3 * ry_r + 3.0 * dt_4 * rz_r * ry_r;
    d.d220 = dt_5 * rx_r2 * ry_r2 + dt_4 * (rx_r2 + ry_r2) + dt_3;
    d.d202 = dt_5 * rx_r2 * rz_r2 + dt_4 * (rx_r2 + rz_r2) + dt_3;
    d.d022 = dt_5 * ry_r2 * rz_r2 + dt_4 * (ry_r2 + rz_r2) + dt_3;
    d.d211 = dt_5 * rx_r2 * ry_r * rz_r + dt_4 * ry_r * rz_r;
    d.d121 = dt_5 * ry_r2 * rx_r * rz_r + dt_4 * rx_r * rz_r;
    d.d112 = dt_5 * rz_r2 * rx_r * ry_r + dt_4 * rx_r * ry_r;
    if (max == 4) return d;

    dt_4 *= r_inv;
    dt_5 *= r_inv;
    d.d500 = dt_6 * rx_r5 + 10.0 * dt_5 * rx_r3 + 15.0 * dt_4 * rx_r;
    d.d050 = dt_6 * ry_r5 + 10.0 * dt_5 * ry_r3 + 15.0 * dt_4 * ry_r;
    d.d005 = dt_6 * rz_r5 + 10.0 * dt_5 * rz_r3 + 15.0 * dt_4 * rz_r;
    d.d410 = dt_6 * rx_r4 * ry_r + 6.0 * dt_5 * rx_r2 * ry_r + 3.0 * dt_4 * ry_r;
    d.d401 = dt_6 * rx_r4 * rz_r + 6.0 * dt_5 * rx_r2 * rz_r + 3.0 * dt_4 * rz_r;
    d.d140 = dt_6 * ry_r4 * rx_r + 6.0 * dt_5 * ry_r2 * rx_r + 3.0 * dt_4 * rx_r;
    d.d041 = dt_6 * ry_r4 * rz_r + 6.0 * dt_5 * ry_r2 * rz_r + 3.0 * dt_4 * rz_r;
    d.d104 = dt_6 * rz_r4 * rx_r + 6.0 * dt_5 * rz_r2 * rx_r + 3.0 * dt_4 * rx_r;
    d.d014 = dt_6 * rz_r4 * ry_r + 6.0 * dt_5 * rz_r2 * ry_r + 3.0 * dt_4 * ry_r;
    d.d320 = dt_6 * rx_r3 * ry_r2 + dt_5 * rx_r3 + 3.0 * dt_5 * rx_r * ry_r2 + 3.0 * dt_4 * rx_r;
    d.d302 = dt_6 * rx_r3 * rz_r2 + dt_5 * rx_r3 + 3.0 * dt_5 * rx_r * rz_r2 + 3.0 * dt_4 * rx_r;
    d.d230 = dt_6 * ry_r3 * rx_r2 + dt_5 * ry_r3 + 3.0 * dt_5 * ry_r * rx_r2 + 3.0 * dt_4 * ry_r;
    d.d032 = dt_6 * ry_r3 * rz_r2 + dt_5 * ry_r3 + 3.0 * dt_5 * ry_r * rz_r2 + 3.0 * dt_4 * ry_r;
    d.d203 = dt_6 * rz_r3 * rx_r2 + dt_5 * rz_r3 + 3.0 * dt_5 * rz_r * rx_r2 + 3.0 * dt_4 * rz_r;
    d.d023 = dt_6 * rz_r3 * ry_r2 + dt_5 * rz_r3 + 3.0 * dt_5 * rz_r * ry_r2 + 3.0 * dt_4 * rz_r;
    d.d311 = dt_6 * rx_r3 * ry_r * rz_r + 3.0 * dt_5 * rx_r * ry_r * rz_r;
    d.d131 = dt_6 * ry_r3 * rx_r * rz_r + 3.0 * dt_5 * rx_r * ry_r * rz_r;
    d.d113 = dt_6 * rz_r3 * rx_r * ry_r + 3.0 * dt_5 * rx_r * ry_r * rz_r;
    d.d122 = dt_6 * rx_r * ry_r2 * rz_r2 + dt_5 * rx_r * ry_r2 + dt_5 * rx_r * rz_r2 + dt_4 * rx_r;
    d.d212 = dt_6 * ry_r * rx_r2 * rz_r2 + dt_5 * ry_r * rx_r2 + dt_5 * ry_r * rz_r2 + dt_4 * ry_r;
    d.d221 = dt_6 * rz_r * rx_r2 * ry_r2 + dt_5 * rz_r * rx_r2 + dt_5 * rz_r * ry_r2 + dt_4 * rz_r;
    return d;
}

// multipole.rs:858-917 (o0_d1 .. o4_d4), 1352-1405 (generic incl. order 5). No dipole term.
double potential_multipole(const Moment& m, const Deriv& d, int order) {
    const int o = std::min(order, 5);
    double phi = -m.m000 * d.d000;
    if (o <= 1) return phi;
    phi -= m.m200 * d.d200 + m.m020 * d.d020 + m.m002 * d.d002;
    phi -= m.m110 * d.d110 + m.m101 * d.d101 + m.m011 * d.d011;
    if (o == 2) return phi;
    phi -= m.m300 * d.d300 + m.m030 * d.d030 + m.m003 * d.d003;
    phi -= m.m210 * d.d210 + m.m201 * d.d201 + m.m120 * d.d120;
    phi -= m.m102 * d.d102 + m.m021 * d.d021 + m.m012 * d.d012;
    phi -= m.m111 * d.d111;
    if (o == 3) return phi;
    phi -= m.m400 * d.d400 + m.m040 * d.d040 + m.m004 * d.d004;
    phi -= m.m310 * d.d310 + m.m301 * d.d301 + m.m130 * d.d130;
    phi -= m.m103 * d.d103 + m.m031 * d.d031 + m.m013 * d.d013;
    phi -= m.m220 * d.d220 + m.m202 * d.d202 + m.m022 * d.d022;
    phi -= m.m211 * d.d211 + m.m121 * d.d121 + m.m112 * d.d112;
    if (o == 4) return phi;
    phi -= m.m500 * d.d500 + m.m050 * d.d050 + m.m005 * d.d005;
    phi -= m.m410 * d.d410 + m.m401 * d.d401 + m.m140 * d.d140;
    phi -= m.m104 * d.d104 + m.m041 * d.d041 + m.m014 * d.d014;
    phi -= m.m320 * d.d320 + m.m302 * d.d302 + m.m230 * d.d230;
    phi -= m.m203 * d.d203 + m.m032 * d.d032 + m.m023 * d.d023;
    phi -= m.m221 * d.d221 + m.m212 * d.d212 + m.m122 * d.d122;
    phi -= m.m311 * d.d311 + m.m131 * d.d131 + m.m113 * d.d113;
    return phi;
}

// multipole.rs:919-1025 (o0_d1 .. o4_d4), 1408-1528 (generic incl. order 5).
// Order p uses moments through order p-1 only.
void accel_multipole(const Moment& m, const Deriv& d, int order, double out[3]) {
    const int o = std::min(order, 5);
    double ax = -m.m000 * d.d100;
    double ay = -m.m000 * d.d010;
    double az = -m.m000 * d.d001;
    if (o >= 2) {
        ax -= m.m100 * d.d200 + m.m010 * d.d110 + m.m001 * d.d101;
        ay -= m.m100 * d.d110 + m.m010 * d.d020 + m.m001 * d.d011;
        az -= m.m100 * d.d101 + m.m010 * d.d011 + m.m001 * d.d002;
    }
    if (o >= 3) {
        ax -= m.m200 * d.d300 + m.m020 * d.d120 + m.m002 * d.d102;
        ax -= m.m110 * d.d210 + m.m101 * d.d201 + m.m011 * d.d111;
        ay -= m.m200 * d.d210 + m.m020 * d.d030 + m.m002 * d.d012;
        ay -= m.m110 * d.d120 + m.m101 * d.d111 + m.m011 * d.d021;
        az -= m.m200 * d.d201 + m.m020 * d.d021 + m.m002 * d.d003;
        az -= m.m110 * d.d111 + m.m101 * d.d102 + m.m011 * d.d012;
    }
    if (o >= 4) {
        ax -= m.m003 * d.d103 + m.m012 * d.d112 + m.m021 * d.d121 + m.m030 * d.d130 +
              m.m102 * d.d202 + m.m111 * d.d211 + m.m120 * d.d220 + m.m201 * d.d301 +
              m.m210 * d.d310 + m.m300 * d.d400;
        ay -= m.m003 * d.d013 + m.m012 * d.d022 + m.m021 * d.d031 + m.m030 * d.d040 +
              m.m102 * d.d112 + m.m111 * d.d121 + m.m120 * d.d130 + m.m201 * d.d211 +
              m.m210 * d.d220 + m.m300 * d.d310;
        az -= m.m003 * d.d004 + m.m012 * d.d013 + m.m021 * d.d022 + m.m030 * d.d031 +
              m.m102 * d.d103 + m.m111 * d.d112 + m.m120 * d.d121 + m.m201 * d.d202 +
              m.m210 * d.d211 + m.m300 * d.d301;
    }
    if (o >= 5) {
        ax -= m.m004 * d.d104 + m.m013 * d.d113 + m.m022 * d.d122 + m.m031 * d.d131 +
              m.m040 * d.d140 + m.m103 * d.d203 + m.m112 * d.d212 + m.m121 * d.d221 +
              m.m130 * d.d230 + m.m202 * d.d302 + m.m211 * d.d311 + m.m220 * d.d320 +
              m.m301 * d.d401 + m.m310 * d.d410 + m.m400 * d.d500;
        ay -= m.m004 * d.d014 + m.m013 * d.d023 + m.m022 * d.d032 + m.m031 * d.d041 +
              m.m040 * d.d050 + m.m103 * d.d113 + m.m112 * d.d122 + m.m121 * d.d131 +
              m.m130 * d.d140 + m.m202 * d.d212 + m.m211 * d.d221 + m.m220 * d.d230 +
              m.m301 * d.d311 + m.m310 * d.d320 + m.m400 * d.d410;
        az -= m.m004 * d.d005 + m.m013 * d.d014 + m.m022 * d.d023 + m.m031 * d.d032 +
              m.m040 * d.d041 + m.m103 * d.d104 + m.m112 * d.d113 + m.m121 * d.d122 +
              m.m130 * d.d131 + m.m202 * d.d203 + m.m211 * d.d212 + m.m220 * d.d221 +
              m.m301 * d.d302 + m.m310 * d.d311 + m.m400 * d.d401;
    }
    out[0] = ax;
    out[1] = ay;
    out[2] = az;
}

// =============================================================== tree.rs
struct Node {  // tree.rs:572-583
    double center[3];
    double half_size;
    double size2;
    bool has_children = false;
    int64_t children[8];
    std::vector<int64_t> indices;
    int32_t depth = 0;       // instrumentation (not in the reference)
    uint64_t path_hi = 0;    // octant digits of levels 1..21, level 1 most significant
    uint64_t path_lo = 0;    // levels 22..42
};

struct Counters {  // instrumentation for the work model
    int64_t visits = 0, accepts = 0, leaf_visits = 0, leaf_particles = 0;
};

struct Octree {  // tree.rs:591-612
    std::vector<double> positions;  // 3n
    bool has_masses = false;
    std::vector<double> masses;
    bool has_softenings = false;
    std::vector<double> softenings;
    std::vector<Node> nodes;
    std::vector<int64_t> first_subnode, next_branch;
    bool has_bh = false;
    std::vector<double> bh_mass, bh_com;  // NodeBh, tree.rs:585-589
    bool has_multipoles = false;
    std::vector<Moment> multipoles;  // compacted per stored order
    bool has_hmax = false;
    std::vector<double> hmax;
    int multipole_order = 0;
    int64_t leaf_capacity = 1;
    int kernel = PLUMMER;
    int64_t n() const { return (int64_t)positions.size() / 3; }
};

// tree.rs:628-654
void bbox_of_points(const std::vector<double>& pts, double center[3], double& half) {
    double minp[3] = {INFINITY, INFINITY, INFINITY};
    double maxp[3] = {-INFINITY, -INFINITY, -INFINITY};
    int64_t n = (int64_t)pts.size() / 3;
    for (int64_t p = 0; p < n; ++p)
        for (int i = 0; i < 3; ++i) {
            double v = pts[3 * p + i];
            if (v < minp[i]) minp[i] = v;
            if (v > maxp[i]) maxp[i] = v;
        }
    for (int i = 0; i < 3; ++i) center[i] = (minp[i] + maxp[i]) / 2.0;
    half = 0.0;
    for (int i = 0; i < 3; ++i) half = rmax(half, (maxp[i] - minp[i]) / 2.0);
    if (half == 0.0) half = 1e-6;
}

// tree.rs:792-802
Node make_node(const double center[3], double half_size, std::vector<int64_t>&& indices) {
    Node nd;
    double s = half_size * 2.0;
    std::memcpy(nd.center, center, sizeof(nd.center));
    nd.half_size = half_size;
    nd.size2 = s * s;
    nd.indices = std::move(indices);
    for (auto& c : nd.children) c = NONE;
    return nd;
}

// tree.rs:804-845
void subdivide_node(Octree& t, int64_t node_idx) {
    double center[3];
    std::memcpy(center, t.nodes[node_idx].center, sizeof(center));
    double half = t.nodes[node_idx].half_size;
    int32_t depth = t.nodes[node_idx].depth;
    uint64_t phi = t.nodes[node_idx].path_hi, plo = t.nodes[node_idx].path_lo;
    std::vector<int64_t> parent_indices = std::move(t.nodes[node_idx].indices);
    t.nodes[node_idx].indices.clear();
    int64_t child_indices[8];
    for (auto& c : child_indices) c = NONE;
    std::vector<int64_t> buckets[8];
    for (int64_t pi : parent_indices) {
        const double* p = &t.positions[3 * pi];
        int oct = 0;
        if (p[0] >= center[0]) oct |= 1;
        if (p[1] >= center[1]) oct |= 2;
        if (p[2] >= center[2]) oct |= 4;
        buckets[oct].push_back(pi);
    }
    for (int oct = 0; oct < 8; ++oct) {
        if (buckets[oct].empty()) continue;
        double child_center[3] = {center[0], center[1], center[2]};
        double offset = half / 2.0;
        child_center[0] += (oct & 1) ? offset : -offset;
        child_center[1] += (oct & 2) ? offset : -offset;
        child_center[2] += (oct & 4) ? offset : -offset;
        Node child = make_node(child_center, offset, std::move(buckets[oct]));
        child.depth = depth + 1;
        child.path_hi = phi;
        child.path_lo = plo;
        if (child.depth <= 21) child.path_hi |= (uint64_t)oct << (3 * (21 - child.depth));
        else if (child.depth <= 42) child.path_lo |= (uint64_t)oct << (3 * (42 - child.depth));
        int64_t idx = (int64_t)t.nodes.size();
        t.nodes.push_back(std::move(child));
        child_indices[oct] = idx;
    }
    t.nodes[node_idx].has_children = true;
    std::memcpy(t.nodes[node_idx].children, child_indices, sizeof(child_indices));
}

// tree.rs:847-864
void build_recursive(Octree& t, int64_t node_idx) {
    bool should_subdivide = (int64_t)t.nodes[node_idx].indices.size() > t.leaf_capacity;
    if (!should_subdivide) return;
    subdivide_node(t, node_idx);
    int64_t children[8];
    std::memcpy(children, t.nodes[node_idx].children, sizeof(children));
    for (int64_t c : children) {
        if (c == NONE) continue;
        build_recursive(t, c);
    }
}

// tree.rs:736-776
void links_rec(const std::vector<Node>& nodes, std::vector<int64_t>& first,
               std::vector<int64_t>& next, int64_t node_idx) {
    if (!nodes[node_idx].has_children) return;
    const int64_t* children = nodes[node_idx].children;
    int64_t last = NONE;
    for (int k = 0; k < 8; ++k) {
        int64_t c = children[k];
        if (c == NONE) continue;
        if (first[node_idx] == NONE) first[node_idx] = c;
        if (last != NONE) next[last] = c;
        last = c;
    }
    if (last != NONE) next[last] = next[node_idx];
    for (int k = 0; k < 8; ++k) {
        int64_t c = children[k];
        if (c == NONE) continue;
        if (nodes[c].has_children) links_rec(nodes, first, next, c);
    }
}
void build_treewalk_links(Octree& t) {
    size_t n = t.nodes.size();
    t.first_subnode.assign(n, NONE);
    t.next_branch.assign(n, NONE);
    t.next_branch[0] = NONE;
    links_rec(t.nodes, t.first_subnode, t.next_branch, 0);
}

// tree.rs:658-734
Octree* octree_from_owned(std::vector<double>&& positions, bool has_m, std::vector<double>&& masses,
                          bool has_h, std::vector<double>&& softenings, int64_t leaf_capacity,
                          int multipole_order, int kernel) {
    Octree* t = new Octree();
    double center[3], half;
    bbox_of_points(positions, center, half);
    t->positions = std::move(positions);
    t->has_masses = has_m;
    t->masses = std::move(masses);
    t->has_softenings = has_h;
    t->softenings = std::move(softenings);
    t->multipole_order = multipole_order;
    t->leaf_capacity = std::max<int64_t>(leaf_capacity, 1);
    t->kernel = kernel;
    int64_t n = t->n();
    std::vector<int64_t> indices(n);
    for (int64_t i = 0; i < n; ++i) indices[i] = i;
    t->nodes.push_back(make_node(center, half, std::move(indices)));
    build_recursive(*t, 0);
    build_treewalk_links(*t);
    return t;
}

// tree.rs:866-932
void build_bh_payload(Octree& t) {
    size_t nn = t.nodes.size();
    t.bh_mass.assign(nn, 0.0);
    t.bh_com.assign(3 * nn, 0.0);
    const double* masses = t.has_masses ? t.masses.data() : nullptr;
    for (int64_t idx = (int64_t)nn - 1; idx >= 0; --idx) {
        double mass = 0.0;
        double com[3] = {0.0, 0.0, 0.0};
        const Node& node = t.nodes[idx];
        if (!node.has_children) {
            if (!node.indices.empty()) {
                if (masses) {
                    for (int64_t pi : node.indices) {
                        const double* p = &t.positions[3 * pi];
                        double m = masses[pi];
                        mass += m;
                        com[0] += p[0] * m;
                        com[1] += p[1] * m;
                        com[2] += p[2] * m;
                    }
                } else {
                    for (int64_t pi : node.indices) {
                        const double* p = &t.positions[3 * pi];
                        mass += 1.0;
                        com[0] += p[0];
                        com[1] += p[1];
                        com[2] += p[2];
                    }
                }
                if (mass > 0.0) {
                    com[0] /= mass;
                    com[1] /= mass;
                    com[2] /= mass;
                }
            }
        } else {
            for (int k = 0; k < 8; ++k) {
                int64_t c = node.children[k];
                if (c == NONE) continue;
                double cm = t.bh_mass[c];
                if (cm == 0.0) continue;
                mass += cm;
                com[0] += t.bh_com[3 * c + 0] * cm;
                com[1] += t.bh_com[3 * c + 1] * cm;
                com[2] += t.bh_com[3 * c + 2] * cm;
            }
            if (mass > 0.0) {
                com[0] /= mass;
                com[1] /= mass;
                com[2] /= mass;
            }
        }
        t.bh_mass[idx] = mass;
        t.bh_com[3 * idx + 0] = com[0];
        t.bh_com[3 * idx + 1] = com[1];
        t.bh_com[3 * idx + 2] = com[2];
    }
    t.has_bh = true;
}

// tree.rs:941-965
void build_hmax_payload(Octree& t) {
    if (!t.has_softenings) {
        t.has_hmax = false;
        t.hmax.clear();
        return;
    }
    size_t nn = t.nodes.size();
    t.hmax.assign(nn, 0.0);
    for (int64_t idx = (int64_t)nn - 1; idx >= 0; --idx) {
        const Node& node = t.nodes[idx];
        double m = 0.0;
        if (!node.has_children) {
            for (int64_t pi : node.indices) m = rmax(m, rmax(t.softenings[pi], MIN_SOFTENING));
        } else {
            for (int k = 0; k < 8; ++k) {
                int64_t c = node.children[k];
                if (c == NONE) continue;
                m = rmax(m, t.hmax[c]);
            }
        }
        t.hmax[idx] = m;
    }
    t.has_hmax = true;
}

// tree.rs:1014-1067
void build_multipole_payload(Octree& t) {
    const double* masses = t.has_masses ? t.masses.data() : nullptr;
    int order = std::min(t.multipole_order, 5);
    size_t nn = t.nodes.size();
    std::vector<Moment> moments(nn);
    for (int64_t idx = (int64_t)nn - 1; idx >= 0; --idx) {
        const Node& node = t.nodes[idx];
        if (t.bh_mass[idx] == 0.0) continue;
        const double* center = &t.bh_com[3 * idx];
        if (!node.has_children) {
            if (node.indices.empty()) continue;
            moments[idx] = moment_from_points(t.positions.data(), masses, node.indices.data(),
                                              (int64_t)node.indices.size(), center, order);
        } else {
            Moment acc;
            for (int k = 0; k < 8; ++k) {
                int64_t c = node.children[k];
                if (c == NONE) continue;
                if (t.bh_mass[c] == 0.0) continue;
                double shift[3] = {center[0] - t.bh_com[3 * c + 0], center[1] - t.bh_com[3 * c + 1],
                                   center[2] - t.bh_com[3 * c + 2]};
                Moment translated = translate_multipole(moments[c], shift, order);
                moment_add_assign(acc, translated);
            }
            moments[idx] = acc;
        }
    }
    for (auto& m : moments) compact_moment(m, order);  // MultipoleMoments::from_full
    t.multipoles = std::move(moments);
    t.has_multipoles = true;
}

// tree.rs:968-1012
void build_mass_payload(Octree& t) {
    build_bh_payload(t);
    build_hmax_payload(t);
    if (t.multipole_order > 0) build_multipole_payload(t);
}

struct Ctx {  // TraversalCtx, tree.rs:614-625
    const Octree* t;
    const double* masses;      // nullable
    const double* softenings;  // nullable (tree.softenings, may differ from what hmax was built on)
    const double* hmax;        // nullable
    double theta2;
    double multipole_eps2;
    int kernel;
};

// tree.rs:55-71
inline bool node_soft_ok(int64_t idx, double dist2, bool has_target_h, double target_h,
                         const Ctx& ctx) {
    if (!ctx.hmax) return true;
    double h = rmax(ctx.hmax[idx], MIN_SOFTENING);
    if (has_target_h) h = rmax(h, rmax(target_h, MIN_SOFTENING));
    if (h <= 0.0) return true;
    double c = multipole_min_separation_factor(ctx.kernel);
    double ch = c * h;
    return dist2 > ch * ch;
}

inline double inv_r_from_r2(double r2) {  // tree.rs:39-43
    double s2 = r2 + R2_TINY;
    return 1.0 / std::sqrt(s2);
}
inline double inv_r3_from_r2(double r2) {  // tree.rs:45-53
    double s2 = r2 + R2_TINY;
    double inv_r = 1.0 / std::sqrt(s2);
    double inv_r2 = inv_r * inv_r;
    return inv_r2 * inv_r;
}

// tree.rs:97-277. The "constant target softening" fast path (:122-171) is restated too:
// it is unreachable from the four entry points but harmless.
void leaf_potential_sum(const Octree& t, const std::vector<int64_t>& indices, const Ctx& ctx,
                        const double target[3], int64_t skip, bool has_target_h, double target_h_in,
                        double& out) {
    const double* positions = t.positions.data();
    const double* masses = ctx.masses;
    const double* hs = ctx.softenings;
    double tx = target[0], ty = target[1], tz = target[2];
    double target_h = rmax(has_target_h ? target_h_in : MIN_SOFTENING, MIN_SOFTENING);
    bool use_softening = hs != nullptr || target_h > 0.0;
    int kernel = ctx.kernel;

    if (use_softening && masses && !hs) {
        double h = target_h;
        if (h <= 0.0) {
            // fall through
        } else if (kernel == SPLINE) {
            double hh = h * h;
            for (int64_t pi : indices) {
                if (pi == skip) continue;
                const double* p = positions + 3 * pi;
                double ddx = p[0] - tx, ddy = p[1] - ty, ddz = p[2] - tz;
                double r2 = std::fma(ddx, ddx, std::fma(ddy, ddy, ddz * ddz));
                double m = masses[pi];
                if (r2 >= hh) {
                    out += -m * inv_r_from_r2(r2);
                } else {
                    double r = std::sqrt(r2 + R2_TINY);
                    out += m * kernel_potential_per_unit_mass(kernel, r, h);
                }
            }
            return;
        } else {
            for (int64_t pi : indices) {
                if (pi == skip) continue;
                const double* p = positions + 3 * pi;
                double ddx = p[0] - tx, ddy = p[1] - ty, ddz = p[2] - tz;
                double r2 = std::fma(ddx, ddx, std::fma(ddy, ddy, ddz * ddz));
                double r = std::sqrt(r2 + R2_TINY);
                out += masses[pi] * kernel_potential_per_unit_mass(kernel, r, h);
            }
            return;
        }
    }

    if (!use_softening) {
        for (int64_t pi : indices) {
            if (pi == skip) continue;
            const double* p = positions + 3 * pi;
            double ddx = p[0] - tx, ddy = p[1] - ty, ddz = p[2] - tz;
            double r2 = std::fma(ddx, ddx, std::fma(ddy, ddy, ddz * ddz));
            double inv_r = inv_r_from_r2(r2);
            if (masses) out += -masses[pi] * inv_r;
            else out += -inv_r;
        }
        return;
    }

    bool kernel_is_spline = kernel == SPLINE;
    for (int64_t pi : indices) {
        if (pi == skip) continue;
        const double* p = positions + 3 * pi;
        double ddx = p[0] - tx, ddy = p[1] - ty, ddz = p[2] - tz;
        double r2 = std::fma(ddx, ddx, std::fma(ddy, ddy, ddz * ddz));
        double m = masses ? masses[pi] : 1.0;
        double h;
        if (hs) {
            double hi = rmax(hs[pi], MIN_SOFTENING);
            h = rmax(hi, target_h);
        } else {
            h = target_h;
        }
        if (h <= 0.0 || (kernel_is_spline && r2 >= h * h)) {
            out += -m * inv_r_from_r2(r2);
        } else {
            double r = std::sqrt(r2 + R2_TINY);
            out += m * kernel_potential_per_unit_mass(kernel, r, h);
        }
    }
}

// tree.rs:279-417
void leaf_acceleration_sum(const Octree& t, const std::vector<int64_t>& indices, const Ctx& ctx,
                           const double target[3], int64_t skip, bool has_target_h,
                           double target_h_in, double out[3]) {
    const double* positions = t.positions.data();
    const double* masses = ctx.masses;
    const double* hs = ctx.softenings;
    double tx = target[0], ty = target[1], tz = target[2];
    double target_h = rmax(has_target_h ? target_h_in : MIN_SOFTENING, MIN_SOFTENING);
    bool use_softening = hs != nullptr || target_h > 0.0;
    int kernel = ctx.kernel;

    if (!use_softening) {
        for (int64_t pi : indices) {
            if (pi == skip) continue;
            const double* p = positions + 3 * pi;
            double ddx = p[0] - tx, ddy = p[1] - ty, ddz = p[2] - tz;
            double r2 = std::fma(ddx, ddx, std::fma(ddy, ddy, ddz * ddz));
            double inv_r3 = inv_r3_from_r2(r2);
            if (masses) {
                double m = masses[pi];
                out[0] += m * ddx * inv_r3;
                out[1] += m * ddy * inv_r3;
                out[2] += m * ddz * inv_r3;
            } else {
                out[0] += ddx * inv_r3;
                out[1] += ddy * inv_r3;
                out[2] += ddz * inv_r3;
            }
        }
        return;
    }

    bool kernel_is_spline = kernel == SPLINE;
    for (int64_t pi : indices) {
        if (pi == skip) continue;
        const double* p = positions + 3 * pi;
        double ddx = p[0] - tx, ddy = p[1] - ty, ddz = p[2] - tz;
        double r2 = std::fma(ddx, ddx, std::fma(ddy, ddy, ddz * ddz));
        double m = masses ? masses[pi] : 1.0;
        double h;
        if (hs) {
            double hi = rmax(hs[pi], MIN_SOFTENING);
            h = rmax(hi, target_h);
        } else {
            h = target_h;
        }
        if (h <= 0.0 || (kernel_is_spline && r2 >= h * h)) {
            double inv_r3 = inv_r3_from_r2(r2);
            out[0] += m * ddx * inv_r3;
            out[1] += m * ddy * inv_r3;
            out[2] += m * ddz * inv_r3;
        } else {
            double r = std::sqrt(r2 + R2_TINY);
            double g = kernel_accel_factor(kernel, r, h);
            out[0] += m * ddx * g;
            out[1] += m * ddy * g;
            out[2] += m * ddz * g;
        }
    }
}

// tree.rs:1069-1206 (potential; no-multipole and with-multipole bodies share control flow)
void potential_traversal(const Octree& t, const Ctx& ctx, const double target[3], int64_t skip,
                         bool has_target_h, double target_h, double& out, Counters* cnt) {
    double tx = target[0], ty = target[1], tz = target[2];
    bool softening_enabled = ctx.hmax != nullptr || has_target_h;
    int64_t idx = 0;
    while (idx != NONE) {
        if (cnt) cnt->visits++;
        double nmass = t.bh_mass[idx];
        if (nmass == 0.0) {
            idx = t.next_branch[idx];
            continue;
        }
        const Node& node = t.nodes[idx];
        if (!node.has_children) {
            if (cnt) {
                cnt->leaf_visits++;
                cnt->leaf_particles += (int64_t)node.indices.size();
            }
            leaf_potential_sum(t, node.indices, ctx, target, skip, has_target_h, target_h, out);
            idx = t.next_branch[idx];
            continue;
        }
        double dx = t.bh_com[3 * idx + 0] - tx;
        double dy = t.bh_com[3 * idx + 1] - ty;
        double dz = t.bh_com[3 * idx + 2] - tz;
        double dist2 = std::fma(dx, dx, std::fma(dy, dy, dz * dz)) + ctx.multipole_eps2;
        bool soft_ok = softening_enabled ? node_soft_ok(idx, dist2, has_target_h, target_h, ctx) : true;
        if (soft_ok && node.size2 < ctx.theta2 * dist2) {
            if (cnt) cnt->accepts++;
            if (!t.has_multipoles) {
                double inv_r = inv_r_from_r2(dist2);  // tree.rs:1127-1128
                out += -nmass * inv_r;
            } else {
                Deriv d = derivatives(dx, dy, dz, ctx.multipole_eps2, std::max(1, std::min(t.multipole_order, 5)));
                out += potential_multipole(t.multipoles[idx], d, t.multipole_order);
            }
            idx = t.next_branch[idx];
        } else {
            idx = t.first_subnode[idx];
        }
    }
}

// tree.rs:1228-1370
void acceleration_traversal(const Octree& t, const Ctx& ctx, const double target[3], int64_t skip,
                            bool has_target_h, double target_h, double out[3], Counters* cnt) {
    double tx = target[0], ty = target[1], tz = target[2];
    bool softening_enabled = ctx.hmax != nullptr || has_target_h;
    int64_t idx = 0;
    while (idx != NONE) {
        if (cnt) cnt->visits++;
        double nmass = t.bh_mass[idx];
        if (nmass == 0.0) {
            idx = t.next_branch[idx];
            continue;
        }
        const Node& node = t.nodes[idx];
        if (!node.has_children) {
            if (cnt) {
                cnt->leaf_visits++;
                cnt->leaf_particles += (int64_t)node.indices.size();
            }
            leaf_acceleration_sum(t, node.indices, ctx, target, skip, has_target_h, target_h, out);
            idx = t.next_branch[idx];
            continue;
        }
        double dx = t.bh_com[3 * idx + 0] - tx;
        double dy = t.bh_com[3 * idx + 1] - ty;
        double dz = t.bh_com[3 * idx + 2] - tz;
        double dist2 = std::fma(dx, dx, std::fma(dy, dy, dz * dz)) + ctx.multipole_eps2;
        bool soft_ok = softening_enabled ? node_soft_ok(idx, dist2, has_target_h, target_h, ctx) : true;
        if (soft_ok && node.size2 < ctx.theta2 * dist2) {
            if (cnt) cnt->accepts++;
            if (!t.has_multipoles) {
                double inv_r = inv_r_from_r2(dist2);  // tree.rs:1285-1290
                double inv_r2 = inv_r * inv_r;
                double inv_r3 = inv_r2 * inv_r;
                out[0] += nmass * dx * inv_r3;
                out[1] += nmass * dy * inv_r3;
                out[2] += nmass * dz * inv_r3;
            } else {
                Deriv d = derivatives(dx, dy, dz, ctx.multipole_eps2, std::max(1, std::min(t.multipole_order, 5)));
                double a[3];
                accel_multipole(t.multipoles[idx], d, t.multipole_order, a);
                out[0] += a[0];
                out[1] += a[1];
                out[2] += a[2];
            }
            idx = t.next_branch[idx];
        } else {
            idx = t.first_subnode[idx];
        }
    }
}

Ctx make_ctx(const Octree& t, double theta) {  // tree.rs:1423-1431
    Ctx c;
    c.t = &t;
    c.masses = t.has_masses ? t.masses.data() : nullptr;
    c.softenings = t.has_softenings ? t.softenings.data() : nullptr;
    c.hmax = t.has_hmax ? t.hmax.data() : nullptr;
    c.theta2 = theta * theta;
    c.multipole_eps2 = R2_TINY;
    c.kernel = t.kernel;
    return c;
}

thread_local std::string g_err;

}  // namespace

// =============================================================== C-ABI of the oracle
extern "C" {

const char* pnbx_oracle_last_error(void) { return g_err.c_str(); }

int pnbx_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void pnbx_oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// Same argument meaning as pnbx_direct (include/pnbx_gravity.h), host pointers only, and
// self mode always covers all n sources (the reference has no shards).
int pnbx_oracle_direct(const double* src_pos, const double* src_mass, const double* src_h, int64_t n,
                       const double* tgt_pos, int64_t m, int kernel, int want, double* out_pot,
                       double* out_acc) {
    if (kernel < 0 && src_h) {
        g_err = "softenings require an explicit kernel; pass kernel=0/1 (or omit softenings)";
        return 1;
    }
    if (kernel > 1) {
        g_err = "kernel must be 0 (Plummer) or 1 (CubicSplineW2)";
        return 1;
    }
    if (!tgt_pos) {
        if (want & 1) {
            if (kernel < 0) direct_potentials(src_pos, src_mass, n, out_pot);
            else direct_potentials_kernel(src_pos, src_mass, src_h, n, kernel, out_pot);
        }
        if (want & 2) {
            if (kernel < 0) direct_accelerations(src_pos, src_mass, n, out_acc);
            else direct_accelerations_kernel(src_pos, src_mass, src_h, n, kernel, out_acc);
        }
    } else {
        if (want & 1) {
            if (kernel < 0) direct_potentials_at_points(src_pos, src_mass, n, tgt_pos, m, out_pot);
            else direct_potentials_kernel_at_points(src_pos, src_mass, src_h, n, tgt_pos, m, kernel, out_pot);
        }
        if (want & 2) {
            if (kernel < 0) direct_accelerations_at_points(src_pos, src_mass, n, tgt_pos, m, out_acc);
            else direct_accelerations_kernel_at_points(src_pos, src_mass, src_h, n, tgt_pos, m, kernel, out_acc);
        }
    }
    return 0;
}

// scalar kernel functions (known-answer tests)
double pnbx_oracle_kernel_potential(int kind, double r, double h) { return kernel_potential_per_unit_mass(kind, r, h); }
double pnbx_oracle_kernel_accel_factor(int kind, double r, double h) { return kernel_accel_factor(kind, r, h); }

// Octree::new (gravity.rs:121-226)
void* pnbx_oracle_tree_create(const double* pos, const double* mass, const double* h, int64_t n,
                              int64_t leaf_capacity, int multipole_order, int kernel) {
    std::vector<double> p(pos, pos + 3 * n);
    std::vector<double> mv, hv;
    if (mass) mv.assign(mass, mass + n);
    if (h) hv.assign(h, h + n);
    Octree* t = octree_from_owned(std::move(p), mass != nullptr, std::move(mv), h != nullptr,
                                  std::move(hv), leaf_capacity, multipole_order, kernel);
    if (t->has_masses) build_mass_payload(*t);
    return t;
}
// Octree.build_mass (gravity.rs:228-239)
int pnbx_oracle_tree_build_mass(void* tp, const double* mass) {
    Octree& t = *static_cast<Octree*>(tp);
    if (mass) {
        t.masses.assign(mass, mass + t.n());
        t.has_masses = true;
    }
    build_mass_payload(t);
    return 0;
}
int pnbx_oracle_tree_set_softenings(void* tp, const double* h) {  // tree.rs:777-782
    Octree& t = *static_cast<Octree*>(tp);
    if (h) {
        t.softenings.assign(h, h + t.n());
        t.has_softenings = true;
    } else {
        t.softenings.clear();
        t.has_softenings = false;
    }
    return 0;
}
int pnbx_oracle_tree_set_kernel(void* tp, int kernel) {  // tree.rs:784-786
    static_cast<Octree*>(tp)->kernel = kernel;
    return 0;
}
void pnbx_oracle_tree_destroy(void* tp) { delete static_cast<Octree*>(tp); }

// tree.rs:1415-1558. counters (nullable): [visits, accepts, leaf_visits, leaf_particles] totals.
// Self mode evaluates own particles [begin, begin+m) (compute_* restricted to a target range: the per-target work of
// tree.rs:1415-1496 is independent of the other targets; begin = 0, m = N is the reference call).
int pnbx_oracle_tree_eval_range(void* tp, const double* tgt_pos, int64_t begin, int64_t m, double theta, int want,
                                double* out_pot, double* out_acc, int64_t* counters);
int pnbx_oracle_tree_eval(void* tp, const double* tgt_pos, int64_t m, double theta, int want,
                          double* out_pot, double* out_acc, int64_t* counters) {
    const Octree& t = *static_cast<Octree*>(tp);
    return pnbx_oracle_tree_eval_range(tp, tgt_pos, 0, tgt_pos ? m : t.n(), theta, want, out_pot, out_acc, counters);
}
int pnbx_oracle_tree_eval_range(void* tp, const double* tgt_pos, int64_t begin, int64_t m, double theta, int want,
                                double* out_pot, double* out_acc, int64_t* counters) {
    const Octree& t = *static_cast<Octree*>(tp);
    if (!t.has_bh) {
        g_err = "mass payload not built; call build_mass() before compute";
        return 3;
    }
    Ctx ctx = make_ctx(t, theta);
    const bool self = tgt_pos == nullptr;
    if (self && (begin < 0 || begin + m > t.n())) {
        g_err = "target range outside [0, N)";
        return 1;
    }
    const int64_t count = m;
    int64_t tot[4] = {0, 0, 0, 0};
    for (int pass = 0; pass < 2; ++pass) {
        const bool do_pot = pass == 0;
        if (do_pot && !(want & 1)) continue;
        if (!do_pot && !(want & 2)) continue;
        int64_t v = 0, a = 0, lv = 0, lp = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : v, a, lv, lp) if (count >= 1024)
        for (int64_t i = 0; i < count; ++i) {
            const int64_t gi = self ? begin + i : i;  // particle index of a self target
            const double* target = self ? &t.positions[3 * gi] : tgt_pos + 3 * i;
            int64_t skip = self ? gi : NONE;
            bool has_th = self && t.has_softenings;
            double th = has_th ? t.softenings[gi] : 0.0;
            Counters c;
            if (do_pot) {
                double o = 0.0;
                potential_traversal(t, ctx, target, skip, has_th, th, o, counters ? &c : nullptr);
                out_pot[i] = o;
            } else {
                double o[3] = {0.0, 0.0, 0.0};
                acceleration_traversal(t, ctx, target, skip, has_th, th, o, counters ? &c : nullptr);
                out_acc[3 * i + 0] = o[0];
                out_acc[3 * i + 1] = o[1];
                out_acc[3 * i + 2] = o[2];
            }
            v += c.visits;
            a += c.accepts;
            lv += c.leaf_visits;
            lp += c.leaf_particles;
        }
        tot[0] += v;
        tot[1] += a;
        tot[2] += lv;
        tot[3] += lp;
    }
    if (counters) std::memcpy(counters, tot, sizeof(tot));
    return 0;
}

// info: [n_particles, n_nodes, n_leaves, depth, stored_coeffs, has_payload, has_hmax]
int pnbx_oracle_tree_info(void* tp, int64_t* info) {
    const Octree& t = *static_cast<Octree*>(tp);
    int64_t leaves = 0, depth = 0;
    for (const Node& nd : t.nodes) {
        if (!nd.has_children) leaves++;
        depth = std::max<int64_t>(depth, nd.depth);
    }
    info[0] = t.n();
    info[1] = (int64_t)t.nodes.size();
    info[2] = leaves;
    info[3] = depth;
    info[4] = t.has_multipoles ? stored_coeffs(t.multipole_order) : (t.has_bh ? 1 : 0);
    info[5] = t.has_bh;
    info[6] = t.has_hmax;
    return 0;
}

// Same fields as pnbx_tree_dump_topology (include/pnbx_gravity.h).
int pnbx_oracle_tree_dump_topology(void* tp, double* center, double* half, int32_t* depth,
                                   int64_t* first_subnode, int64_t* next_branch, int64_t* leaf_start,
                                   int64_t* leaf_count, int64_t* leaf_particles, uint64_t* path_hi,
                                   uint64_t* path_lo) {
    const Octree& t = *static_cast<Octree*>(tp);
    int64_t cursor = 0;
    for (size_t i = 0; i < t.nodes.size(); ++i) {
        const Node& nd = t.nodes[i];
        if (center) std::memcpy(center + 3 * i, nd.center, 3 * sizeof(double));
        if (half) half[i] = nd.half_size;
        if (depth) depth[i] = nd.depth;
        if (first_subnode) first_subnode[i] = t.first_subnode[i];
        if (next_branch) next_branch[i] = t.next_branch[i];
        if (path_hi) path_hi[i] = nd.path_hi;
        if (path_lo) path_lo[i] = nd.path_lo;
        if (!nd.has_children) {
            if (leaf_start) leaf_start[i] = cursor;
            if (leaf_count) leaf_count[i] = (int64_t)nd.indices.size();
            if (leaf_particles)
                for (size_t k = 0; k < nd.indices.size(); ++k) leaf_particles[cursor + k] = nd.indices[k];
            cursor += (int64_t)nd.indices.size();
        } else {
            if (leaf_start) leaf_start[i] = -1;
            if (leaf_count) leaf_count[i] = -1;
        }
    }
    return 0;
}

// moments: stored_coeffs per node in field order (zeros if no multipole payload beyond m000).
int pnbx_oracle_tree_dump_payload(void* tp, double* mass, double* com, double* hmax, double* moments) {
    const Octree& t = *static_cast<Octree*>(tp);
    if (!t.has_bh) {
        g_err = "mass payload not built";
        return 3;
    }
    size_t nn = t.nodes.size();
    if (mass) std::memcpy(mass, t.bh_mass.data(), nn * sizeof(double));
    if (com) std::memcpy(com, t.bh_com.data(), 3 * nn * sizeof(double));
    if (hmax && t.has_hmax) std::memcpy(hmax, t.hmax.data(), nn * sizeof(double));
    if (moments && t.has_multipoles) {
        int k = stored_coeffs(t.multipole_order);
        for (size_t i = 0; i < nn; ++i)
            std::memcpy(moments + (size_t)k * i, &t.multipoles[i], (size_t)k * sizeof(double));
    }
    return 0;
}

// Stand-alone multipole entry points for the restated single_node / translate tests.
// moments: 56 doubles in field order.
void pnbx_oracle_p2m(const double* pos, const double* mass, int64_t n, const double* center, int order,
                     double* moments) {
    std::vector<int64_t> idx(n);
    for (int64_t i = 0; i < n; ++i) idx[i] = i;
    Moment m = moment_from_points(pos, mass, idx.data(), n, center, order);
    std::memcpy(moments, &m, sizeof(m));
}
void pnbx_oracle_m2m(const double* child, const double* shift, int order, double* out) {
    Moment c;
    std::memcpy(&c, child, sizeof(c));
    Moment o = translate_multipole(c, shift, order);
    std::memcpy(out, &o, sizeof(o));
}
// gravity_potential_multipole / gravity_accel_multipole with PotentialDerivatives::new(.., 5)
void pnbx_oracle_m2p(const double* moments, const double* dxyz, double eps2, int order, double* pot,
                     double* acc) {
    Moment m;
    std::memcpy(&m, moments, sizeof(m));
    Deriv d = derivatives(dxyz[0], dxyz[1], dxyz[2], eps2, 5);
    if (pot) *pot = potential_multipole(m, d, order);
    if (acc) accel_multipole(m, d, order, acc);
}

}  // extern "C"
