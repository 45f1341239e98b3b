"""CPU check of the closed forms the fp32 GPU kernels evaluate instead of the reference's literal expressions
(float64 numpy here, so only the ALGEBRA is on trial, not fp32 rounding):

  * csrc/spline.cuh `w2_terms_f32`: branch-free W2 / W2' with 1/u = h/r  vs  kernel.rs:41-128 (oracle)
  * csrc/multipole.cuh `m2p_fast` (orders 2, 3): traceless quadrupole T = 3(S - trS/3 I) and folded cubic
    C'(u) = 15 C(u) - 3 (w.u)(u.u), as packed by tree_build.cu `pack_walk_moments`  vs  the reference's
    derivative-tensor M2P (multipole.rs:591-1025, oracle `m2p`)
  * squared-softening rule of the fp32 leaf / direct loops: max(h_s, h_t)^2 == max(h_s^2, h_t^2) for clamped h.
"""
import numpy as np

from oracle import oracle as O

# field order of the first 20 coefficients (multipole.rs:11-74)
I000, I100, I010, I001, I200, I020, I002, I110, I101, I011 = range(10)
I300, I030, I003, I210, I201, I120, I102, I021, I012, I111 = range(10, 20)


def w2_terms(r2, rinv, h, hinv):
    u, uinv = r2 * rinv * hinv, h * rinv
    u2 = u * u
    wi = u2 * (u2 * (6.4 * u - 9.6) + 16.0 / 3.0) - 2.8
    wo = u2 * (u * (u * (-32.0 / 15.0 * u + 9.6) - 16.0) + 32.0 / 3.0) + (uinv / 15.0 - 3.2)
    kpot = np.where(u < 0.5, wi, wo) * hinv
    pi = u * (u2 * (32.0 * u - 38.4) + 32.0 / 3.0)
    po = u * (u * (u * (-32.0 / 3.0 * u + 38.4) - 48.0) + 64.0 / 3.0) - uinv * uinv / 15.0
    kacc = np.where(u < 0.5, pi, po) * hinv * hinv * rinv
    return kpot, kacc


def test_branch_free_w2_matches_reference_kernel():
    rng = np.random.default_rng(1)
    h = rng.uniform(1e-3, 2.0, 4000)
    r = h * np.concatenate([rng.uniform(1e-6, 1.0, 3990), [0.5, 0.5 - 1e-12, 0.5 + 1e-12, 1e-9, 0.999999, 0.25, 0.75, 0.1, 0.9, 0.49]])
    r2 = r * r
    kpot, kacc = w2_terms(r2, 1.0 / np.sqrt(r2), h, 1.0 / h)
    ref_p = np.array([O.kernel_potential(1, ri, hi) for ri, hi in zip(r, h)])
    ref_a = np.array([O.kernel_accel_factor(1, ri, hi) for ri, hi in zip(r, h)])
    assert np.max(np.abs(kpot - ref_p) / np.abs(ref_p)) < 1e-12
    # W2'(u) ~ 32/3 u near 0: compare on the scale of the leading term
    assert np.max(np.abs(kacc - ref_a) / np.maximum(np.abs(ref_a), 1e-300)) < 1e-9


def pack_fast_record(m, order):
    tr = m[I200] + m[I020] + m[I002]
    rec = np.zeros(20)
    rec[0] = m[I000]
    rec[1:7] = [3 * m[I200] - tr, 3 * m[I020] - tr, 3 * m[I002] - tr, 1.5 * m[I110], 1.5 * m[I101], 1.5 * m[I011]]
    if order == 3:
        wx = 3 * m[I300] + m[I120] + m[I102]
        wy = 3 * m[I030] + m[I210] + m[I012]
        wz = 3 * m[I003] + m[I201] + m[I021]
        rec[8:18] = [15 * m[I300] - 3 * wx, 15 * m[I030] - 3 * wy, 15 * m[I003] - 3 * wz, 15 * m[I210] - 3 * wy,
                     15 * m[I201] - 3 * wz, 15 * m[I120] - 3 * wx, 15 * m[I102] - 3 * wx, 15 * m[I021] - 3 * wz,
                     15 * m[I012] - 3 * wy, 15 * m[I111]]
    return rec


def m2p_fast(rec, d, order):
    r2 = d @ d
    ri = 1.0 / np.sqrt(r2)
    ri2, ri3 = ri * ri, ri * ri * ri
    u = d * ri
    M = rec[0]
    Txx, Tyy, Tzz, Txy, Txz, Tyz = rec[1:7]
    q = np.array([Txx * u[0] + Txy * u[1] + Txz * u[2], Txy * u[0] + Tyy * u[1] + Tyz * u[2],
                  Txz * u[0] + Tyz * u[1] + Tzz * u[2]])
    st = q @ u
    phi = -M * ri - ri3 * st
    acc = M * ri3 * d
    if order >= 3:
        c300, c030, c003, c210, c201, c120, c102, c021, c012, c111 = rec[8:18]
        cx = c300 * u[0] + c210 * u[1] + c201 * u[2]
        cy = c030 * u[1] + c120 * u[0] + c021 * u[2]
        cz = c003 * u[2] + c102 * u[0] + c012 * u[1]
        C = u[0] ** 2 * cx + u[1] ** 2 * cy + u[2] ** 2 * cz + c111 * u[0] * u[1] * u[2]
        phi += ri2 * ri2 * C
        acc = acc + 2.0 * ri2 * ri2 * (2.5 * st * u - q)
    return phi, acc


def test_folded_moment_records_match_reference_m2p():
    rng = np.random.default_rng(2)
    for order in (2, 3):
        for _ in range(50):
            # moments of a random clump about its centre of mass (dipole ~ 0, as in the tree)
            n = 40
            pos = rng.normal(0.0, 0.3, (n, 3))
            mass = rng.uniform(0.5, 1.5, n)
            com = (mass[:, None] * pos).sum(0) / mass.sum()
            mom = np.asarray(O.p2m(pos, mass, com, order), dtype=float)
            d = rng.normal(0.0, 1.0, 3)
            d *= rng.uniform(2.0, 6.0) / np.linalg.norm(d)     # source COM - target, outside the clump
            full = np.zeros(56)
            full[:len(mom)] = mom
            phi_ref, acc_ref = O.m2p(mom, d, order)
            phi, acc = m2p_fast(pack_fast_record(full, order), d, order)
            assert abs(phi - phi_ref) < 1e-12 * abs(phi_ref)
            # the fast form drops the dipole (rounding noise about the COM): compare on the monopole scale
            assert np.linalg.norm(acc - acc_ref) < 1e-11 * np.linalg.norm(acc_ref)


def test_squared_softening_rule():
    rng = np.random.default_rng(3)
    hs = np.maximum(rng.uniform(-0.5, 1.0, 1000), 0.0)
    ht = np.maximum(rng.uniform(-0.5, 1.0, 1000), 0.0)
    assert np.array_equal(np.maximum(hs, ht) ** 2, np.maximum(hs ** 2, ht ** 2))
