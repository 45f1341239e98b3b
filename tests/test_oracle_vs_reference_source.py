"""The oracle against numbers produced by EXECUTING THE REFERENCE'S SOURCE TEXT (tests/golden/reference_exec.npz,
made by tests/golden/make_reference_exec.py: a mechanical Rust->Python translation of kernel.rs / multipole.rs function
bodies run in IEEE binary64). Everything here is BIT-EXACT equality: same formulas, same operation order, same rounding.
This is what pins the reference's quirks independently of the hand-written port: no dipole term in the potential,
accelerations from moments through order p-1 (with their dipole terms), order 0/1 = monopole, W2 branch points, the
P2M factor order and the M2M term order / zero-skipping.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_exec.npz"))

# the oracle's coefficient order must be the reference's struct field order (multipole.rs:11-74)
FIELDS = [str(f) for f in G["field_order"]]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def test_field_order_is_the_reference_struct_order():
    assert len(FIELDS) == 56 and FIELDS[:10] == ["m000", "m100", "m010", "m001", "m200", "m020", "m002", "m110", "m101", "m011"]
    # P2M of a single unit mass at (x,y,z) about the origin puts x^l y^m z^n / (l! m! n!) into field m{l}{m}{n}
    from math import factorial as f
    x, y, z = 0.3, -0.7, 1.1
    m = O.p2m(np.array([[x, y, z]]), np.array([1.0]), np.zeros(3), 5)
    for i, name in enumerate(FIELDS):
        l, mm, n = (int(c) for c in name[1:])
        assert m[i] == pytest.approx(x ** l * y ** mm * z ** n / (f(l) * f(mm) * f(n)), rel=1e-14), name


@pytest.mark.parametrize("kind", [0, 1])
def test_kernel_functions_bit_exact(kind):
    r, h = G["kernel_r"], G["kernel_h"]
    pot = np.array([O.kernel_potential(kind, a, b) for a, b in zip(r, h)])
    acc = np.array([O.kernel_accel_factor(kind, a, b) for a, b in zip(r, h)])
    assert np.array_equal(bits(pot), bits(G[f"kernel_pot_{kind}"]))
    assert np.array_equal(bits(acc), bits(G[f"kernel_acc_{kind}"]))


def test_w2_and_derivative_bit_exact_through_the_spline_kernel():
    # K(r, h=1) = W2(r) and g(r, h=1) = W2'(r) / r  (kernel.rs:46-54, 72-80 with h_inv = 1)
    u = G["w2_u"]
    ok = u > 0
    w2 = np.array([O.kernel_potential(1, a, 1.0) for a in u[ok]])
    assert np.array_equal(bits(w2), bits(G["w2"][ok]))
    g = np.array([O.kernel_accel_factor(1, a, 1.0) for a in u[ok]])
    assert np.array_equal(bits(g), bits(G["w2_prime"][ok] * (1.0 * 1.0) / u[ok]))


@pytest.mark.parametrize("order", range(6))
def test_p2m_bit_exact(order):
    m = O.p2m(G["p2m_pos"], G["p2m_mass"], G["p2m_center"], order)
    assert np.array_equal(bits(m), bits(G[f"p2m_o{order}"]))


def test_p2m_unit_masses_and_index_order():
    pos = G["p2m_pos"][[7, 2, 9]]
    m = O.p2m(pos, None, G["p2m_center"], 5)
    assert np.array_equal(bits(m), bits(G["p2m_unit_idx729"]))


@pytest.mark.parametrize("order", range(6))
def test_m2m_bit_exact(order):
    t = O.m2m(G["p2m_o5"], G["m2m_shift"], order)
    assert np.array_equal(bits(t), bits(G[f"m2m_o{order}"]))


@pytest.mark.parametrize("order", range(6))
def test_m2p_potential_and_acceleration_bit_exact(order):
    # random moments INCLUDING non-zero dipole coefficients: the reference ignores them in the potential and uses them
    # in the acceleration (orders >= 2) — a port that "fixed" either would fail here
    pot, acc = [], []
    for d, m in zip(G["m2p_dxyz"], G["m2p_moments"]):
        p, a = O.m2p(m, d, order)
        pot.append(p)
        acc.append(a)
    assert np.array_equal(bits(np.array(pot)), bits(G[f"m2p_pot_o{order}"]))
    assert np.array_equal(bits(np.array(acc)), bits(G[f"m2p_acc_o{order}"]))


def test_quirks_visible_in_the_golden_numbers():
    # order 0 == order 1 (monopole), and the dipole changes the acceleration but not the potential at order >= 2
    assert np.array_equal(G["m2p_pot_o0"], G["m2p_pot_o1"]) and np.array_equal(G["m2p_acc_o0"], G["m2p_acc_o1"])
    m = G["m2p_moments"][0].copy()
    d = G["m2p_dxyz"][0]
    p_with, a_with = O.m2p(m, d, 3)
    m[1:4] = 0.0
    p_without, a_without = O.m2p(m, d, 3)
    assert p_with == p_without and not np.array_equal(a_with, a_without)


# ---- direct.rs: all eight solvers, both size branches (serial symmetric pair loop for N < 512, per-target sums above)
def _direct_case(tag):
    return G[f"direct_{tag}_pos"], G[f"direct_{tag}_mass"], G[f"direct_{tag}_h"], G[f"direct_{tag}_tgt"]


@pytest.mark.parametrize("tag", ["small", "large"])
def test_direct_newtonian_bit_exact(tag):
    pos, m, h, tgt = _direct_case(tag)
    p, a = O.direct(pos, m, None, kernel=None)
    assert np.array_equal(bits(a), bits(G[f"direct_{tag}_acc"])) and np.array_equal(bits(p), bits(G[f"direct_{tag}_pot"]))
    _, a1 = O.direct(pos, None, None, kernel=None, want=2)
    assert np.array_equal(bits(a1), bits(G[f"direct_{tag}_acc_unit"]))
    pq, aq = O.direct(pos, m, None, targets=tgt, kernel=None)
    assert np.array_equal(bits(aq), bits(G[f"direct_{tag}_acc_pts"])) and np.array_equal(bits(pq), bits(G[f"direct_{tag}_pot_pts"]))


@pytest.mark.parametrize("tag", ["small", "large"])
@pytest.mark.parametrize("kind", [0, 1])
def test_direct_softened_bit_exact(tag, kind):
    # signed softenings: self mode uses max(h_i, h_j) of the SIGNED values, at-points max(h_j, 0)
    pos, m, h, tgt = _direct_case(tag)
    p, a = O.direct(pos, m, h, kernel=kind)
    assert np.array_equal(bits(p), bits(G[f"direct_{tag}_k{kind}_pot"]))
    assert np.array_equal(bits(a), bits(G[f"direct_{tag}_k{kind}_acc"]))
    pq, aq = O.direct(pos, m, h, targets=tgt, kernel=kind)
    assert np.array_equal(bits(pq), bits(G[f"direct_{tag}_k{kind}_pot_pts"]))
    assert np.array_equal(bits(aq), bits(G[f"direct_{tag}_k{kind}_acc_pts"]))
    p0, _ = O.direct(pos, m, None, kernel=kind, want=1)
    assert np.array_equal(bits(p0), bits(G[f"direct_{tag}_k{kind}_pot_noh"]))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["small", "large"])
def test_gpu_direct_against_reference_source_goldens(tag):
    # the CUDA path against the reference-produced numbers directly (not via the oracle): float64 mode to ~1e-13
    # (different summation order), fp32 mode within the north_star tolerance
    import pynbodyext._rust as r
    pos, m, h, tgt = _direct_case(tag)

    def rel(a, ref):
        a, ref = np.asarray(a).reshape(len(ref), -1), np.asarray(ref).reshape(len(ref), -1)
        return np.sqrt(((a - ref) ** 2).sum(1) / (ref ** 2).sum(1)).max()

    assert rel(r.direct_accelerations_py(pos, m, precision="f64"), G[f"direct_{tag}_acc"]) < 1e-11
    assert rel(r.direct_potentials_py(pos, m, precision="f64"), G[f"direct_{tag}_pot"]) < 1e-11
    assert rel(r.direct_potentials_at_points_py(pos, tgt, m, precision="f64"), G[f"direct_{tag}_pot_pts"]) < 1e-11
    for kind in (0, 1):
        assert rel(r.direct_potentials_py(pos, m, 0, h, kind, precision="f64"), G[f"direct_{tag}_k{kind}_pot"]) < 1e-11
        assert rel(r.direct_accelerations_py(pos, m, 0, h, kind, precision="f64"), G[f"direct_{tag}_k{kind}_acc"]) < 1e-10
        assert rel(r.direct_accelerations_at_points_py(pos, tgt, m, 0, h, kind, precision="f64"),
                   G[f"direct_{tag}_k{kind}_acc_pts"]) < 1e-10
        assert rel(r.direct_potentials_py(pos, m, 0, h, kind, precision="f32"), G[f"direct_{tag}_k{kind}_pot"]) < 1e-5
        assert rel(r.direct_accelerations_at_points_py(pos, tgt, m, 0, h, kind, precision="f32"),
                   G[f"direct_{tag}_k{kind}_acc_pts"]) < 1e-4


# ---- tree.rs: the oracle against an INDEPENDENT second restatement of its control flow (PyOctree in
# tests/golden/make_reference_exec.py) that evaluates the mechanically translated reference arithmetic. Bit-exact.
TREE_CASES = [str(c) for c in G["tree_cases"]]


def _oracle_tree(tag):
    cap, order, kernel, theta = G[f"tree_{tag}_params"]
    pos = G[f"tree_{tag}_pos"]
    mass = G[f"tree_{tag}_mass"] if G[f"tree_{tag}_mass"].size else None
    h = G[f"tree_{tag}_h"] if G[f"tree_{tag}_h"].size else None
    t = O.Tree(pos, mass, int(cap), int(order), h, int(kernel) if h is not None else None)
    if mass is None:
        t.build_mass(None)  # gravity.rs:210-220 builds payloads only when masses are given; unit masses need build_mass()
    return t, float(theta)


@pytest.mark.parametrize("tag", TREE_CASES)
def test_tree_topology_and_payload_bit_exact(tag):
    t, _ = _oracle_tree(tag)
    topo = t.topology()
    assert np.array_equal(bits(topo["center"]), bits(G[f"tree_{tag}_center"]))
    assert np.array_equal(bits(topo["half"]), bits(G[f"tree_{tag}_half"]))
    assert np.array_equal(topo["first_subnode"], G[f"tree_{tag}_first"])
    assert np.array_equal(topo["next_branch"], G[f"tree_{tag}_next"])
    assert np.array_equal(topo["leaf_count"], G[f"tree_{tag}_leaf_count"])
    ids = np.nonzero(topo["leaf_count"] >= 0)[0]
    flat = np.concatenate([topo["leaf_particles"][topo["leaf_start"][i]:topo["leaf_start"][i] + topo["leaf_count"][i]] for i in ids])
    assert np.array_equal(flat, G[f"tree_{tag}_leaf_particles"])
    pay = t.payload()
    assert np.array_equal(bits(pay["mass"]), bits(G[f"tree_{tag}_bh_mass"]))
    assert np.array_equal(bits(pay["com"]), bits(G[f"tree_{tag}_bh_com"]))
    if f"tree_{tag}_hmax" in G:
        assert np.array_equal(bits(pay["hmax"]), bits(G[f"tree_{tag}_hmax"]))
    if f"tree_{tag}_moments" in G:
        k = pay["moments"].shape[1]  # compact storage keeps the first k coefficients (MultipoleMoments::from_full)
        assert np.array_equal(bits(pay["moments"]), bits(G[f"tree_{tag}_moments"][:, :k]))


@pytest.mark.parametrize("tag", TREE_CASES)
def test_tree_walk_results_bit_exact(tag):
    t, theta = _oracle_tree(tag)
    p, a = t.eval(theta)
    assert np.array_equal(bits(p), bits(G[f"tree_{tag}_pot"]))
    assert np.array_equal(bits(a), bits(G[f"tree_{tag}_acc"]))
    pq, aq = t.eval(theta, targets=G[f"tree_{tag}_pts"])
    assert np.array_equal(bits(pq), bits(G[f"tree_{tag}_pot_pts"]))
    assert np.array_equal(bits(aq), bits(G[f"tree_{tag}_acc_pts"]))
