"""Golden fixtures (tests/golden/*.npz, generated from the CPU oracle by tests/golden/make_golden.py — the
reference itself cannot be built here).

not gpu: the oracle still reproduces its frozen outputs bit for bit (catches accidental edits of the oracle).
gpu:     the CUDA path matches the frozen values without needing the oracle at run time: topology/payload
         bit-exact, results within the fp32 tolerance (1e-5 RMS) / fp64 tolerance (1e-11).
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
D = np.load(os.path.join(HERE, "golden", "direct_golden.npz"))
T = np.load(os.path.join(HERE, "golden", "tree_golden.npz"))

DIRECT = (("newton", None, False), ("plummer", 0, True), ("spline", 1, True))


@pytest.fixture(autouse=True, params=["lane", "warp", "hybrid"])
def walk_kernel_choice(request, monkeypatch):
    """Every test runs with the fp32 walk kernels in all three arrangements: one target per lane (large calls), one
    target per warp (calls with few targets; PNBX_WPT_MAX_TARGETS is the switch-over size, default 16384 particles /
    131072 query points), and — query points only — the hybrid split of larger point sets (warps of 32 points whose walk in the
    lane-per-target kernel outgrows PNBX_WALK_HYBRID_COST are handed to the warp-per-target kernel; lowered here so
    that both parts are populated at test sizes)."""
    monkeypatch.setenv("PNBX_WPT_MAX_TARGETS", "4000000000" if request.param == "warp" else "0")
    monkeypatch.setenv("PNBX_WALK_HYBRID_COST", "6000" if request.param == "hybrid" else "0")


def rms_rel_vec(a, ref):
    return np.sqrt((((a - ref) ** 2).sum(1) / (ref ** 2).sum(1)).mean())


def rms_rel(p, ref):
    return np.sqrt((((p - ref) / ref) ** 2).mean())


def test_oracle_reproduces_direct_golden():
    from oracle import oracle as O
    for name, kern, use_h in DIRECT:
        h = D["h"] if use_h else None
        p, a = O.direct(D["pos"], D["mass"], h, kernel=kern)
        assert np.array_equal(p, D[f"{name}_self_pot"]) and np.array_equal(a, D[f"{name}_self_acc"])
        p, a = O.direct(D["pos"], D["mass"], h, targets=D["q"], kernel=kern)
        assert np.array_equal(p, D[f"{name}_pts_pot"]) and np.array_equal(a, D[f"{name}_pts_acc"])
    p, a = O.direct(D["small_pos"], D["small_mass"])
    assert np.array_equal(p, D["small_pot"]) and np.array_equal(a, D["small_acc"])


def test_oracle_reproduces_tree_golden():
    from oracle import oracle as O
    for order in (0, 2, 3, 5):
        t = O.Tree(T["pos"], T["mass"], 8, order, T["h"], 1)
        p, a = t.eval(0.7)
        assert np.array_equal(p, T[f"o{order}_self_pot"]) and np.array_equal(a, T[f"o{order}_self_acc"])
        p, a = t.eval(0.7, targets=T["q"])
        assert np.array_equal(p, T[f"o{order}_pts_pot"]) and np.array_equal(a, T[f"o{order}_pts_acc"])
    t = O.Tree(T["pos"], T["mass"], 8, 3, T["h"], 1)
    topo = t.topology()
    for k in ("center", "half", "depth", "first_subnode", "next_branch", "leaf_count", "path_hi", "path_lo"):
        assert np.array_equal(topo[k], T[f"topo_{k}"]), k
    pay = t.payload()
    assert np.array_equal(pay["moments"], T["pay_moments"]) and np.array_equal(pay["com"], T["pay_com"])


@pytest.mark.gpu
def test_gpu_direct_matches_golden():
    import pynbodyext._rust as r
    for name, kern, use_h in DIRECT:
        h = D["h"] if use_h else None
        p = r.direct_potentials_py(D["pos"], D["mass"], 0, h, kern)
        a = r.direct_accelerations_py(D["pos"], D["mass"], 0, h, kern)
        assert rms_rel(p, D[f"{name}_self_pot"]) < 1e-5 and rms_rel_vec(a, D[f"{name}_self_acc"]) < 1e-5
        p = r.direct_potentials_at_points_py(D["pos"], D["q"], D["mass"], 0, h, kern, precision="f64")
        a = r.direct_accelerations_at_points_py(D["pos"], D["q"], D["mass"], 0, h, kern, precision="f64")
        assert rms_rel(p, D[f"{name}_pts_pot"]) < 1e-11 and rms_rel_vec(a, D[f"{name}_pts_acc"]) < 1e-11
    p = r.direct_potentials_py(D["small_pos"], D["small_mass"])
    assert rms_rel(p, D["small_pot"]) < 1e-5


@pytest.mark.gpu
def test_gpu_tree_matches_golden():
    import pynbodyext._rust as r
    for order in (0, 2, 3, 5):
        g = r.Octree(T["pos"], T["mass"], 8, order, T["h"], 1)
        p, a = g._eval(None, 0.7, 3)
        assert rms_rel(p, T[f"o{order}_self_pot"]) < 1e-5 and rms_rel_vec(a, T[f"o{order}_self_acc"]) < 1e-5
        p, a = g._eval(T["q"], 0.7, 3, precision="f64")
        assert rms_rel(p, T[f"o{order}_pts_pot"]) < 1e-11 and rms_rel_vec(a, T[f"o{order}_pts_acc"]) < 1e-11
    g = r.Octree(T["pos"], T["mass"], 8, 3, T["h"], 1)
    topo = g.topology()
    for k in ("center", "half", "depth", "first_subnode", "next_branch", "leaf_count", "path_hi", "path_lo"):
        assert np.array_equal(topo[k], T[f"topo_{k}"]), k
    leaves = np.nonzero(topo["leaf_count"] >= 0)[0]
    lp = np.concatenate([topo["leaf_particles"][topo["leaf_start"][i]: topo["leaf_start"][i] + topo["leaf_count"][i]]
                         for i in leaves])
    assert np.array_equal(lp, T["topo_leaf_particles_in_node_order"])
    pay = g.payload()
    assert np.array_equal(pay["mass"], T["pay_mass"]) and np.array_equal(pay["com"], T["pay_com"])
    assert np.array_equal(pay["hmax"], T["pay_hmax"]) and np.array_equal(pay["moments"], T["pay_moments"])
    c = g.walk_counters(0.7)
    assert [c["visits"], c["accepts"], c["leaf_visits"], c["leaf_particles"]] == T["counters"].tolist()
