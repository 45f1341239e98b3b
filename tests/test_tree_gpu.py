"""GPU parity of the octree: topology + payloads bit-exact against the oracle's dump, tree results
against the oracle's tree at the same (theta, leaf_capacity, order, kernel).

Tolerances: node numbering, links, centres, half sizes, path keys, leaf particle lists, node mass,
COM, hmax and multipole moments: EXACT (float64 bit equality). Tree potentials/accelerations:
float64 verification mode <= 1e-11 RMS relative; fp32 interaction arithmetic <= 1e-5 RMS relative
(BASELINE.json north_star), measured ~1e-7.
"""
import numpy as np
import pytest

from benchmarks.synthetic import hernquist, nfw_disc, plummer, uniform_cube
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["lane", "warp", "hybrid"])
def walk_kernel_choice(request, monkeypatch):
    """Every test runs with the fp32 walk kernels in all three arrangements: one target per lane (large calls), one
    target per warp (calls with few targets; PNBX_WPT_MAX_TARGETS is the switch-over size, default 16384 particles /
    131072 query points), and — query points only — the hybrid split of larger point sets (warps of 32 points whose walk in the
    lane-per-target kernel outgrows PNBX_WALK_HYBRID_COST are handed to the warp-per-target kernel; lowered here so
    that both parts are populated at test sizes)."""
    monkeypatch.setenv("PNBX_WPT_MAX_TARGETS", "4000000000" if request.param == "warp" else "0")
    monkeypatch.setenv("PNBX_WALK_HYBRID_COST", "6000" if request.param == "hybrid" else "0")

TOL32 = 1e-5
TOL64 = 1e-11


def rms_rel_vec(a, ref):
    return np.sqrt((((a - ref) ** 2).sum(1) / (ref ** 2).sum(1)).mean())


def rms_rel(p, ref):
    return np.sqrt((((p - ref) / ref) ** 2).mean())


def R():
    import pynbodyext._rust as r
    return r


def leaf_sets(topo):
    out = {}
    for i in np.nonzero(topo["leaf_count"] >= 0)[0]:
        s, c = topo["leaf_start"][i], topo["leaf_count"][i]
        out[int(i)] = topo["leaf_particles"][s:s + c].tolist()
    return out


def assert_same_topology(g, o):
    tg, to = g.topology(), o.topology()
    ig, io = g.info(), o.info()
    assert ig["n_nodes"] == io["n_nodes"] and ig["n_leaves"] == io["n_leaves"] and ig["depth"] == io["depth"]
    for k in ("center", "half", "depth", "first_subnode", "next_branch", "leaf_count", "path_hi", "path_lo"):
        assert np.array_equal(tg[k], to[k]), k
    assert leaf_sets(tg) == leaf_sets(to)  # same particles, same (ascending) order in every leaf


def assert_same_payload(g, o):
    pg, po = g.payload(), o.payload()
    assert np.array_equal(pg["mass"], po["mass"])
    assert np.array_equal(pg["com"], po["com"])
    if po["hmax"] is not None:
        assert np.array_equal(pg["hmax"], po["hmax"])
    assert pg["moments"].shape == po["moments"].shape
    assert np.array_equal(pg["moments"], po["moments"])


@pytest.mark.parametrize("n,cap,order", [(1, 8, 3), (7, 8, 0), (9, 8, 3), (300, 1, 2), (5000, 8, 3), (5000, 32, 5),
                                         (20000, 8, 4), (3000, 128, 1),
                                         # leaf capacities > 32: sliding-window maximum (van Herk) + global leaf ordering
                                         (1000, 33, 3), (50000, 200, 2), (4097, 4096, 0), (2000, 2000, 3), (6000, 1000, 2)])
def test_topology_and_payload_bit_exact(n, cap, order):
    pos, m = plummer(n, seed=100 + n)
    g = R().Octree(pos, m, cap, order)
    o = O.Tree(pos, m, cap, order)
    assert_same_topology(g, o)
    assert_same_payload(g, o)


def test_topology_clustered_deep_tree_uses_second_key_word():
    # two tight clumps far apart: depth > 21 forces the 42-level key path
    rng = np.random.default_rng(3)
    a = rng.normal(0.0, 1e-9, (40, 3))
    b = rng.normal(0.0, 1e-9, (40, 3)) + 1.0
    pos = np.concatenate([a, b, rng.random((200, 3)) * 4 - 2])
    m = rng.random(len(pos)) + 0.5
    g = R().Octree(pos, m, 4, 3)
    o = O.Tree(pos, m, 4, 3)
    assert o.info()["depth"] > 21
    assert_same_topology(g, o)
    assert_same_payload(g, o)
    p_g = g.compute_potentials(0.6)
    p_o, _ = o.eval(0.6, want=1)
    assert rms_rel(p_g, p_o) < TOL32


def test_coincident_points_exceeding_leaf_capacity_raise():
    pos = np.zeros((20, 3))
    pos[10:] = 1.0
    with pytest.raises(ValueError, match="deeper than 42 levels"):
        R().Octree(pos, np.ones(20), 4, 0)


def test_unit_mass_and_softening_payloads():
    pos, _ = uniform_cube(4000, 5, with_masses=False)
    h = np.random.default_rng(6).uniform(0.0, 0.1, 4000)
    g = R().Octree(pos, None, 8, 3, h, 1)
    o = O.Tree(pos, None, 8, 3, h, 1)
    assert_same_topology(g, o)
    with pytest.raises(ValueError, match="mass payload not built; call build_mass\\(\\) before compute_potentials"):
        g.compute_potentials(0.5)
    g.build_mass()
    o.build_mass()
    assert_same_payload(g, o)
    p_g, p_o = g.compute_potentials(0.5), o.eval(0.5, want=1)[0]
    assert rms_rel(p_g, p_o) < TOL32


ORDERS = [0, 1, 2, 3, 4, 5]


@pytest.mark.parametrize("order", ORDERS)
def test_self_eval_matches_oracle_tree(order):
    pos, m = hernquist(6000, seed=21)
    g = R().Octree(pos, m, 8, order)
    o = O.Tree(pos, m, 8, order)
    for theta in (0.5, 0.7):
        p_o, a_o = o.eval(theta)
        p64, a64 = g._eval(None, theta, 3, precision="f64")
        assert rms_rel(p64, p_o) < TOL64 and rms_rel_vec(a64, a_o) < TOL64
        p32, a32 = g._eval(None, theta, 3)
        assert rms_rel(p32, p_o) < TOL32 and rms_rel_vec(a32, a_o) < TOL32
        # the fused (pot+acc) and single-output kernels may contract multiply-adds differently: fp32 rounding level
        assert rms_rel(g.compute_potentials(theta), p32) < 1e-6 and rms_rel_vec(g.compute_accelerations(theta), a32) < 1e-6


@pytest.mark.parametrize("order", [0, 3, 5])
def test_at_points_matches_oracle_tree(order):
    pos, m = plummer(5000, seed=22)
    q, _ = plummer(1500, seed=23, a=2.0)
    q[:5] = pos[:5]  # points on top of particles: no skip in at-points mode (tree.rs:1516)
    q[5] = [1e3, -2e3, 5e2]  # far outside the root cube
    g = R().Octree(pos, m, 8, order, np.full(5000, 0.02), 0)
    o = O.Tree(pos, m, 8, order, np.full(5000, 0.02), 0)
    p_o, a_o = o.eval(0.7, targets=q)
    p64, a64 = g._eval(q, 0.7, 3, precision="f64")
    assert rms_rel(p64, p_o) < TOL64 and rms_rel_vec(a64, a_o) < TOL64
    assert rms_rel(g.potentials_at_points(q, 0.7), p_o) < TOL32
    assert rms_rel_vec(g.accelerations_at_points(q, 0.7), a_o) < TOL32


def test_hybrid_point_walk_is_a_per_point_choice_of_the_two_kernels(monkeypatch):
    # Larger query sets start in the lane-per-target kernel; a warp (32 consecutive points in path-key order) whose walk
    # outgrows PNBX_WALK_HYBRID_COST gives up and its points are walked one per warp by the other kernel. Each point's
    # value is bit-equal to what the kernel that finished it gives for the whole set.
    from benchmarks.synthetic import rz_grid_targets, zoom_set
    n = 200_000
    pos, m, h = zoom_set(n, seed=4)
    q = rz_grid_targets(30_011, seed=5)  # not a multiple of 32: a partial trailing warp
    t = R().Octree(pos, m, 8, 3, h, 1)

    def run(wpt, cost):
        monkeypatch.setenv("PNBX_WPT_MAX_TARGETS", wpt)
        monkeypatch.setenv("PNBX_WALK_HYBRID_COST", cost)
        return t._eval(q, 0.7, 3)

    lane, warp = run("0", "0"), run("4000000000", "0")
    mixed = 0
    for cost in ("2000", "8000", "16000", "32000", "64000", "1000000000"):
        hyb = run("0", cost)
        eq_lane = (hyb[0] == lane[0]) & (hyb[1] == lane[1]).all(1)
        eq_warp = (hyb[0] == warp[0]) & (hyb[1] == warp[1]).all(1)
        assert (eq_lane | eq_warp).all()
        mixed += bool((~eq_lane).any() and (~eq_warp).any())
    assert mixed >= 2  # thresholds at which both classes are populated
    o = O.Tree(pos, m, 8, 3, h, 1)
    sub = np.random.default_rng(3).choice(q.shape[0], 400, replace=False)
    p_o, a_o = o.eval(0.7, targets=q[sub])
    hyb = run("0", "16000")
    assert rms_rel(hyb[0][sub], p_o) < TOL32 and rms_rel_vec(hyb[1][sub], a_o) < TOL32
    # default threshold, default switch-over: same values again from whichever kernels are chosen, to fp32 accuracy
    monkeypatch.delenv("PNBX_WPT_MAX_TARGETS"); monkeypatch.delenv("PNBX_WALK_HYBRID_COST")
    assert rms_rel(t.potentials_at_points(q[sub], 0.7), p_o) < TOL32


@pytest.mark.parametrize("kernel", [0, 1])
@pytest.mark.parametrize("hmode", ["const", "var"])
def test_softened_tree_matches_oracle(kernel, hmode):
    n = 8000
    pos, m, h = nfw_disc(n, seed=31)
    if hmode == "const":
        h = np.full(n, 0.01)
    else:
        h = h * np.random.default_rng(7).uniform(1.0, 30.0, n)
    g = R().Octree(pos, m, 8, 3, h, kernel)
    o = O.Tree(pos, m, 8, 3, h, kernel)
    assert_same_topology(g, o)
    assert_same_payload(g, o)
    p_o, a_o = o.eval(0.7)
    p64, a64 = g._eval(None, 0.7, 3, precision="f64")
    assert rms_rel(p64, p_o) < TOL64 and rms_rel_vec(a64, a_o) < TOL64
    p32, a32 = g._eval(None, 0.7, 3)
    assert rms_rel(p32, p_o) < TOL32 and rms_rel_vec(a32, a_o) < TOL32
    q, _ = plummer(500, seed=8, a=0.2)
    p_o, a_o = o.eval(0.7, targets=q)
    p32, a32 = g._eval(q, 0.7, 3)
    assert rms_rel(p32, p_o) < TOL32 and rms_rel_vec(a32, a_o) < TOL32


def test_reference_property_theta0_tree_equals_direct():
    # gravity_tests.rs:57-126 on the GPU path (float64 mode, the reference's 1e-10 bound)
    r = R()
    pos, m = uniform_cube(256, 1)
    g = r.Octree(pos, m, 32, 2)
    p, a = g._eval(None, 0.0, 3, precision="f64")
    p_d, a_d = O.direct(pos, m)
    assert np.abs(p - p_d).max() < 1e-10 and np.abs(a - a_d).max() < 1e-10
    q, _ = uniform_cube(128, 13, with_masses=False)
    src, ms = uniform_cube(512, 11)
    g = r.Octree(src, ms, 32, 2)
    p, a = g._eval(q, 0.0, 3, precision="f64")
    p_d, a_d = O.direct(src, ms, targets=q)
    assert np.abs(p - p_d).max() < 1e-10 and np.abs(a - a_d).max() < 1e-10


def test_reference_property_error_decreases_with_order():
    # gravity_tests.rs:128-202 on the GPU path
    r = R()
    pos, m = uniform_cube(800, 21)
    p_ref, a_ref = O.direct(pos, m)
    ea, ep = [], []
    for order in (0, 3, 4, 5):
        a = r.Octree(pos, m, 64, order).compute_accelerations(0.7)
        ea.append(np.sqrt(((a - a_ref) ** 2).sum(1).mean()))
    for order in (0, 2, 3, 4, 5):
        p = r.Octree(pos, m, 64, order).compute_potentials(0.7)
        ep.append(np.sqrt(((p - p_ref) ** 2).mean()))
    assert all(ea[i] <= ea[i - 1] for i in range(1, len(ea))) and ea[-1] <= 0.8 * ea[0]
    assert all(ep[i] <= ep[i - 1] for i in range(1, len(ep)))


def test_setters_follow_reference_semantics():
    # set_softenings does not rebuild hmax (tree.rs:777-782); set_kernel switches leaf sums + gate factor
    r = R()
    pos, m = plummer(3000, seed=41)
    h0 = np.full(3000, 0.05)
    g = r.Octree(pos, m, 8, 3, h0, 0)
    o = O.Tree(pos, m, 8, 3, h0, 0)
    h1 = np.random.default_rng(1).uniform(0.0, 0.3, 3000)
    g.set_softenings(h1); o.set_softenings(h1)
    g.set_kernel(1); o.set_kernel(1)
    p_o, a_o = o.eval(0.7)
    p, a = g._eval(None, 0.7, 3, precision="f64")
    assert rms_rel(p, p_o) < TOL64 and rms_rel_vec(a, a_o) < TOL64
    g.set_softenings(None); o.set_softenings(None)
    p_o, _ = o.eval(0.7, want=1)
    assert rms_rel(g._eval(None, 0.7, 1, precision="f64")[0], p_o) < TOL64
    m2 = m * np.random.default_rng(2).uniform(0.5, 2.0, 3000)
    g.build_mass(m2); o.build_mass(m2)
    assert_same_payload(g, o)


def test_shard_eval_equals_full_eval():
    r = R()
    pos, m = hernquist(10000, seed=51)
    g = r.Octree(pos, m, 8, 3)
    p_full, a_full = g._eval(None, 0.7, 3)
    for lo, hi in ((0, 2500), (2500, 7001), (7001, 10000)):
        p, a = g._eval(None, 0.7, 3, tgt_begin=lo, count=hi - lo)
        assert np.array_equal(p, p_full[lo:hi]) and np.array_equal(a, a_full[lo:hi])


def test_gravity_api_tree_path_and_cache_semantics():
    from pynbodyext.gravity import Gravity, KernelKind
    pos, m = plummer(4000, seed=61)
    g = Gravity(pos, m, softening=0.01, kernel=KernelKind.Spline, leaf_capacity=16, multipole_order=2)
    # tree_* use their own defaults (8, 3): a throw-away tree, not the cached (16, 2) one
    p = g.tree_potentials(theta=0.6)
    assert g._tree is None
    o = O.Tree(pos, m, 8, 3, np.full(4000, 0.01), 1)
    assert rms_rel(p, o.eval(0.6, want=1)[0]) < TOL32
    p2 = g.tree_potentials(theta=0.6, leaf_capacity=16, multipole_order=2)
    assert g._tree is not None
    o2 = O.Tree(pos, m, 16, 2, np.full(4000, 0.01), 1)
    assert rms_rel(p2, o2.eval(0.6, want=1)[0]) < TOL32
    a = g.tree_accelerations(positions=pos[:100] + 0.01, theta=0.7)
    a_o = O.Tree(pos, m, 8, 3, np.full(4000, 0.01), 1).eval(0.7, targets=pos[:100] + 0.01, want=2)[1]
    assert rms_rel_vec(a, a_o) < TOL32


def test_bench_gravity_shape_config():
    # benchmarks/bench_gravity.py:148-159: theta 0.7, softening 0.001, kernel=1, order 3, leaf 8, construct + potentials
    from pynbodyext.gravity import Gravity
    pos, m = hernquist(15682, seed=71)  # size of halo 0 of subhalo_103 (tests/conftest.py:52)
    g = Gravity(pos, m, softening=0.001, kernel=1, leaf_capacity=8, multipole_order=3)
    p = g.tree_potentials(theta=0.7, leaf_capacity=8, multipole_order=3, kernel=1)
    o = O.Tree(pos, m, 8, 3, np.full(len(m), 0.001), 1)
    assert rms_rel(p, o.eval(0.7, want=1)[0]) < TOL32


def test_larger_tree_1e5_subsample_against_oracle():
    pos, m = plummer(100_000, seed=1)
    g = R().Octree(pos, m, 8, 3)
    o = O.Tree(pos, m, 8, 3)
    assert_same_topology(g, o)
    assert_same_payload(g, o)
    p_o, a_o = o.eval(0.7)
    p, a = g._eval(None, 0.7, 3)
    assert rms_rel(p, p_o) < TOL32 and rms_rel_vec(a, a_o) < TOL32
    assert np.abs((p - p_o) / p_o).max() < 1e-4


def test_walk_counters_equal_oracle_counters():
    # lane-exact acceptance: the GPU walk visits / accepts / sums exactly what tree.rs:1069-1134 would
    r = R()
    pos, m = hernquist(20000, seed=81)
    h = np.full(20000, 0.005)
    g = r.Octree(pos, m, 8, 3, h, 1)
    o = O.Tree(pos, m, 8, 3, h, 1)
    for theta in (0.5, 0.7, 1.0):
        _, _, c_o = o.eval(theta, want=1, counters=True)
        c_g = g.walk_counters(theta)
        assert {k: c_g[k] for k in c_o} == c_o and c_g["warp_visits"] * 32 >= c_g["visits"]
    q, _ = plummer(3000, seed=82, a=3.0)
    _, _, c_o = o.eval(0.7, targets=q, want=1, counters=True)
    c_g = g.walk_counters(0.7, points=q)
    assert {k: c_g[k] for k in c_o} == c_o


def test_device_tree_api_matches_host_api():
    import torch
    from pynbodyext.gravity import device as gdev
    pos, m = plummer(30000, seed=91)
    h = np.full(30000, 0.01)
    host = R().Octree(pos, m, 8, 3, h, 0)
    d = torch.device("cuda", 0)
    tree = gdev.OctreeDevice(torch.from_numpy(pos).to(d), torch.from_numpy(m).to(d), 8, 3, torch.from_numpy(h).to(d), 0)
    assert tree.info()["n_nodes"] == host.info()["n_nodes"]
    p_h, a_h = host._eval(None, 0.7, 3)
    p_d, a_d = tree.eval(0.7, 3)
    torch.cuda.synchronize()
    assert np.array_equal(p_d.cpu().numpy(), p_h) and np.array_equal(a_d.cpu().numpy(), a_h)
    p_s, _ = tree.eval(0.7, 3, tgt_begin=1000, count=5000)
    assert np.array_equal(p_s.cpu().numpy(), p_h[1000:6000])
    # tree-order shards: same numbers, delivered in tree order with the scatter map
    p_t, a_t = tree.eval(0.7, 3, tgt_begin=7000, count=9000, tree_order=True)
    idx = tree.order(7000, 9000).cpu().numpy()
    assert np.array_equal(p_t.cpu().numpy(), p_h[idx]) and np.array_equal(a_t.cpu().numpy(), a_h[idx])
    assert np.array_equal(np.sort(tree.order().cpu().numpy()), np.arange(30000))
    # block-cyclic shards of 3 ranks: disjoint, complete, identical numbers
    seen = []
    for rk in range(3):
        p_s, a_s = tree.eval(0.7, 3, shard=(rk, 3))
        ix = tree.order(shard=(rk, 3)).cpu().numpy()
        assert len(ix) == gdev.shard_count(30000, 3, rk)
        assert np.array_equal(p_s.cpu().numpy(), p_h[ix]) and np.array_equal(a_s.cpu().numpy(), a_h[ix])
        seen.append(ix)
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(30000))
    q = torch.from_numpy(pos[:777] * 1.5).to(d)
    p_q, _ = tree.eval(0.7, 1, targets=q)
    assert np.array_equal(p_q.cpu().numpy(), host.potentials_at_points(pos[:777] * 1.5, 0.7))  # same kernel variant


def test_sharded_entry_points_single_rank():
    # world = 1 path of the multi-GPU API (the all-gather is a no-op); multi-rank host logic is in test_sharded_gloo.py
    from pynbodyext.gravity.sharded import direct_sharded, tree_sharded
    pos, m = hernquist(20000, seed=92)
    h = np.full(20000, 0.01)
    p, a, (lo, hi) = direct_sharded(pos, m, h, kernel=0, want=3, rank=0, world=1, device=0)
    assert (lo, hi) == (0, 20000)
    p_o, a_o = O.direct(pos, m, h, kernel=0)
    assert rms_rel(p, p_o) < TOL32 and rms_rel_vec(a, a_o) < TOL32
    p, a, idx = tree_sharded(pos, m, h, kernel=0, want=3, theta=0.7, rank=0, world=1, device=0)
    o = O.Tree(pos, m, 8, 3, h, 0)
    p_o, a_o = o.eval(0.7)
    assert np.array_equal(np.sort(idx), np.arange(20000))
    assert rms_rel(p, p_o[idx]) < TOL32 and rms_rel_vec(a, a_o[idx]) < TOL32


def test_zero_mass_particles_and_nodes():
    # zero-mass nodes are skipped as a whole (tree.rs:1087-1090), zero-mass children are left out of COM / M2M
    # (tree.rs:913, 1051); a clump of massless tracers forms zero-mass leaves and subtrees
    r = R()
    pos, m = plummer(4000, seed=101)
    tracers = np.random.default_rng(102).normal(0.0, 0.01, (600, 3)) + np.array([3.0, 3.0, 3.0])
    pos = np.concatenate([pos, tracers])
    m = np.concatenate([m, np.zeros(600)])
    m[::7] = 0.0
    g = r.Octree(pos, m, 8, 3)
    o = O.Tree(pos, m, 8, 3)
    assert_same_topology(g, o)
    assert_same_payload(g, o)
    assert (o.payload()["mass"] == 0.0).sum() > 10
    p_o, a_o = o.eval(0.7)
    p, a = g._eval(None, 0.7, 3, precision="f64")
    assert rms_rel(p, p_o) < TOL64 and rms_rel_vec(a, a_o) < TOL64
    p, a = g._eval(None, 0.7, 3)
    assert rms_rel(p, p_o) < TOL32 and rms_rel_vec(a, a_o) < TOL32
    c_o = o.eval(0.7, want=1, counters=True)[2]
    c_g = g.walk_counters(0.7)
    assert {k: c_g[k] for k in c_o} == c_o


def test_negative_and_zero_softenings_follow_reference_clamps():
    # tree leaf sums clamp every h at 0 (tree.rs:37,115,234); direct self mode does not
    r = R()
    pos, m = plummer(2000, seed=103)
    h = np.random.default_rng(104).uniform(-0.02, 0.05, 2000)
    h[:100] = 0.0
    for kernel in (0, 1):
        g = r.Octree(pos, m, 8, 3, h, kernel)
        o = O.Tree(pos, m, 8, 3, h, kernel)
        assert_same_payload(g, o)
        p_o, a_o = o.eval(0.7)
        p, a = g._eval(None, 0.7, 3, precision="f64")
        assert rms_rel(p, p_o) < TOL64 and rms_rel_vec(a, a_o) < TOL64
        p_d, a_d = O.direct(pos, m, h, kernel=kernel)
        p = r.direct_potentials_py(pos, m, 0, h, kernel, precision="f64")
        a = r.direct_accelerations_py(pos, m, 0, h, kernel, precision="f64")
        assert rms_rel(p, p_d) < TOL64 and rms_rel_vec(a, a_d) < TOL64


def test_concurrent_queries_on_one_tree_from_threads():
    # compute methods take &self in the reference and release the GIL (gravity.rs:103-111, 267-444): concurrent
    # queries on one Octree from several Python threads are legal; here every thread gets its own stream + temporaries
    import threading
    r = R()
    pos, m = hernquist(40000, seed=111)
    g = r.Octree(pos, m, 8, 3, np.full(40000, 0.01), 1)
    ref_p = g.compute_potentials(0.7)
    ref_a = g.compute_accelerations(0.6)
    q = pos[:5000] * 1.01
    ref_q = g.potentials_at_points(q, 0.7)
    out, errs = {}, []

    def work(k):
        try:
            for _ in range(3):
                if k % 3 == 0:
                    out[k] = np.array_equal(g.compute_potentials(0.7), ref_p)
                elif k % 3 == 1:
                    out[k] = np.array_equal(g.compute_accelerations(0.6), ref_a)
                else:
                    out[k] = np.array_equal(g.potentials_at_points(q, 0.7), ref_q)
        except Exception as exc:  # pragma: no cover
            errs.append(exc)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(6)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs and all(out[k] for k in range(6))


@pytest.mark.parametrize("cap", [1, 3, 8, 32])
@pytest.mark.parametrize("kernel", [0, 1])
def test_leaf_runs_dense_softened_core(cap, kernel):
    # A core whose softening exceeds the inter-particle spacing: the hmax gate (tree.rs:55-71) opens whole
    # subtrees, most work is direct leaf sums with r < h. The walk merges sibling leaves into runs and sums
    # them in fp32 with squared softenings (tree_walk.cu); lists and results must still be the reference's.
    n = 6000
    pos, m = hernquist(n, seed=131, a=0.05)
    rng = np.random.default_rng(132)
    h = np.where(rng.uniform(size=n) < 0.5, 0.02, rng.uniform(1e-4, 5e-2, n))
    g = R().Octree(pos, m, cap, 3, h, kernel)
    o = O.Tree(pos, m, cap, 3, h, kernel)
    assert_same_topology(g, o)
    p_o, a_o = o.eval(0.7)
    c_o = o.eval(0.7, want=1, counters=True)[2]  # one traversal (want=3 would count the pot and acc walks)
    c_g = g.walk_counters(0.7)
    assert {k: c_g[k] for k in c_o} == c_o
    assert c_o["leaf_particles"] > 50 * n  # the regime this test is about
    p64, a64 = g._eval(None, 0.7, 3, precision="f64")
    assert rms_rel(p64, p_o) < TOL64 and rms_rel_vec(a64, a_o) < TOL64
    p32, a32 = g._eval(None, 0.7, 3)
    assert rms_rel(p32, p_o) < TOL32 and rms_rel_vec(a32, a_o) < TOL32
    assert np.abs((p32 - p_o) / p_o).max() < 20 * TOL32
    # single-output kernels agree with the fused one to fp32 rounding
    p1 = g.compute_potentials(0.7)
    a2 = g.compute_accelerations(0.7)
    assert rms_rel(p1, p_o) < TOL32 and rms_rel_vec(a2, a_o) < TOL32
    q = pos[:400] + rng.normal(0.0, 1e-3, (400, 3))
    p_q, a_q = o.eval(0.7, targets=q)
    p32, a32 = g._eval(q, 0.7, 3)
    assert rms_rel(p32, p_q) < TOL32 and rms_rel_vec(a32, a_q) < TOL32


def test_asv_benchmark_classes_of_the_reference_run_and_match_the_oracle():
    # benchmarks/asv_gravity.py mirrors the reference's ASV classes (bench_gravity.py:74-166); pynbody's test data are
    # absent here, so it runs on the synthetic stand-in of halo 0 (15 682 particles). Every timed call works, and the
    # `TimeTreeGravityFull` / `main()` configuration (theta 0.7, softening 0.001, spline, order 3) matches the oracle.
    import itertools
    from benchmarks import asv_gravity as B
    b = B.TimeTreeConstruct()
    assert len(b.pos) == 15_682
    for lc, soft, order in itertools.product([8, 128], [None, 0.288], [0, 5]):
        b.time_construct_tree(lc, soft, order)
    B.TimeTreeGravityTheta().time_construct_and_tree_potentials(1.0)
    B.TimeTreeGravityOrder().time_construct_and_tree_potentials(4)
    full = B.TimeTreeGravityFull()
    full.time_construct_and_tree_potentials()
    p = full._construct_and_tree_potentials(theta=0.7, softening=0.001, kernel=1, multipole_order=3)
    o = O.Tree(full.pos, full.mass, 8, 3, np.full(len(full.mass), 0.001), 1)
    assert rms_rel(p, o.eval(0.7, want=1)[0]) < TOL32
    # theta-only case: Gravity(kernel=None, order 0) -> unsoftened monopole tree
    p0 = B.TimeTreeGravityTheta()._construct_and_tree_potentials(theta=0.7)
    assert rms_rel(p0, O.Tree(full.pos, full.mass, 8, 0).eval(0.7, want=1)[0]) < TOL32
