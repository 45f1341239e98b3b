"""BASELINE.json configurations at their FULL sizes on the GPU, checked through size-independent properties and
oracle / float64 subsamples (the oracle cannot run the full sizes in test time).

config 2: Hernquist N = 1e6, Plummer eps = 0.01, direct sum fp32 vs float64 reference on a target subsample
config 3: NFW + disc N = 1e7, spline per-particle softening, tree theta 0.5 / 0.7, order 3, leaf 8 — against the
          oracle's tree built from the same 1e7 particles (bit-exact tree, equal walks on a 2e4-target subsample)
config 4/5 shape: dm/gas/star zoom set at N = 1e7 (the oracle cannot hold 1e8): oracle tree, self + (R,z) grid
          targets, float64 direct sum as truth for the grid potentials. The N = 1e8 runs themselves are measured and
          spot-checked by bench.py --gpus 8 (tree_1e8 section).
"""
import numpy as np
import pytest

from benchmarks.synthetic import hernquist, nfw_disc, rz_grid_targets, zoom_range, zoom_set

pytestmark = pytest.mark.gpu


def rms_rel_vec(a, ref):
    return np.sqrt((((a - ref) ** 2).sum(1) / (ref ** 2).sum(1)).mean())


def rms_rel(p, ref):
    return np.sqrt((((p - ref) / ref) ** 2).mean())


@pytest.fixture(scope="module")
def config2():
    pos, m = hernquist(1_000_000, seed=2)
    return pos, m, np.full(1_000_000, 0.01)


def test_config2_direct_fp32_vs_f64_oracle_subsample(config2):
    from oracle import oracle as O
    from pynbodyext.gravity import Gravity, KernelKind
    pos, m, h = config2
    g = Gravity(pos, m, softening=0.01, kernel=KernelKind.Plummer)
    acc = g.direct_accelerations()
    pot = g.direct_potentials()
    idx = np.random.default_rng(0).choice(len(m), 1000, replace=False)
    # oracle at-points on the subsample: the self term of a softened pair at zero separation is exactly 0 for
    # accelerations and -m/eps for the potential, which the self-mode call skips by index
    p_o, a_o = O.direct(pos, m, h, targets=np.ascontiguousarray(pos[idx]), kernel=0)
    p_o = p_o + m[idx] / 0.01
    assert rms_rel_vec(acc[idx], a_o) < 1e-5  # north_star tolerance; measured ~7e-8
    assert rms_rel(pot[idx], p_o) < 1e-5
    # Newton's third law over the whole set: sum_i m_i a_i = 0 (pairwise antisymmetry survives fp32 to ~1e-7)
    net = (m[:, None] * acc).sum(0)
    assert np.linalg.norm(net) / (m * np.linalg.norm(acc, axis=1)).sum() < 1e-6
    # float64 GPU mode on the same subsample agrees with the oracle to rounding
    import pynbodyext._rust as r
    a64 = r.direct_accelerations_at_points_py(pos, np.ascontiguousarray(pos[idx]), m, 0, h, 0, precision="f64")
    assert rms_rel_vec(a64, a_o) < 1e-11


def test_config2_target_shards_reassemble(config2):
    # multi-GPU decomposition property: target shards of the self-mode sum concatenate to the full result (the fp32
    # tile partials are grouped differently per launch shape, so equality holds to fp32-accumulation level, not bitwise)
    import torch
    from pynbodyext.gravity import device as gdev
    pos, m, h = config2
    d = torch.device("cuda", 0)
    dp, dm, dh = (torch.from_numpy(a).to(d) for a in (pos, m, h))
    _, full = gdev.direct_device(dp, dm, dh, kernel=0, want=2)
    parts = [gdev.direct_device(dp, dm, dh, kernel=0, want=2, tgt_begin=lo, count=hi - lo)[1]
             for lo, hi in ((0, 250_000), (250_000, 600_001), (600_001, 1_000_000))]
    got = torch.cat(parts)
    rel = (got - full).norm(dim=1) / full.norm(dim=1)
    assert float(rel.max()) < 1e-5 and float(rel.pow(2).mean().sqrt()) < 1e-6


def flat_leaves(topo):
    """Particle lists of all leaves, concatenated in ascending node id (vectorised leaf_sets for 1e7-size trees)."""
    ids = np.nonzero(topo["leaf_count"] >= 0)[0]
    start, cnt = topo["leaf_start"][ids], topo["leaf_count"][ids]
    off = np.concatenate([[0], np.cumsum(cnt)])
    k = np.arange(off[-1]) - np.repeat(off[:-1], cnt) + np.repeat(start, cnt)
    return ids, cnt, topo["leaf_particles"][k]


def assert_same_tree_as_oracle(g, o):
    """Topology (reference numbering, links, geometry, path keys, leaf particle lists) and payloads: bit-exact."""
    ig, io = g.info(), o.info()
    assert ig["n_nodes"] == io["n_nodes"] and ig["n_leaves"] == io["n_leaves"] and ig["depth"] == io["depth"]
    tg, to = g.topology(), o.topology()
    for k in ("center", "half", "depth", "first_subnode", "next_branch", "leaf_count", "path_hi", "path_lo"):
        assert np.array_equal(tg[k], to[k]), k
    for x, y in zip(flat_leaves(tg), flat_leaves(to)):
        assert np.array_equal(x, y)
    del tg, to
    pg, po = g.payload(), o.payload()
    for k in ("mass", "com", "hmax", "moments"):
        assert np.array_equal(pg[k], po[k]), k


def check_against_oracle_subsample(tree, otree, theta, lo, cnt, grid):
    """GPU fp32 / f64 walk vs the ORACLE's walk on a target subsample: identical interaction lists (counters equal),
    results within 1e-5 (fp32, north_star) / 1e-11 (f64). Self targets [lo, lo+cnt) and at-points grid targets."""
    p_o, a_o, c_o = otree.eval(theta, begin=lo, count=cnt, counters=True)
    c_g = tree.walk_counters(theta, tgt_begin=lo, count=cnt)
    # the oracle counts both of its passes (potential + acceleration); the traversal is the same in both
    assert {k: 2 * c_g[k] for k in c_o} == c_o
    p32, a32 = tree._eval(None, theta, 3, tgt_begin=lo, count=cnt)
    p64, a64 = tree._eval(None, theta, 3, tgt_begin=lo, count=cnt, precision="f64")
    assert rms_rel(p32, p_o) < 1e-5 and rms_rel_vec(a32, a_o) < 1e-5
    assert rms_rel(p64, p_o) < 1e-11 and rms_rel_vec(a64, a_o) < 1e-11
    pq_o, aq_o, cq_o = otree.eval(theta, targets=grid, counters=True)
    cq_g = tree.walk_counters(theta, points=grid)
    assert {k: 2 * cq_g[k] for k in cq_o} == cq_o
    pq32, aq32 = tree._eval(grid, theta, 3)
    pq64, aq64 = tree._eval(grid, theta, 3, precision="f64")
    assert rms_rel(pq32, pq_o) < 1e-5 and rms_rel_vec(aq32, aq_o) < 1e-5
    assert rms_rel(pq64, pq_o) < 1e-11 and rms_rel_vec(aq64, aq_o) < 1e-11
    return p64, a64, pq64, aq64


@pytest.fixture(scope="module")
def config3():
    import pynbodyext._rust as r
    pos, m, h = nfw_disc(10_000_000, seed=3)
    return pos, m, h, r.Octree(pos, m, 8, 3, h, 1)


def test_config3_matches_oracle_tree_at_full_size(config3):
    # BASELINE config 3 at its full size against the ORACLE (serial reference build, ~15 s at 1e7): bit-exact tree,
    # equal interaction lists and results on 2e4 self targets + 2e4 grid points, theta 0.5 and 0.7
    from oracle import oracle as O
    pos, m, h, tree = config3
    ot = O.Tree(pos, m, 8, 3, h, 1)
    assert_same_tree_as_oracle(tree, ot)
    grid = rz_grid_targets(20_000, seed=5, rmax=1.0)
    for theta in (0.5, 0.7):
        check_against_oracle_subsample(tree, ot, theta, 4_000_000, 20_000, grid)


@pytest.fixture(scope="module")
def config4():
    import pynbodyext._rust as r
    pos, m, h = zoom_set(10_000_000, seed=4)
    return pos, m, h, r.Octree(pos, m, 8, 3, h, 1)


def test_config4_zoom_matches_oracle_tree(config4):
    # BASELINE configs 4 / 5 shape (dm/gas/star zoom set, spline softening, gas h ∝ spacing, theta 0.7, order 3) at
    # N = 1e7: depth >= 18 and a softened core where the hmax gate opens whole subtrees (tree.rs:55-71) — the regime
    # of the N = 1e8 runs. Bit-exact tree, equal interaction lists, results vs the oracle on self and grid targets.
    from oracle import oracle as O
    pos, m, h, tree = config4
    ot = O.Tree(pos, m, 8, 3, h, 1)
    assert tree.info()["depth"] >= 18
    assert_same_tree_as_oracle(tree, ot)
    grid = rz_grid_targets(20_000, seed=5)
    p64, a64, pq64, aq64 = check_against_oracle_subsample(tree, ot, 0.7, 5_000_000, 20_000, grid)
    # config 5's truth: float64 DIRECT sum at the grid points (at-points: h = max(h_j, 0), no self term) — the tree's
    # truncation error at theta 0.7 / order 3 is far below 1e-2 (single_node.rs bound); measured median ~1e-4
    import pynbodyext._rust as r
    sub = grid[::10]
    p_d = r.direct_potentials_at_points_py(pos, np.ascontiguousarray(sub), m, 0, h, 1, precision="f64")
    a_d = r.direct_accelerations_at_points_py(pos, np.ascontiguousarray(sub), m, 0, h, 1, precision="f64")
    ep = np.abs(pq64[::10] - p_d) / np.abs(p_d)
    ea = np.linalg.norm(aq64[::10] - a_d, axis=1) / np.linalg.norm(a_d, axis=1)
    assert np.median(ep) < 2e-4 and ep.max() < 1e-2
    assert np.median(ea) < 2e-3 and np.percentile(ea, 90) < 1e-2


def test_zoom_set_is_independent_of_the_rank_layout():
    # the N = 1e8 benchmark generates each rank's shard separately: any split reproduces the same global set
    n = 3_000_000
    pos, m, h = zoom_set(n, seed=4)
    for world in (2, 8):
        b = [(n * r) // world for r in range(world + 1)]
        parts = [zoom_range(n, lo, hi, seed=4) for lo, hi in zip(b[:-1], b[1:])]
        assert np.array_equal(np.concatenate([p[0] for p in parts]), pos)
        assert np.array_equal(np.concatenate([p[1] for p in parts]), m)
        assert np.array_equal(np.concatenate([p[2] for p in parts]), h)


def test_config3_tree_structure_and_payload_invariants(config3):
    pos, m, h, tree = config3
    info = tree.info()
    assert info["n_particles"] == 10_000_000 and 0.3 < info["n_nodes"] / 1e7 < 0.6
    pay = tree.payload()
    assert pay["mass"][0] == pytest.approx(m.sum(), rel=1e-12)
    assert pay["com"][0] == pytest.approx((pos * m[:, None]).sum(0) / m.sum(), rel=1e-9, abs=1e-12)
    assert pay["hmax"][0] == h.max()
    topo = tree.topology()
    assert np.array_equal(np.sort(topo["leaf_particles"]), np.arange(10_000_000))
    assert topo["leaf_count"].max() <= 8


@pytest.mark.parametrize("theta", [0.5, 0.7])
def test_config3_tree_fp32_vs_f64_and_direct(config3, theta):
    import pynbodyext._rust as r
    pos, m, h, tree = config3
    lo, cnt = 4_000_000, 200_000
    p32, a32 = tree._eval(None, theta, 3, tgt_begin=lo, count=cnt)
    p64, a64 = tree._eval(None, theta, 3, tgt_begin=lo, count=cnt, precision="f64")
    assert rms_rel(p32, p64) < 1e-5 and rms_rel_vec(a32, a64) < 1e-5  # same interaction lists, fp32 vs f64 arithmetic
    # tree vs float64 direct sum on 2000 of those targets: truncation error of theta / order 3, far below 1e-2
    idx = lo + np.random.default_rng(1).choice(cnt, 2000, replace=False)
    q = np.ascontiguousarray(pos[idx])
    p_d = r.direct_potentials_at_points_py(pos, q, m, 0, h, 1, precision="f64")
    a_d = r.direct_accelerations_at_points_py(pos, q, m, 0, h, 1, precision="f64")
    # at-points direct has h = max(h_j, 0) and includes the particle itself (r = 0 -> spline W2(0)/h self potential,
    # zero self force); the tree self-mode skips it and uses h = max(h_i, h_j): compare accelerations only through
    # particles whose own softening is not the larger one -> use accelerations with a loose truncation bound
    err = np.linalg.norm(a64[idx - lo] - a_d, axis=1) / np.linalg.norm(a_d, axis=1)
    assert np.median(err) < (2e-3 if theta == 0.7 else 8e-4)
    assert np.isfinite(p_d).all()


def test_config3_block_cyclic_shards_cover_everything(config3):
    # the 8-GPU decomposition, emulated on one GPU through the host API's contiguous shards: results reassemble exactly
    pos, m, h, tree = config3
    full = tree.compute_potentials(0.7)
    for lo, hi in ((0, 1_250_000), (8_750_000, 10_000_000)):
        part = tree._eval(None, 0.7, 1, tgt_begin=lo, count=hi - lo)[0]
        assert np.array_equal(part, full[lo:hi])
    c = tree.walk_counters(0.7)
    assert 300 < c["visits"] / 1e7 < 3000 and c["accepts"] < c["visits"]
