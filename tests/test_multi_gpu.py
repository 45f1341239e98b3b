"""PNBX_DEVICES: several GPUs behind the unchanged host API (csrc/multi.cu).

The multi-device result must be BIT-IDENTICAL to the single-GPU shard calls it is made of: direct sums are the
contiguous target shards (tgt_begin / count), tree self-evaluations are a pure function of (tree, target), so the
assembled array equals the single-GPU evaluation exactly. Needs >= 2 GPUs (skipped otherwise): run with
`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`.
"""
import os

import numpy as np
import pytest

from benchmarks.synthetic import hernquist, nfw_disc, rz_grid_targets

pytestmark = pytest.mark.gpu


def _ndev():
    import pynbodyext._rust as r
    return r._load().pnbx_device_count()


@pytest.fixture()
def multi_env():
    if _ndev() < 2:
        pytest.skip("needs >= 2 GPUs")
    keys = ("PNBX_DEVICES", "PNBX_MULTI_MIN_WORK", "PNBX_MULTI_MIN_N", "PNBX_MULTI_MIN_TARGETS")
    saved = {k: os.environ.get(k) for k in keys}

    def on(devs="all"):
        os.environ.update(PNBX_DEVICES=devs, PNBX_MULTI_MIN_WORK="0", PNBX_MULTI_MIN_N="0", PNBX_MULTI_MIN_TARGETS="0")

    def off():
        os.environ.pop("PNBX_DEVICES", None)

    yield on, off
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _shards(n, w):
    return [(n * r) // w for r in range(w + 1)]


@pytest.mark.parametrize("mode", ["plummer_const", "spline_pair", "newton"])
def test_direct_multi_equals_single_gpu_shards(multi_env, mode):
    import pynbodyext._rust as r
    from pynbodyext._rust import _load, _opts, _ptr
    import ctypes as C
    on, off = multi_env
    n = 40_003
    pos, m = hernquist(n, seed=11)
    rng = np.random.default_rng(3)
    h = {"plummer_const": np.full(n, 0.01), "spline_pair": rng.uniform(0.005, 0.05, n), "newton": None}[mode]
    kern = {"plummer_const": 0, "spline_pair": 1, "newton": None}[mode]
    on()
    w = _ndev()
    acc_m = r.direct_accelerations_py(pos, m, 0, h, kern)
    pot_m = r.direct_potentials_py(pos, m, 0, h, kern)
    q = np.ascontiguousarray(pos[::7] * 1.01)
    pot_q = r.direct_potentials_at_points_py(pos, q, m, 0, h, kern)
    off()
    b = _shards(n, min(w, 8))
    for lo, hi in zip(b[:-1], b[1:]):
        cnt = hi - lo
        o = _opts(0, None)
        acc1 = np.empty((cnt, 3))
        rc = _load().pnbx_direct(_ptr(pos), _ptr(m), _ptr(h), n, None, cnt, lo, -1 if kern is None else kern, 2,
                                 None, _ptr(acc1), C.byref(o))
        assert rc == 0
        assert np.array_equal(acc_m[lo:hi], acc1)
        pot1 = np.empty(cnt)
        rc = _load().pnbx_direct(_ptr(pos), _ptr(m), _ptr(h), n, None, cnt, lo, -1 if kern is None else kern, 1,
                                 _ptr(pot1), None, C.byref(o))
        assert rc == 0
        assert np.array_equal(pot_m[lo:hi], pot1)
    bq = _shards(len(q), min(w, 8))
    for lo, hi in zip(bq[:-1], bq[1:]):
        ref = r.direct_potentials_at_points_py(pos, np.ascontiguousarray(q[lo:hi]), m, 0, h, kern, device=0)
        assert np.array_equal(pot_q[lo:hi], ref)


def test_tree_multi_equals_single_gpu(multi_env):
    import pynbodyext._rust as r
    on, off = multi_env
    n = 300_007
    pos, m, h = nfw_disc(n, seed=5)
    q = rz_grid_targets(5000, seed=5, rmax=1.0)
    off()
    t1 = r.Octree(pos, m, 8, 3, h, 1, device=0)
    p1, a1 = t1.compute_potentials(0.7), t1.compute_accelerations(0.7)
    pq1 = t1.potentials_at_points(q, 0.7)
    on()
    tm = r.Octree(pos, m, 8, 3, h, 1)
    pm, am = tm.compute_potentials(0.7), tm.compute_accelerations(0.7)
    pqm = tm.accelerations_at_points(q, 0.7), tm.potentials_at_points(q, 0.7)
    assert np.array_equal(pm, p1) and np.array_equal(am, a1)
    assert np.array_equal(pqm[1], pq1)
    # topology of the primary copy is the reference topology (same as the single-GPU build)
    assert tm.info() == t1.info()
    # setters reach every copy: kernel switch (gate factor), new masses, new softenings
    tm.set_kernel(0); t1.set_kernel(0)
    assert np.array_equal(tm.compute_potentials(0.6), t1.compute_potentials(0.6))
    m2 = m * np.random.default_rng(1).uniform(0.5, 1.5, n)
    tm.build_mass(m2); t1.build_mass(m2)
    assert np.array_equal(tm.compute_accelerations(0.7), t1.compute_accelerations(0.7))
    h2 = np.ascontiguousarray(h * 2.0)
    tm.set_softenings(h2); t1.set_softenings(h2)
    assert np.array_equal(tm.compute_potentials(0.7), t1.compute_potentials(0.7))
    # partial ranges stay on the primary device and still work
    part = tm._eval(None, 0.7, 1, tgt_begin=1000, count=5000)[0]
    assert np.array_equal(part, t1._eval(None, 0.7, 1, tgt_begin=1000, count=5000)[0])
    del tm, t1


def test_gravity_api_uses_all_devices(multi_env):
    # the reference-facing call, unchanged: Gravity(...).direct_accelerations() / tree_potentials()
    from pynbodyext.gravity import Gravity, KernelKind
    on, off = multi_env
    pos, m = hernquist(30_000, seed=2)
    off()
    g = Gravity(pos, m, softening=0.01, kernel=KernelKind.Plummer)
    a1, p1 = g.direct_accelerations(), g.tree_potentials(theta=0.7)
    on()
    g2 = Gravity(pos, m, softening=0.01, kernel=KernelKind.Plummer)
    a2, p2 = g2.direct_accelerations(), g2.tree_potentials(theta=0.7)
    assert p2.shape == p1.shape and np.array_equal(p2, p1)
    # direct: shard launches differ from the one full launch in their split shape -> fp32-accumulation level
    rel = np.linalg.norm(a2 - a1, axis=1) / np.linalg.norm(a1, axis=1)
    assert rel.max() < 1e-5


def test_bad_device_list_is_an_error(multi_env):
    import pynbodyext._rust as r
    on, off = multi_env
    on("0,99")
    pos, m = hernquist(2000, seed=2)
    with pytest.raises(ValueError, match="PNBX_DEVICES"):
        r.direct_potentials_py(pos, m)


def test_point_evaluation_hybrid_walk_multi_equals_single(multi_env, monkeypatch):
    # larger point sets: lane-per-target walk with a cost budget per warp + hand-over to the warp-per-target kernel.
    # The devices take blocks of 256 points of the path-key order, i.e. whole warps of the single-device call: the same
    # warps give up, so the assembled result is bit-identical to the single-device one.
    import pynbodyext._rust as r
    on, off = multi_env
    monkeypatch.setenv("PNBX_WPT_MAX_TARGETS", "0")
    monkeypatch.setenv("PNBX_WALK_HYBRID_COST", "6000")
    n = 300_007
    pos, m, h = nfw_disc(n, seed=5)
    q = rz_grid_targets(20_000, seed=6, rmax=1.0)
    off()
    t1 = r.Octree(pos, m, 8, 3, h, 1, device=0)
    p1, a1 = t1._eval(q, 0.7, 3)
    monkeypatch.setenv("PNBX_WALK_HYBRID_COST", "0")
    p_lane = t1._eval(q, 0.7, 1)[0]
    monkeypatch.setenv("PNBX_WALK_HYBRID_COST", "6000")
    assert 0 < (p1 != p_lane).sum() < q.shape[0]  # some warps were handed over, not all
    on()
    tm = r.Octree(pos, m, 8, 3, h, 1)
    pm, am = tm._eval(q, 0.7, 3)
    assert np.array_equal(pm, p1) and np.array_equal(am, a1)
    del tm, t1
