"""Round-2 additions around the boundary: device-pointer setters (`pnbx_tree_build_mass_ex` /
`pnbx_tree_set_softenings_ex`), the library's caching allocator (`pnbx_trim_memory`), the `precision=` plumbing
(ADVICE: the reference is float64 throughout), stream / lifetime ordering of trees used from several streams, and the
current-device guard."""
import numpy as np
import pytest

from benchmarks.synthetic import hernquist, plummer
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def rms_rel(p, ref):
    return np.sqrt((((p - ref) / ref) ** 2).mean())


def rms_rel_vec(a, ref):
    return np.sqrt((((a - ref) ** 2).sum(1) / (ref ** 2).sum(1)).mean())


def test_device_setters_match_host_setters():
    import torch
    import pynbodyext._rust as r
    from pynbodyext.gravity import device as gdev
    n = 20000
    pos, m = plummer(n, seed=3)
    h = np.random.default_rng(1).uniform(0.005, 0.05, n)
    d = torch.device("cuda", 0)
    host = r.Octree(pos, m, 8, 3, h, 1)
    dev = gdev.OctreeDevice(torch.from_numpy(pos).to(d), torch.from_numpy(m).to(d), 8, 3, torch.from_numpy(h).to(d), 1)
    m2 = m * np.random.default_rng(2).uniform(0.5, 2.0, n)
    h2 = np.ascontiguousarray(h * 1.7)
    host.build_mass(m2); dev.build_mass(torch.from_numpy(m2).to(d))
    host.set_softenings(h2); dev.set_softenings(torch.from_numpy(h2).to(d))
    host.set_kernel(0); dev.set_kernel(0)
    p_h, a_h = host._eval(None, 0.7, 3)
    p_d, a_d = dev.eval(0.7, 3)
    assert np.array_equal(p_d.cpu().numpy(), p_h) and np.array_equal(a_d.cpu().numpy(), a_h)
    # against the oracle driven through the same setter sequence (hmax is NOT rebuilt by set_softenings, tree.rs:777-782)
    o = O.Tree(pos, m, 8, 3, h, 1)
    o.build_mass(m2); o.set_softenings(h2); o.set_kernel(0)
    p_o, a_o = o.eval(0.7)
    assert rms_rel(p_h, p_o) < 1e-5 and rms_rel_vec(a_h, a_o) < 1e-5
    with pytest.raises(ValueError, match="masses must be length N"):
        dev.build_mass(torch.zeros(5, dtype=torch.float64, device=d))


def test_tree_can_be_dropped_while_walks_on_other_streams_are_queued():
    # ADVICE (round 1): tree buffers were freed on the creation stream with no ordering against evaluations queued on
    # other streams. Now every evaluation makes the tree's stream wait for it: results stay correct when the tree is
    # destroyed (and its memory immediately reused by the next build) right after the calls were queued.
    import torch
    from pynbodyext.gravity import device as gdev
    d = torch.device("cuda", 0)
    pos, m = hernquist(200_000, seed=4)
    dp, dm = torch.from_numpy(pos).to(d), torch.from_numpy(m).to(d)
    ref = gdev.OctreeDevice(dp, dm, 8, 3).eval(0.7, 1)[0].clone()
    streams = [torch.cuda.Stream(d) for _ in range(3)]
    torch.cuda.synchronize()
    outs = []
    for rep in range(4):
        tree = gdev.OctreeDevice(dp, dm, 8, 3)
        for s in streams:
            with torch.cuda.stream(s):
                outs.append(tree.eval(0.7, 1)[0])
        del tree  # frees are ordered after the three queued walks
    torch.cuda.synchronize()
    for o in outs:
        assert torch.equal(o, ref)


def test_trim_memory_and_reuse():
    import pynbodyext._rust as r
    from pynbodyext.gravity import device as gdev
    pos, m = plummer(50_000, seed=9)
    a = r.Octree(pos, m, 8, 3).compute_potentials(0.7)
    gdev.trim_memory()          # nothing is in use: every cached block goes back to the driver
    b = r.Octree(pos, m, 8, 3).compute_potentials(0.7)
    gdev.trim_memory()
    gdev.trim_memory()          # idempotent
    c = r.direct_potentials_py(pos[:4000], m[:4000])
    assert np.array_equal(a, b) and np.isfinite(c).all()


def test_precision_plumbing_and_auto_default(monkeypatch):
    import pynbodyext._rust as r
    from pynbodyext.gravity import Gravity, KernelKind
    pos, m = plummer(3000, seed=21)
    p_o, a_o = O.direct(pos, m)
    # unsoftened small direct sums default to float64 (the reference's small-N validation use of method="direct")
    assert r.resolve_precision(None, unsoftened_pairs=3000 * 3000) == "f64" and r.resolve_precision(None) == "f32"
    g = Gravity(pos, m)
    assert rms_rel(g.direct_potentials(), p_o) < 1e-12 and rms_rel_vec(g.direct_accelerations(), a_o) < 1e-12
    p32 = g.direct_potentials(precision="f32")
    assert 1e-12 < rms_rel(p32, p_o) < 1e-5
    # Gravity(precision=...) is the default of its methods, a per-call value overrides it
    h = np.full(3000, 0.02)
    g64 = Gravity(pos, m, softening=0.02, kernel=KernelKind.Plummer, precision="f64")
    p_os, _ = O.direct(pos, m, h, kernel=0, want=1)
    assert rms_rel(g64.direct_potentials(), p_os) < 1e-12
    assert rms_rel(g64.direct_potentials(precision="f32"), p_os) > 1e-12
    t_o = O.Tree(pos, m, 8, 3, h, 0).eval(0.7, want=1)[0]
    assert rms_rel(g64.tree_potentials(), t_o) < 1e-11
    assert 1e-12 < rms_rel(Gravity(pos, m, softening=0.02, kernel=KernelKind.Plummer).tree_potentials(), t_o) < 1e-5
    # process-wide default
    monkeypatch.setenv("PNBX_PRECISION", "f64")
    assert rms_rel(Gravity(pos, m, softening=0.02, kernel=KernelKind.Plummer).direct_potentials(), p_os) < 1e-12
    monkeypatch.setenv("PNBX_PRECISION", "bogus")
    with pytest.raises(ValueError, match="PNBX_PRECISION"):
        g.direct_potentials()
    with pytest.raises(ValueError, match="precision must be"):
        g.direct_potentials(precision="f16")


def test_calls_leave_the_current_device_alone():
    import torch
    import pynbodyext._rust as r
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    pos, m = plummer(5000, seed=1)
    torch.cuda.set_device(0)
    t = r.Octree(pos, m, 8, 3, device=1)
    p = t.compute_potentials(0.7)
    r.direct_potentials_py(pos, m, device=1)
    assert torch.cuda.current_device() == 0
    del t
    assert torch.cuda.current_device() == 0
    assert np.array_equal(p, r.Octree(pos, m, 8, 3, device=0).compute_potentials(0.7))


def test_direct_calls_are_stream_ordered_without_host_syncs():
    # the variant decision (constant vs per-particle softening, equal masses) is made on the device: queuing many
    # device-pointer calls on one stream gives the same results as running them one by one
    import torch
    from pynbodyext.gravity import device as gdev
    d = torch.device("cuda", 0)
    pos, m = hernquist(30_000, seed=8)
    rng = np.random.default_rng(5)
    cases = [(np.full(30_000, 0.01), 0), (rng.uniform(0.005, 0.02, 30_000), 0), (rng.uniform(0.005, 0.02, 30_000), 1),
             (None, None), (rng.uniform(-0.01, 0.02, 30_000), 0)]
    dp, dm = torch.from_numpy(pos).to(d), torch.from_numpy(m * rng.uniform(0.9, 1.1, 30_000)).to(d)
    queued = []
    for h, k in cases * 3:
        dh = None if h is None else torch.from_numpy(h).to(d)
        queued.append(gdev.direct_device(dp, dm, dh, kernel=k, want=3))
    torch.cuda.synchronize()
    for i, (h, k) in enumerate(cases):
        dh = None if h is None else torch.from_numpy(h).to(d)
        p, a = gdev.direct_device(dp, dm, dh, kernel=k, want=3)
        torch.cuda.synchronize()
        for rep in range(3):
            assert torch.equal(queued[i + rep * len(cases)][0], p) and torch.equal(queued[i + rep * len(cases)][1], a)
        p_o, a_o = O.direct(pos, dm.cpu().numpy(), h, kernel=k)
        tol = 1e-5 if h is not None else 1e-3  # unsoftened fp32: close pairs below the coordinate resolution
        assert rms_rel(p.cpu().numpy(), p_o) < tol and rms_rel_vec(a.cpu().numpy(), a_o) < tol


def test_gravity_init_aliases_contiguous_float64_inputs():
    # documented difference: the reference's Gravity.__init__ copies (`astype`, base.py:199-200); here
    # float64 C-contiguous inputs are NOT copied — the device upload at call time is the copy — so a caller who mutates
    # its arrays between construction and a call sees the mutation. Other dtypes / layouts are converted (= copied).
    from pynbodyext.gravity import Gravity
    pos, m = plummer(2000, seed=5)
    g = Gravity(pos, m)
    assert g.pos is pos or np.shares_memory(g.pos, pos)
    before = g.direct_potentials(precision="f64")
    pos *= 2.0  # same shape, all separations doubled -> potentials halved
    after = g.direct_potentials(precision="f64")
    assert np.allclose(after, 0.5 * before, rtol=1e-12)
    g32 = Gravity(pos.astype(np.float32), m)  # converted: an independent float64 copy
    assert not np.shares_memory(g32.pos, pos)
