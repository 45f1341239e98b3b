"""GPU parity: the CUDA direct-summation path (through the drop-in API / C-ABI) against the CPU oracle.

Tolerances (stated per BASELINE.json north_star): fp32 interaction arithmetic vs the float64
reference sum: RMS over targets of |da|/|a| and |dphi|/|phi| <= 1e-5 (measured ~1e-7);
float64 verification mode (precision="f64"): <= 1e-11.
"""
import numpy as np
import pytest

from benchmarks.synthetic import hernquist, plummer, uniform_cube
from oracle import oracle as O

pytestmark = pytest.mark.gpu

TOL32 = 1e-5
TOL64 = 1e-11


@pytest.fixture(autouse=True, params=["caller_order", "sorted"])
def direct_particle_order(request, monkeypatch):
    """Self calls with per-particle softenings (and spline sums at points) sort the particles — by softening for Plummer,
    along a Morton curve for the spline — so that whole source tiles resolve max(h_i, h_j) / the r < h test at once;
    the default only does so from 65536 particles (PNBX_DIRECT_SORT_MIN). Every test runs both ways."""
    monkeypatch.setenv("PNBX_DIRECT_SORT_MIN", "0" if request.param == "sorted" else "-1")


def rms_rel_vec(a, ref):
    return np.sqrt((((a - ref) ** 2).sum(1) / (ref ** 2).sum(1)).mean())


def rms_rel(p, ref):
    return np.sqrt((((p - ref) / ref) ** 2).mean())


def backend():
    import pynbodyext._rust as r
    return r


CASES = [
    ("newton", None, None),
    ("plummer_const", 0, "const"),
    ("plummer_pair", 0, "var"),
    ("spline_const", 1, "const"),
    ("spline_pair", 1, "var"),
]


def _soft(mode, n, seed):
    if mode is None:
        return None
    if mode == "const":
        return np.full(n, 0.03)
    return np.random.default_rng(seed).uniform(0.005, 0.15, n)


@pytest.mark.parametrize("name,kernel,hmode", CASES)
@pytest.mark.parametrize("n", [300, 511, 1023, 2000, 5003])  # 511 / 1023: odd N whose padded pair count fills the last tile
def test_self_matches_oracle(name, kernel, hmode, n):
    r = backend()
    pos, m = plummer(n, seed=n)
    h = _soft(hmode, n, n + 1)
    p_o, a_o = O.direct(pos, m, h, kernel=kernel)
    p = r.direct_potentials_py(pos, m, 0, h, kernel)
    a = r.direct_accelerations_py(pos, m, 0, h, kernel)
    assert rms_rel(p, p_o) < TOL32 and rms_rel_vec(a, a_o) < TOL32
    assert np.abs((p - p_o) / p_o).max() < 1e-4
    p64 = r.direct_potentials_py(pos, m, 0, h, kernel, precision="f64")
    a64 = r.direct_accelerations_py(pos, m, 0, h, kernel, precision="f64")
    assert rms_rel(p64, p_o) < TOL64 and rms_rel_vec(a64, a_o) < TOL64


@pytest.mark.parametrize("name,kernel,hmode", CASES)
def test_at_points_matches_oracle(name, kernel, hmode):
    r = backend()
    n, mq = 3001, 777
    pos, m = hernquist(n, seed=3)
    q, _ = plummer(mq, seed=4, a=0.5)
    h = _soft(hmode, n, 5)
    p_o, a_o = O.direct(pos, m, h, targets=q, kernel=kernel)
    p = r.direct_potentials_at_points_py(pos, q, m, 0, h, kernel)
    a = r.direct_accelerations_at_points_py(pos, q, m, 0, h, kernel)
    assert rms_rel(p, p_o) < TOL32 and rms_rel_vec(a, a_o) < TOL32
    p64 = r.direct_potentials_at_points_py(pos, q, m, 0, h, kernel, precision="f64")
    a64 = r.direct_accelerations_at_points_py(pos, q, m, 0, h, kernel, precision="f64")
    assert rms_rel(p64, p_o) < TOL64 and rms_rel_vec(a64, a_o) < TOL64


def test_unit_masses_when_none():
    r = backend()
    pos, _ = uniform_cube(1000, 9, with_masses=False)
    p_o, a_o = O.direct(pos, None)
    assert rms_rel(r.direct_potentials_py(pos), p_o) < TOL32
    assert rms_rel_vec(r.direct_accelerations_py(pos), a_o) < TOL32


def test_self_skip_is_by_index_not_distance():
    # two distinct particles at the same place, softened: the reference keeps their mutual term
    # (skip is `j == i`, direct.rs:421) — a distance-based skip would drop it.
    r = backend()
    pos, m = uniform_cube(600, 10)
    pos[17] = pos[400]
    h = np.full(600, 0.05)
    p_o, a_o = O.direct(pos, m, h, kernel=0)
    p = r.direct_potentials_py(pos, m, 0, h, 0)
    a = r.direct_accelerations_py(pos, m, 0, h, 0)
    assert abs(p[17] - p_o[17]) / abs(p_o[17]) < 1e-5
    assert rms_rel(p, p_o) < TOL32 and rms_rel_vec(a, a_o) < TOL32


def test_empty_and_tiny_inputs():
    r = backend()
    e = np.zeros((0, 3))
    assert r.direct_potentials_py(e).shape == (0,)
    assert r.direct_accelerations_py(e).shape == (0, 3)
    pos = np.array([[0.0, 0.0, 0.0], [2.0, 0.0, 0.0]])
    m = np.array([3.0, 5.0])
    assert r.direct_potentials_py(pos, m) == pytest.approx([-2.5, -1.5], rel=1e-6)
    # no sources -> zeros at the targets (direct.rs:195-197)
    assert np.all(r.direct_potentials_at_points_py(e, pos) == 0.0)
    # a single particle feels nothing
    assert r.direct_potentials_py(pos[:1], m[:1])[0] == 0.0


def test_offset_coordinates_are_recentred():
    # a system far from the origin: float32 coordinates would lose the separations without recentring
    r = backend()
    pos, m = plummer(2000, seed=11, a=1e-3)
    pos = pos + np.array([5.0e3, -7.0e3, 9.0e3])
    p_o, a_o = O.direct(pos, m)
    assert rms_rel(r.direct_potentials_py(pos, m), p_o) < TOL32
    assert rms_rel_vec(r.direct_accelerations_py(pos, m), a_o) < 3e-5


def test_deterministic_run_to_run():
    r = backend()
    pos, m = plummer(20000, seed=12)
    a1 = r.direct_accelerations_py(pos, m)
    a2 = r.direct_accelerations_py(pos, m)
    assert np.array_equal(a1, a2)


def test_config1_plummer_1e5_subsample():
    # BASELINE config 1 (Plummer N=1e5, Newtonian): GPU on all targets, oracle on a 2000-target
    # at-points subsample (identical sums except the skipped self term, which is 0 for j == i here
    # because the oracle at-points call includes a -m/sqrt(TINY) self term — so compare via self-mode
    # oracle on the subsample indices computed with explicit skip).
    r = backend()
    pos, m = plummer(100_000, seed=1)
    from pynbodyext.gravity import Gravity
    g = Gravity(pos, m)
    p = g.direct_potentials()
    a = g.direct_accelerations()
    idx = np.random.default_rng(0).choice(100_000, 500, replace=False)
    for i in idx[:50]:
        d = np.delete(pos, i, axis=0) - pos[i]
        mm = np.delete(m, i)
        r2 = (d * d).sum(1)
        assert abs(p[i] + (mm / np.sqrt(r2)).sum()) / abs(p[i]) < 1e-5
        aref = (mm[:, None] * d / r2[:, None] ** 1.5).sum(0)
        assert np.linalg.norm(a[i] - aref) / np.linalg.norm(aref) < 1e-4


def test_linearity_in_mass_full_size():
    # size-independent property at a larger N: doubling all masses doubles phi and a exactly
    # (power-of-two scaling commutes with every rounding step).
    r = backend()
    pos, m = hernquist(50_000, seed=13)
    h = np.full(50_000, 0.01)
    a1 = r.direct_accelerations_py(pos, m, 0, h, 0)
    a2 = r.direct_accelerations_py(pos, 2.0 * m, 0, h, 0)
    assert np.array_equal(2.0 * a1, a2)


@pytest.mark.parametrize("kernel", [0, 1])
def test_signed_softenings_follow_reference(kernel):
    # direct.rs:402,426 take h = max(h_i, h_j) of the SIGNED values in self mode (Plummer then uses h*h, the spline treats
    # h <= 0 as Newtonian, kernel.rs:46-48); at points h = max(h_j, 0) (direct.rs:560). The packed per-pair kernels
    # work on squared clamped softenings, so self-mode Plummer with a negative entry must take the scalar kernel:
    # results have to agree with the oracle either way, and mixed / all-positive arrays both stay within tolerance.
    r = backend()
    n = 2500
    pos, m = plummer(n, seed=77)
    rng = np.random.default_rng(78)
    for h in (rng.uniform(-0.05, 0.1, n), -rng.uniform(0.01, 0.1, n), rng.uniform(0.0, 0.1, n)):
        p_o, a_o = O.direct(pos, m, h, kernel=kernel)
        p = r.direct_potentials_py(pos, m, 0, h, kernel)
        a = r.direct_accelerations_py(pos, m, 0, h, kernel)
        assert rms_rel(p, p_o) < TOL32 and rms_rel_vec(a, a_o) < TOL32
        q, _ = plummer(300, seed=79, a=0.5)
        p_o, a_o = O.direct(pos, m, h, targets=q, kernel=kernel)
        p = r.direct_potentials_at_points_py(pos, q, m, 0, h, kernel)
        a = r.direct_accelerations_at_points_py(pos, q, m, 0, h, kernel)
        assert rms_rel(p, p_o) < TOL32 and rms_rel_vec(a, a_o) < TOL32


def test_packed_per_pair_kernels_match_scalar_kernels_full_tiles():
    # N large enough that most tiles take the packed FP32x2 loop (and the spline's second pass), shard with an offset
    r = backend()
    n = 20011
    pos, m = hernquist(n, seed=91)
    h = np.random.default_rng(92).uniform(0.002, 0.2, n)
    m = m * np.random.default_rng(93).uniform(0.5, 2.0, n)
    idx = np.random.default_rng(94).choice(n, 1500, replace=False)
    for kernel in (0, 1):
        p_o, a_o = O.direct(pos, m, h, targets=pos[idx], kernel=kernel)  # at-points oracle: h = max(h_j, 0)
        p = r.direct_potentials_at_points_py(pos, pos[idx], m, 0, h, kernel)
        a = r.direct_accelerations_at_points_py(pos, pos[idx], m, 0, h, kernel)
        # targets sit exactly on sources: r = 0 pairs deep inside the softening of their own particle
        assert rms_rel(p, p_o) < TOL32 and rms_rel_vec(a, a_o) < TOL32
        p_s, a_s = O.direct(pos, m, h, kernel=kernel)
        p = r.direct_potentials_py(pos, m, 0, h, kernel)
        a = r.direct_accelerations_py(pos, m, 0, h, kernel)
        assert rms_rel(p, p_s) < TOL32 and rms_rel_vec(a, a_s) < TOL32


@pytest.mark.parametrize("n", [20011, 20012, 700])
def test_packed_per_pair_kernels_equal_mass_variant(n):
    # equal source masses + per-particle softenings take the equal-mass variant of the packed per-pair kernels (mass
    # factored out, the spline zeroes 1/r of inside pairs): odd N (zero-mass pad in the last tile), even N, N < one tile;
    # unit masses (masses=None) are equal masses too
    r = backend()
    pos, m = hernquist(n, seed=191)
    h = np.random.default_rng(192).uniform(0.002, 0.2, n)
    m = np.full(n, 0.37 / n)
    idx = np.random.default_rng(194).choice(n, min(n, 1200), replace=False)
    for kernel in (0, 1):
        p_s, a_s = O.direct(pos, m, h, kernel=kernel)
        assert rms_rel(r.direct_potentials_py(pos, m, 0, h, kernel), p_s) < TOL32
        assert rms_rel_vec(r.direct_accelerations_py(pos, m, 0, h, kernel), a_s) < TOL32
        # targets exactly on sources (r = 0 pairs inside their own softening), as in the general-mass test above; an
        # offset of 1e-3 at |x| ~ 70 would only test the fp32 resolution of box-centred coordinates (4e-6)
        q = np.ascontiguousarray(pos[idx])
        p_o, a_o = O.direct(pos, m, h, targets=q, kernel=kernel)
        assert rms_rel(r.direct_potentials_at_points_py(pos, q, m, 0, h, kernel), p_o) < TOL32
        assert rms_rel_vec(r.direct_accelerations_at_points_py(pos, q, m, 0, h, kernel), a_o) < TOL32
        p_u, a_u = O.direct(pos, None, h, kernel=kernel)
        assert rms_rel(r.direct_potentials_py(pos, None, 0, h, kernel), p_u) < TOL32
        assert rms_rel_vec(r.direct_accelerations_py(pos, None, 0, h, kernel), a_u) < TOL32


def max_rel_vec(a, ref):
    return (np.linalg.norm(a - ref, axis=1) / np.linalg.norm(ref, axis=1)).max()


@pytest.mark.parametrize("kernel", [0, 1])
@pytest.mark.parametrize("pattern", ["random_small", "families", "one_wide"])
@pytest.mark.parametrize("equal_mass", [True, False])
def test_tile_regimes_match_oracle(kernel, pattern, equal_mass):
    # direct_kernel_f2h picks a cheaper loop for whole (target block, source tile) combinations: Plummer with the tile's
    # softenings all below / all above the block's (max(h_i, h_j) known), spline with the two bounding boxes further
    # apart than every softening radius involved (Newtonian). 80 source tiles here; softenings well below the
    # particle spacing so that most spline tiles qualify once the particles are Morton-sorted, contiguous families of
    # equal softening (the usual snapshot layout: the regimes apply without sorting), and one particle with a huge
    # softening radius (its tile and block never qualify, every pair with it is inside). A pair wrongly treated as
    # Newtonian would show up as an O(1) error on its two particles: the check is on the MAXIMUM relative error.
    r = backend()
    n = 40_001
    pos, m = hernquist(n, seed=7)
    rng = np.random.default_rng(11)
    if not equal_mass:
        m = m * rng.uniform(0.5, 2.0, n)
    if pattern == "random_small":
        h = rng.uniform(0.005, 0.02, n)
    elif pattern == "families":
        h = np.concatenate([np.full(n // 3, 0.01), np.full(n // 3, 0.04), np.full(n - 2 * (n // 3), 0.002)])
    else:
        h = rng.uniform(0.005, 0.02, n)
        h[12345] = 50.0
    p_o, a_o = O.direct(pos, m, h, kernel=kernel)
    p = r.direct_potentials_py(pos, m, 0, h, kernel)
    a = r.direct_accelerations_py(pos, m, 0, h, kernel)
    assert rms_rel(p, p_o) < TOL32 and rms_rel_vec(a, a_o) < TOL32
    assert np.abs((p - p_o) / p_o).max() < 1e-4 and max_rel_vec(a, a_o) < 2e-3
    # at-points: h = max(h_j, 0), every Plummer tile takes its softening from the source records
    q = hernquist(3000, seed=8)[0]  # independent points: no separation below the fp32 resolution of the coordinates
    p_q, a_q = O.direct(pos, m, h, targets=q, kernel=kernel)
    assert np.abs((r.direct_potentials_at_points_py(pos, q, m, 0, h, kernel) - p_q) / p_q).max() < 1e-4
    assert max_rel_vec(r.direct_accelerations_at_points_py(pos, q, m, 0, h, kernel), a_q) < 2e-3


def test_sorted_sweep_is_a_reordering_of_the_same_pair_terms(monkeypatch):
    # sorting changes the summation order only: sorted and caller-order results agree to fp32 accumulation accuracy,
    # and constant softenings are not reordered at all by the (stable) Plummer key: bit-equal
    r = backend()
    n = 70_000
    pos, m = hernquist(n, seed=3)
    h = np.random.default_rng(4).uniform(0.005, 0.02, n)
    out = {}
    for mode in ("-1", "0"):
        monkeypatch.setenv("PNBX_DIRECT_SORT_MIN", mode)
        out[mode] = (r.direct_accelerations_py(pos, m, 0, h, 0), r.direct_accelerations_py(pos, m, 0, h, 1),
                     r.direct_accelerations_py(pos, m, 0, np.full(n, 0.01), 0))
    for k in (0, 1):
        assert 0 < max_rel_vec(out["0"][k], out["-1"][k]) < 1e-4
    assert np.array_equal(out["0"][2], out["-1"][2])
    monkeypatch.delenv("PNBX_DIRECT_SORT_MIN")  # default: sorted from 65536 particles
    assert np.array_equal(r.direct_accelerations_py(pos, m, 0, h, 1), out["0"][1])


@pytest.mark.parametrize("kernel", [0, 1])
def test_self_shards_with_per_particle_softening_match_oracle(kernel):
    # a self call on a target shard sorts the shard's own particles to the front of the source arrays (then target k
    # is source k again: self-skip by index) and the rest behind them: every shard of the sum equals the oracle's
    # slice, and the shards concatenate to the whole-array result to fp32 accumulation accuracy
    import torch
    from pynbodyext.gravity import device as gdev
    n = 20_011
    pos, m = hernquist(n, seed=31)
    rng = np.random.default_rng(32)
    h = rng.uniform(0.005, 0.05, n)
    m = m * rng.uniform(0.5, 2.0, n)
    p_o, a_o = O.direct(pos, m, h, kernel=kernel)
    d = torch.device("cuda", 0)
    dp, dm, dh = (torch.from_numpy(x).to(d) for x in (pos, m, h))
    parts_p, parts_a = [], []
    for lo, hi in ((0, 5000), (5000, 12_345), (12_345, n)):
        p, a = gdev.direct_device(dp, dm, dh, kernel=kernel, want=3, tgt_begin=lo, count=hi - lo)
        p, a = p.cpu().numpy(), a.cpu().numpy()
        assert rms_rel(p, p_o[lo:hi]) < TOL32 and rms_rel_vec(a, a_o[lo:hi]) < TOL32
        assert max_rel_vec(a, a_o[lo:hi]) < 2e-3
        parts_p.append(p); parts_a.append(a)
    p_w, a_w = gdev.direct_device(dp, dm, dh, kernel=kernel, want=3)
    assert max_rel_vec(np.concatenate(parts_a), a_w.cpu().numpy()) < 1e-4
    assert np.abs(np.concatenate(parts_p) / p_w.cpu().numpy() - 1).max() < 1e-5
