"""the first in-package consumers of the gravity path (pynbodyext/gravity/consumers.py) —
CenPos(mode="pot") / ShiftPosTo("pot") equivalents, binding energy and rotation curve — fed by the GPU potential /
acceleration instead of a pre-existing sim["phi"] (reference properties/generic.py:51-52, transforms/shift.py:17-24).
Driven with the fake pynbody of tests/fake_pynbody; truth from the CPU oracle."""
import numpy as np
import pytest

from test_snapshot_api import fake_pynbody  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def make_sim(pynbody, n=20000, seed=12, a=2.0, shift=(3.0, -1.0, 0.5)):
    from benchmarks.synthetic import plummer
    pos, m = plummer(n, seed=seed, a=a)
    rng = np.random.default_rng(seed)
    vel = rng.normal(0.0, 50.0, (n, 3)) + np.array([10.0, -20.0, 5.0])
    pos = pos + np.asarray(shift)
    sim = pynbody.snapshot.SimSnap(pos=pynbody.array.SimArray(pos, "kpc"), mass=pynbody.array.SimArray(m * 1e11, "Msol"),
                                   vel=pynbody.array.SimArray(vel, "km s**-1"))
    return sim, pos, m * 1e11, vel


def test_potential_center_and_shift(fake_pynbody):  # noqa: F811
    from oracle import oracle as O
    from pynbodyext.gravity import KernelKind
    from pynbodyext.gravity.consumers import PotentialCenter, shift_to_potential_minimum
    sim, pos, m, _ = make_sim(fake_pynbody)
    h = 0.05
    phi_o, _ = O.Tree(pos, m, 8, 3, np.full(len(m), h), 1).eval(0.7, want=1)
    node = PotentialCenter(softening=h, kernel=KernelKind.Spline)
    cen = node(sim)
    assert np.array_equal(np.asarray(cen), pos[phi_o.argmin()]) and cen.sim is sim
    assert np.linalg.norm(np.asarray(cen) - np.array([3.0, -1.0, 0.5])) < 0.5  # the Plummer centre, within the core
    # an existing phi is used only on request (the reference's behaviour)
    sim._a["phi"] = fake_pynbody.array.SimArray(-np.arange(len(m), dtype=float), "km**2 s**-2")
    assert np.array_equal(np.asarray(PotentialCenter(use_existing=True)(sim)), pos[-1])
    assert np.array_equal(np.asarray(node(sim)), np.asarray(cen))
    pos0 = pos.copy()  # the snapshot's array is a view of `pos`: the in-place shift moves both
    got = shift_to_potential_minimum(sim, softening=h, kernel=KernelKind.Spline)
    assert np.array_equal(np.asarray(got), np.asarray(cen))
    assert np.allclose(np.asarray(sim["pos"]), pos0 - np.asarray(cen), rtol=0, atol=1e-12)
    assert node.instance_signature() == PotentialCenter(softening=h, kernel=KernelKind.Spline).instance_signature()


def test_binding_energy(fake_pynbody):  # noqa: F811
    from oracle import oracle as O
    from pynbodyext.gravity import KernelKind
    from pynbodyext.gravity.consumers import binding_energy
    u = fake_pynbody.units
    sim, pos, m, vel = make_sim(fake_pynbody, n=8000)
    f_pot = u.G.si * u.Msol.si / u.kpc.si / 1e6
    phi_o, _ = O.direct(pos, m, np.full(len(m), 0.05), kernel=0, want=1)
    e = binding_energy(sim, softening=0.05, kernel=KernelKind.Plummer, method="direct")
    vc = (vel * m[:, None]).sum(0) / m.sum()
    ref = 0.5 * ((vel - vc) ** 2).sum(1) + phi_o * f_pot
    assert e.sim is sim and e.units.si == pytest.approx(1e6)
    assert np.allclose(np.asarray(e), ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())
    assert (np.asarray(e) < 0).mean() > 0.5  # most particles of this cold set are bound


def test_rotation_curve(fake_pynbody):  # noqa: F811
    from oracle import oracle as O
    from pynbodyext.gravity import KernelKind
    from pynbodyext.gravity.consumers import rotation_curve
    u = fake_pynbody.units
    a = 2.0
    sim, pos, m, _ = make_sim(fake_pynbody, n=40000, a=a, shift=(0.0, 0.0, 0.0))
    radii = np.array([0.5, 1.0, 2.0, 4.0, 8.0])
    n_phi = 8
    vc = rotation_curve(sim, radii, n_phi=n_phi, softening=0.05, kernel=KernelKind.Spline, theta=0.6)
    # against the oracle's tree accelerations at the same ring points
    ang = 2.0 * np.pi * (np.arange(n_phi) + 0.5) / n_phi
    ring = np.stack([np.cos(ang), np.sin(ang), np.zeros(n_phi)], axis=1)
    pts = (radii[:, None, None] * ring[None]).reshape(-1, 3)
    _, a_o = O.Tree(pos, m, 8, 3, np.full(len(m), 0.05), 1).eval(0.6, targets=pts, want=2)
    f_acc = u.G.si * u.Msol.si / u.kpc.si ** 2 / 1e3
    a_r = -(a_o.reshape(len(radii), n_phi, 3) * ring[None]).sum(2).mean(1) * f_acc
    ref = np.sqrt(a_r * radii * u.kpc.si / 1e3)
    assert np.allclose(np.asarray(vc), ref, rtol=1e-5)
    # and against the analytic Plummer curve v_c^2 = G M R^2 / (R^2 + a^2)^(3/2) (finite, truncated sample: several per cent at the innermost ring)
    M = m.sum()
    vc_an = np.sqrt(u.G.si * M * u.Msol.si * (radii * u.kpc.si) ** 2 / ((radii ** 2 + a ** 2) ** 1.5 * u.kpc.si ** 3)) / 1e3
    assert np.allclose(np.asarray(vc), vc_an, rtol=0.1)  # sampling noise of 4e4 particles inside 0.5 kpc: ~6 %
    assert vc.units.si == pytest.approx(1e3) and vc.sim is sim
