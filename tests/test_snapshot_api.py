"""Snapshot-level API (calculate_potential / calculate_acceleration, reference pyn_gravity.py:31-216) driven with a
minimal fake pynbody (tests/fake_pynbody) because pynbody is absent from this image: units, SimArray coercion of
positions / softening, method switch, kwargs handling and error behaviour."""
import os
import sys

import numpy as np
import pytest

FAKE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fake_pynbody")


@pytest.fixture()
def fake_pynbody():
    sys.path.insert(0, FAKE)
    for k in [k for k in sys.modules if k == "pynbody" or k.startswith("pynbody.")]:
        del sys.modules[k]
    import pynbody  # noqa: F401
    yield sys.modules["pynbody"]
    sys.path.remove(FAKE)
    for k in [k for k in sys.modules if k == "pynbody" or k.startswith("pynbody.")]:
        del sys.modules[k]


def vec_rel(a, ref):
    return np.sqrt((((a - ref) ** 2).sum(1) / (ref ** 2).sum(1)).mean())


def make_sim(pynbody, n=3000, seed=5):
    from benchmarks.synthetic import plummer
    pos, m = plummer(n, seed=seed, a=2.0)  # kpc, Msol-ish
    sim = pynbody.snapshot.SimSnap(pos=pynbody.array.SimArray(pos, pynbody.units.kpc),
                                   mass=pynbody.array.SimArray(m * 1e10, pynbody.units.Msol))
    return sim, pos, m * 1e10


def test_unknown_method_raises(fake_pynbody):
    from pynbodyext.gravity import calculate_potential
    sim, _, _ = make_sim(fake_pynbody, 50)
    with pytest.raises(ValueError, match="Unknown method: fmm"):
        calculate_potential(sim, method="fmm")


@pytest.mark.gpu
def test_units_and_methods(fake_pynbody):
    from oracle import oracle as O
    from pynbodyext.gravity import KernelKind, calculate_acceleration, calculate_potential
    u = fake_pynbody.units
    sim, pos, m = make_sim(fake_pynbody)
    # G Msol / kpc -> km^2 s^-2 and G Msol / kpc^2 -> km s^-2
    f_pot = u.G.si * u.Msol.si / u.kpc.si / 1e6
    f_acc = u.G.si * u.Msol.si / u.kpc.si ** 2 / 1e3
    p_o, a_o = O.direct(pos, m)
    pd = calculate_potential(sim, method="direct")
    ad = calculate_acceleration(sim, method="direct")
    assert pd.sim is sim and pd.units.si == pytest.approx(1e6)
    assert np.allclose(np.asarray(pd), p_o * f_pot, rtol=1e-5)
    assert vec_rel(np.asarray(ad), a_o * f_acc) < 1e-5
    # tree path: leaf_capacity / multipole_order kwargs are NOT forwarded to the tree call: results
    # equal the (8, 3) tree whatever is passed
    h = 0.05
    o = O.Tree(pos, m, 8, 3, np.full(len(m), h), 1)
    pt = calculate_potential(sim, softening=h, method="tree", kernel=KernelKind.Spline, theta=0.6, leaf_capacity=64,
                             multipole_order=0)
    assert np.allclose(np.asarray(pt), o.eval(0.6, want=1)[0] * f_pot, rtol=1e-5)
    # SimArray softening in other units and SimArray target positions are converted to the position units
    soft = fake_pynbody.array.SimArray(np.full(len(m), h * 1e-3), "Mpc")
    tg = fake_pynbody.array.SimArray(pos[:64] * 1e-3 + 1e-4, "Mpc")
    at = calculate_acceleration(sim, positions=tg, softening=soft, method="tree", kernel=KernelKind.Spline)
    a_ref = o.eval(0.7, targets=np.asarray(tg) * 1e3, want=2)[1]
    assert vec_rel(np.asarray(at), a_ref * f_acc) < 1e-5
    # softening without a kernel is an error from the backend
    with pytest.raises(ValueError, match="softenings require an explicit kernel"):
        calculate_potential(sim, softening=0.01, method="direct")


@pytest.mark.gpu
def test_property_nodes_use_the_active_view(fake_pynbody):
    # a filtered / shifted view reaches gravity only as its pos / mass arrays: the node must evaluate
    # exactly what calculate_potential gives for that view
    from pynbodyext.gravity import calculate_potential
    from pynbodyext.gravity.properties import GravityAcceleration, GravityPotential
    sim, pos, m = make_sim(fake_pynbody, n=4000, seed=9)
    keep = (pos ** 2).sum(1) < 4.0  # "Sphere(2 kpc)" filter
    view = fake_pynbody.snapshot.SimSnap(pos=fake_pynbody.array.SimArray(pos[keep] - pos[keep].mean(0), "kpc"),
                                         mass=fake_pynbody.array.SimArray(m[keep], "Msol"))
    node = GravityPotential(softening=0.05, kernel=1, theta=0.6)
    got = node(view)
    ref = calculate_potential(view, softening=0.05, kernel=1, theta=0.6)
    assert np.array_equal(np.asarray(got), np.asarray(ref)) and got.sim is view
    acc = GravityAcceleration(method="direct")(view)
    assert acc.shape == (int(keep.sum()), 3) and np.isfinite(np.asarray(acc)).all()
    assert node.instance_signature() == GravityPotential(softening=0.05, kernel=1, theta=0.6).instance_signature()
