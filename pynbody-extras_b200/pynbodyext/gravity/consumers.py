"""First in-package consumers of the gravity path.

In the reference the potential-based centre reads a potential that must already be in the snapshot
(``CenPos(mode="pot")``: ``i = sim["phi"].argmin()``, properties/generic.py:51-52) and ``ShiftPosTo("pot")``
(transforms/shift.py:17-24) translates to it; nothing in the package computes ``phi``. The nodes below compute it on
the GPU through :func:`calculate_potential` / :func:`calculate_acceleration` (same keyword arguments: ``softening``,
``method``, ``kernel``, ``theta`` ...), for the ACTIVE view they are handed (filters and transforms of a calculator
chain select the sources):

* :class:`PotentialCenter` — position of the particle at the potential minimum (the ``mode="pot"`` centre);
* :func:`shift_to_potential_minimum` — the ``ShiftPosTo("pot")`` translation applied to ``sim["pos"]``;
* :func:`binding_energy` — specific binding energy ``0.5 |v - v_cen|^2 + phi`` per particle;
* :func:`rotation_curve` — ``v_c(R) = sqrt(R |a_R|)`` from accelerations evaluated at ring points (the at-points path).

They derive from the reference's ``PropertyBase`` when that framework is importable and are plain callables otherwise
(pynbody and the calculator framework are absent from this image: the integration is exercised with the fake snapshot
of tests/fake_pynbody only).
"""
from __future__ import annotations

from typing import Any

import numpy as np

from .base import KernelKind
from .pyn_gravity import _pynbody, calculate_acceleration, calculate_potential

try:  # inside pynbody-extras: a real calculator node
    from pynbodyext.calculate import PropertyBase as _Base  # type: ignore
    _HAVE_FRAMEWORK = True
except Exception:  # stand-alone
    _HAVE_FRAMEWORK = False

    class _Base:  # type: ignore
        def __call__(self, sim, *args: Any, **kwargs: Any):
            return self.calculate(sim)

__all__ = ["PotentialCenter", "shift_to_potential_minimum", "binding_energy", "rotation_curve"]


class PotentialCenter(_Base):
    """``CenPos(mode="pot")`` with the potential computed here instead of read from ``sim["phi"]``.

    ``use_existing=True`` keeps the reference behaviour when the snapshot already carries ``phi``.
    """

    def __init__(self, softening=None, method="tree", threads=0, kernel=KernelKind.No, use_existing=False, **kwargs: Any):
        if _HAVE_FRAMEWORK:
            super().__init__()
        self.softening = softening
        self.method = method
        self.threads = threads
        self.kernel = KernelKind(kernel)
        self.use_existing = use_existing
        self.kwargs = dict(kwargs)

    def instance_signature(self):
        soft = self.softening if self.softening is None or isinstance(self.softening, (int, float)) else id(self.softening)
        return (type(self).__name__, soft, self.method, self.kernel.name, self.use_existing, tuple(sorted(self.kwargs.items())))

    def potential(self, sim):
        if self.use_existing:
            try:
                return sim["phi"]
            except KeyError:
                pass
        return calculate_potential(sim, None, self.softening, self.method, self.threads, kernel=self.kernel, **self.kwargs)

    def calculate(self, sim, params: Any = None):
        i = int(np.asarray(self.potential(sim)).argmin())  # properties/generic.py:51
        cen = sim["pos"][i].copy()                          # properties/generic.py:52
        if hasattr(cen, "sim"):
            cen.sim = sim
        return cen


def shift_to_potential_minimum(sim, **kwargs: Any):
    """``ShiftPosTo("pot")``: translate ``sim["pos"]`` so that the potential minimum sits at the origin.
    Returns the centre that was subtracted (position units). Keyword arguments as :class:`PotentialCenter`."""
    cen = PotentialCenter(**kwargs)(sim)
    sim["pos"][...] = np.asarray(sim["pos"]) - np.asarray(cen)
    return cen


def binding_energy(sim, vcen=None, **kwargs: Any):
    """Specific binding energy per particle, ``0.5 |v - v_cen|^2 + phi`` in km^2 s^-2.

    ``sim["vel"]`` is converted to km/s; ``vcen`` defaults to the mass-weighted mean velocity. Keyword arguments
    (softening, method, kernel, theta, ...) go to :func:`calculate_potential`."""
    _units, SimArray = _pynbody()
    phi = calculate_potential(sim, None, kwargs.pop("softening", None), kwargs.pop("method", "tree"),
                              kwargs.pop("threads", 0), kernel=kwargs.pop("kernel", KernelKind.No), **kwargs)
    vel = sim["vel"]
    vel = np.asarray(vel.in_units("km s**-1")) if hasattr(vel, "in_units") else np.asarray(vel, dtype=np.float64)
    mass = np.asarray(sim["mass"], dtype=np.float64)
    if vcen is None:
        vcen = (vel * mass[:, None]).sum(0) / mass.sum()
    e = 0.5 * ((vel - np.asarray(vcen)) ** 2).sum(1) + np.asarray(phi)
    out = SimArray(e, "km**2 s**-2")
    out.sim = sim
    return out


def rotation_curve(sim, radii, n_phi=16, **kwargs: Any):
    """Circular velocity ``v_c(R) = sqrt(R |a_R|)`` in km/s at cylindrical radii ``radii`` (position units, or a
    SimArray) in the z = 0 plane: the radial acceleration is averaged over ``n_phi`` points per ring and evaluated at
    points by :func:`calculate_acceleration` (direct sum or tree, same keyword arguments)."""
    _units, SimArray = _pynbody()
    pos_units = sim["pos"].units
    r = radii.in_units(pos_units) if isinstance(radii, SimArray) else radii
    r = np.asarray(r, dtype=np.float64).ravel()
    ang = 2.0 * np.pi * (np.arange(n_phi) + 0.5) / n_phi
    ring = np.stack([np.cos(ang), np.sin(ang), np.zeros(n_phi)], axis=1)     # (n_phi, 3) unit vectors
    pts = (r[:, None, None] * ring[None, :, :]).reshape(-1, 3)
    acc = calculate_acceleration(sim, pts, kwargs.pop("softening", None), kwargs.pop("method", "tree"),
                                 kwargs.pop("threads", 0), kernel=kwargs.pop("kernel", KernelKind.No), **kwargs)
    a_r = -(np.asarray(acc).reshape(len(r), n_phi, 3) * ring[None, :, :]).sum(2).mean(1)  # inward pull, km s^-2
    r_km = r * (pos_units.ratio(_units.km) if hasattr(pos_units, "ratio") else float(pos_units.in_units("km")))
    vc = np.sqrt(np.maximum(a_r, 0.0) * r_km)
    out = SimArray(vc, "km s**-1")
    out.sim = sim
    return out
