"""Snapshot-level API, signature-compatible with the reference's
``pynbodyext/gravity/pyn_gravity.py`` (``calculate_potential`` :31-123,
``calculate_acceleration`` :125-216, ``_coerce_softening`` :14-29).

pynbody is imported inside the functions (the reference imports it at module top,
pyn_gravity.py:7-9), so importing ``pynbodyext.gravity`` works without pynbody.
``leaf_capacity`` / ``multipole_order`` kwargs reach ``Gravity.__init__`` only, exactly as in the
reference: the tree path then runs with (8, 3).
"""
from __future__ import annotations

from typing import Any

import numpy as np

from .base import Gravity, KernelKind


def _pynbody():
    try:
        import pynbody  # noqa: F401
        from pynbody import units
        from pynbody.array import SimArray
    except ImportError as exc:  # pragma: no cover
        raise ImportError("calculate_potential / calculate_acceleration need pynbody (SimSnap / SimArray)") from exc
    return units, SimArray


def _coerce_softening(sim, softening, SimArray=None):
    if softening is None:
        return None
    if SimArray is not None and isinstance(softening, SimArray):
        soft = softening.in_units(sim["pos"].units)
        arr = np.asarray(soft, dtype=np.float64)
        return float(arr) if arr.ndim == 0 else arr
    if isinstance(softening, (float, int)):
        return float(softening)
    return np.asarray(softening, dtype=np.float64)


def _run(sim, positions, softening, method, threads, kernel, kwargs, want_acc):
    units, SimArray = _pynbody()
    helper = Gravity(
        sim["pos"],
        sim["mass"],
        softening=_coerce_softening(sim, softening, SimArray),
        kernel=kernel,
        leaf_capacity=kwargs.get("leaf_capacity", 8),
        multipole_order=kwargs.get("multipole_order", 3),
        precision=kwargs.get("precision"),  # additive: "f32" | "f64" | None (auto), see pynbodyext._rust
    )
    if isinstance(positions, SimArray):
        positions = positions.in_units(sim["pos"].units)
    if method == "direct":
        out = helper.direct_accelerations(positions, threads) if want_acc else helper.direct_potentials(positions, threads)
    elif method == "tree":
        theta = kwargs.get("theta", 0.7)
        out = (helper.tree_accelerations(positions, theta, threads) if want_acc
               else helper.tree_potentials(positions, theta, threads))
    else:
        raise ValueError(f"Unknown method: {method}")
    if want_acc:
        res = SimArray(out, units.G * sim["mass"].units / sim["pos"].units ** 2)
        res.sim = sim
        return res.in_units("km s**-2")
    res = SimArray(out, units.G * sim["mass"].units / sim["pos"].units)
    res.sim = sim
    return res.in_units("km**2 s**-2")


def calculate_potential(sim, positions=None, softening=None, method="tree", threads=0, *,
                        kernel: KernelKind = KernelKind.No, **kwargs: Any):
    """Potentials of a pynbody snapshot (at its particles or at ``positions``), in km^2 s^-2."""
    return _run(sim, positions, softening, method, threads, kernel, kwargs, want_acc=False)


def calculate_acceleration(sim, positions=None, softening=None, method="tree", threads=0, *,
                           kernel: KernelKind = KernelKind.No, **kwargs: Any):
    """Accelerations of a pynbody snapshot (at its particles or at ``positions``), in km s^-2."""
    return _run(sim, positions, softening, method, threads, kernel, kwargs, want_acc=True)
