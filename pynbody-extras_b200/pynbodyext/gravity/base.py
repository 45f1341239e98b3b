"""Array-level gravity API, signature-compatible with the reference's
``pynbodyext/gravity/base.py`` (``Gravity`` :132-454, ``KernelKind`` :71-80, ``TreeOptions`` :82-100).

Everything below the five ``pynbodyext._rust`` names runs on the GPU (libpnbx_gravity.so).
Behaviour kept from the reference, including its quirks:

* a scalar ``softening`` is broadcast to a per-particle array (base.py:192-193);
* ``tree_*`` methods take their own ``leaf_capacity`` / ``multipole_order`` defaults (8, 3) and
  build a throw-away tree whenever they differ from the instance options (base.py:213-228;
  SURVEY F11) — only the instance-option tree is cached;
* ``KernelKind.No`` with a softening is an error raised by the backend.

One deliberate difference: inputs that are already float64 C-contiguous are not copied in
``__init__`` (the reference ``astype`` copies, base.py:199-200); the device upload is the copy.

Additive keyword (not in the reference): ``precision=`` on ``Gravity(...)`` and on every method
(``"f32"`` | ``"f64"`` | ``None`` = auto, see ``pynbodyext._rust``).
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum

import numpy as np

try:
    from pynbodyext._rust import (
        Octree as _Octree,
        direct_accelerations_at_points_py as _direct_accelerations_at_points_py,
        direct_accelerations_py as _direct_accelerations_py,
        direct_potentials_at_points_py as _direct_potentials_at_points_py,
        direct_potentials_py as _direct_potentials_py,
    )
except ImportError as exc:  # pragma: no cover
    raise ImportError(
        "pynbodyext.gravity requires the backend module `pynbodyext._rust` "
        "(ctypes shim over libpnbx_gravity.so); build it with __graft_entry__.build()."
    ) from exc

from pynbodyext.log import logger

__all__ = ["Gravity", "KernelKind", "TreeOptions"]


class KernelKind(Enum):
    """Softening kernel selector; ``.value`` is the backend code (None / 0 / 1)."""

    No = None
    Plummer = 0
    Spline = 1


@dataclass(eq=True, frozen=True)
class TreeOptions:
    """Octree construction options (hashable; equality decides tree-cache hits)."""

    leaf_capacity: int = 8
    multipole_order: int = 3
    kernel: KernelKind = KernelKind.No


def _build_tree(positions, masses, softening, options: TreeOptions, precision=None):
    return _Octree(
        positions,
        masses,
        options.leaf_capacity,
        options.multipole_order,
        softening,
        options.kernel.value,
        precision=precision,
    )


def _targets(positions):
    pos = np.asarray(positions, dtype=np.float64)
    assert pos.ndim == 2 and pos.shape[1] == 3, "positions must be of shape (N, 3)"
    return pos


class Gravity:
    """Direct-summation and tree gravity for a fixed particle set.

    Parameters mirror the reference: ``positions`` (N,3), ``masses`` (N,), ``softening``
    (None | float | (N,) array), ``kernel`` (:class:`KernelKind`), ``leaf_capacity``,
    ``multipole_order``.
    """

    def __init__(self, positions, masses, softening=None, kernel=KernelKind.No, leaf_capacity=8, multipole_order=3, *,
                 precision=None):
        pos, mass = map(np.asarray, (positions, masses))
        if pos.ndim != 2 or pos.shape[1] != 3:
            raise ValueError("positions must be a float64 array of shape (N, 3)")
        if mass.shape != (pos.shape[0],):
            raise ValueError("masses must be a float64 array of shape (N,)")
        if softening is None:
            soft_arr = None
        elif np.isscalar(softening):
            soft_arr = np.full((pos.shape[0],), float(softening), dtype=np.float64)
        else:
            soft_arr = np.ascontiguousarray(softening, dtype=np.float64)
            if soft_arr.shape != (pos.shape[0],):
                raise ValueError("softening must be a float64 array of shape (N,)")

        self.pos = np.ascontiguousarray(pos, dtype=np.float64)
        self.mass = np.ascontiguousarray(mass, dtype=np.float64)
        self.softening = soft_arr
        self.tree_options = TreeOptions(leaf_capacity, multipole_order, kernel=KernelKind(kernel))
        self.precision = precision
        self._tree = None  # built lazily

    # ------------------------------------------------------------------ tree cache
    def get_tree(self, leaf_capacity=8, multipole_order=3, kernel=KernelKind.No):
        """Cached tree if the options equal the instance options, else a fresh (uncached) one."""
        options = TreeOptions(leaf_capacity, multipole_order, kernel=KernelKind(kernel))
        if options == self.tree_options:
            return self.tree
        logger.debug("Building new Octree with leaf_capacity=%d, multipole_order=%d", leaf_capacity, multipole_order)
        return _build_tree(self.pos, self.mass, self.softening, options, self.precision)

    @property
    def tree(self):
        if self._tree is None:
            self._tree = _build_tree(self.pos, self.mass, self.softening, self.tree_options, self.precision)
        return self._tree

    def _prec(self, precision):
        return self.precision if precision is None else precision

    def _kernel(self, kernel):
        return self.tree_options.kernel if kernel is None else KernelKind(kernel)

    # ------------------------------------------------------------------ direct summation
    def direct_potentials(self, positions=None, threads=0, kernel=None, *, precision=None):
        k = self._kernel(kernel)
        p = self._prec(precision)
        if positions is None:
            return _direct_potentials_py(self.pos, self.mass, threads, self.softening, k.value, precision=p)
        return _direct_potentials_at_points_py(self.pos, _targets(positions), self.mass, threads, self.softening, k.value,
                                               precision=p)

    def direct_accelerations(self, positions=None, threads=0, kernel=None, *, precision=None):
        k = self._kernel(kernel)
        p = self._prec(precision)
        if positions is None:
            return _direct_accelerations_py(self.pos, self.mass, threads, self.softening, k.value, precision=p)
        return _direct_accelerations_at_points_py(self.pos, _targets(positions), self.mass, threads, self.softening,
                                                  k.value, precision=p)

    # ------------------------------------------------------------------ tree
    def tree_potentials(self, positions=None, theta=0.7, threads=0, leaf_capacity=8, multipole_order=3, kernel=None, *,
                        precision=None):
        tree = self.get_tree(leaf_capacity=leaf_capacity, multipole_order=multipole_order, kernel=self._kernel(kernel))
        p = self._prec(precision)
        if positions is None:
            return tree.compute_potentials(theta, threads, precision=p)
        return tree.potentials_at_points(_targets(positions), theta, threads, precision=p)

    def tree_accelerations(self, positions=None, theta=0.7, threads=0, leaf_capacity=8, multipole_order=3, kernel=None, *,
                           precision=None):
        tree = self.get_tree(leaf_capacity=leaf_capacity, multipole_order=multipole_order, kernel=self._kernel(kernel))
        p = self._prec(precision)
        if positions is None:
            return tree.compute_accelerations(theta, threads, precision=p)
        return tree.accelerations_at_points(_targets(positions), theta, threads, precision=p)
