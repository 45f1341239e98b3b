"""Calculator-graph entry for the gravity path (the reference has the hook points —
``PropertyBase.calculate(sim, params)``, core/calculate/properties.py:125-145 — but no gravity node).

``GravityPotential`` / ``GravityAcceleration`` are property nodes whose ``calculate`` hands the ACTIVE view it is given
(``sim[filter]`` after the transforms of the calculator chain, core/calculate/base.py:980-1008) to
``calculate_potential`` / ``calculate_acceleration``: the filters and transforms select the sources, the GPU path
does the sums. When this package is dropped into pynbody-extras the classes derive from its ``PropertyBase`` (so
``.filter(Sphere(...)).transform(ShiftPosTo(...))`` chains work); stand-alone they are plain callables.
The integration with the real framework is untested here (pynbody and the framework are absent from this image).
"""
from __future__ import annotations

from typing import Any

from .base import KernelKind
from .pyn_gravity import calculate_acceleration, calculate_potential

try:  # inside pynbody-extras: real calculator nodes
    from pynbodyext.calculate import PropertyBase as _Base  # type: ignore
    _HAVE_FRAMEWORK = True
except Exception:  # stand-alone: minimal callable base
    _HAVE_FRAMEWORK = False

    class _Base:  # type: ignore
        def __call__(self, sim, *args: Any, **kwargs: Any):
            return self.calculate(sim)

__all__ = ["GravityPotential", "GravityAcceleration"]


class _GravityNode(_Base):
    _want_acc = False

    def __init__(self, positions=None, softening=None, method="tree", threads=0, kernel=KernelKind.No, **kwargs: Any):
        if _HAVE_FRAMEWORK:
            super().__init__()
        self.positions = positions
        self.softening = softening
        self.method = method
        self.threads = threads
        self.kernel = KernelKind(kernel)
        self.kwargs = dict(kwargs)  # theta, leaf_capacity, multipole_order

    def instance_signature(self):
        pos_key = None if self.positions is None else id(self.positions)
        soft_key = self.softening if self.softening is None or isinstance(self.softening, (int, float)) else id(self.softening)
        return (type(self).__name__, pos_key, soft_key, self.method, self.kernel.name, tuple(sorted(self.kwargs.items())))

    def calculate(self, sim, params: Any = None):
        fn = calculate_acceleration if self._want_acc else calculate_potential
        return fn(sim, self.positions, self.softening, self.method, self.threads, kernel=self.kernel, **self.kwargs)


class GravityPotential(_GravityNode):
    """Potential of the active snapshot view at its particles (or at ``positions``), km^2 s^-2."""
    _want_acc = False


class GravityAcceleration(_GravityNode):
    """Acceleration of the active snapshot view at its particles (or at ``positions``), km s^-2."""
    _want_acc = True
