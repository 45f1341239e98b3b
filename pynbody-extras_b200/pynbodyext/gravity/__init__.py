"""pynbodyext.gravity — same public names as the reference package
(pynbodyext/gravity/__init__.py:15-30): ``Gravity``, ``KernelKind``,
``calculate_potential``, ``calculate_acceleration``, ``GRAVITY_RUST_AVAILABLE``.

The snapshot-level functions import pynbody lazily, so the array-level ``Gravity`` API works
without pynbody.
"""
from pynbodyext.util.deps import GRAVITY_RUST_AVAILABLE

__all__ = ["GRAVITY_RUST_AVAILABLE"]

if GRAVITY_RUST_AVAILABLE:
    from .base import Gravity, KernelKind, TreeOptions
    from .pyn_gravity import calculate_acceleration, calculate_potential

    __all__ += ["Gravity", "KernelKind", "TreeOptions", "calculate_potential", "calculate_acceleration"]
else:  # pragma: no cover
    import warnings

    warnings.warn(
        "pynbodyext.gravity: backend module pynbodyext._rust not importable; gravity is unavailable.",
        ImportWarning,
        stacklevel=2,
    )
