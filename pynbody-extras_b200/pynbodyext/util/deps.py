"""Feature flags the gravity package reads (reference: pynbodyext/util/deps.py:14-19).

Unlike the reference this module does not call ``importlib.metadata.version("pynbody")`` at
import time: the array-level API must work without pynbody installed.
"""
from importlib.util import find_spec

__all__ = ["GRAVITY_RUST_AVAILABLE", "PYNBODY_AVAILABLE", "module_available"]


def module_available(name: str) -> bool:
    """True if ``name`` is importable."""
    try:
        return find_spec(name) is not None
    except (ImportError, ValueError):
        return False


# Same name as the reference flag; here "_rust" is the ctypes shim over libpnbx_gravity.so.
GRAVITY_RUST_AVAILABLE: bool = module_available("pynbodyext._rust")
PYNBODY_AVAILABLE: bool = module_available("pynbody")
