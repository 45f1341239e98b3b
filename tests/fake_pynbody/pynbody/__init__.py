"""Minimal stand-in for pynbody (absent from this image) — just enough of units / SimArray / SimSnap for the
snapshot-level gravity API (pynbodyext/gravity/pyn_gravity.py) to run in tests. Units carry an SI scale factor."""
from . import array, snapshot, units  # noqa: F401
