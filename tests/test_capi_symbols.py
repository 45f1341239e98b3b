"""The C-ABI library loads and exports every symbol include/pnbx_gravity.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "pnbx_gravity.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pnbx_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_expected_entry_points():
    names = _declared()
    for must in ("pnbx_direct", "pnbx_tree_create", "pnbx_tree_eval", "pnbx_tree_destroy", "pnbx_last_error"):
        assert must in names


def test_library_exports_all_declared_symbols():
    import pynbodyext._rust as backend
    lib = backend._load()
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.pnbx_abi_version() == 1
    assert lib.pnbx_device_count() >= 0


def test_product_does_not_reference_oracle():
    # the product path must never route through the oracle (parity claims depend on it)
    pkg = os.path.join(ROOT, "pynbody-extras_b200")
    for d, _, files in os.walk(pkg):
        if os.sep + "build" in d:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "pnbx_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, (d, f)


def test_argument_errors_match_reference_messages():
    import numpy as np
    import pynbodyext._rust as r

    pos = np.zeros((4, 3))
    with pytest.raises(ValueError, match="masses must be length N"):
        r.direct_potentials_py(pos, np.ones(3))
    with pytest.raises(ValueError, match="softenings must be length N"):
        r.direct_potentials_py(pos, np.ones(4), 0, np.ones(5), 0)
    with pytest.raises(ValueError, match="softenings require an explicit kernel"):
        r.direct_accelerations_py(pos, np.ones(4), 0, np.ones(4), None)
    with pytest.raises(ValueError, match=r"kernel must be 0 \(Plummer\) or 1 \(CubicSplineW2\)"):
        r.direct_accelerations_py(pos, np.ones(4), 0, None, 2)
    with pytest.raises(ValueError, match=r"targets must be \(N,3\) float64 array"):
        r.direct_potentials_at_points_py(pos, np.zeros((4, 2)))
    with pytest.raises(TypeError):
        r.direct_potentials_py(pos.astype(np.float32))


def test_python_api_surface():
    import inspect

    from pynbodyext.gravity import Gravity, KernelKind, calculate_acceleration, calculate_potential
    assert [k.name for k in KernelKind] == ["No", "Plummer", "Spline"] and KernelKind.Spline.value == 1
    sig = inspect.signature(Gravity.tree_potentials)
    assert list(sig.parameters) == ["self", "positions", "theta", "threads", "leaf_capacity", "multipole_order", "kernel"]
    assert sig.parameters["theta"].default == 0.7 and sig.parameters["multipole_order"].default == 3
    sig = inspect.signature(calculate_potential)
    assert list(sig.parameters)[:5] == ["sim", "positions", "softening", "method", "threads"]
    assert sig.parameters["method"].default == "tree"
    import pynbodyext._rust as r
    sig = inspect.signature(r.Octree.__init__)
    assert [sig.parameters[p].default for p in ("masses", "leaf_capacity", "multipole_order", "softenings", "kernel")] == [None, 32, 0, None, None]
    assert calculate_acceleration is not None


def test_compute_fails_loudly_without_a_gpu():
    # no CPU fallback: on a machine without a CUDA device every compute entry point must raise, not return numbers
    import numpy as np
    import pynbodyext._rust as r

    if r._load().pnbx_device_count() > 0:
        pytest.skip("a CUDA device is present")
    pos = np.random.default_rng(0).random((16, 3))
    with pytest.raises(RuntimeError, match="no CUDA device available"):
        r.direct_potentials_py(pos)
    with pytest.raises(RuntimeError, match="no CUDA device available"):
        r.Octree(pos, np.ones(16))
    from pynbodyext.gravity import Gravity
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Gravity(pos, np.ones(16)).tree_potentials()
