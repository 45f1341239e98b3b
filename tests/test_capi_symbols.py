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
    # the reference's positional parameters, in order; additive extensions are keyword-only
    assert list(sig.parameters)[:7] == ["self", "positions", "theta", "threads", "leaf_capacity", "multipole_order", "kernel"]
    assert all(p.kind is inspect.Parameter.KEYWORD_ONLY for p in list(sig.parameters.values())[7:])
    assert sig.parameters["theta"].default == 0.7 and sig.parameters["multipole_order"].default == 3
    sig = inspect.signature(Gravity.__init__)
    assert list(sig.parameters)[:7] == ["self", "positions", "masses", "softening", "kernel", "leaf_capacity", "multipole_order"]
    assert all(p.kind is inspect.Parameter.KEYWORD_ONLY for p in list(sig.parameters.values())[7:])
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


def test_capi_argument_validation_without_gpu():
    # argument errors are reported before any device work: status code + thread-local message through the C-ABI
    import ctypes as C

    import numpy as np
    import pynbodyext._rust as r

    lib = r._load()
    pos = np.zeros((4, 3))
    out = np.zeros(4)
    o = r._opts()
    p = lambda a: a.ctypes.data  # noqa: E731
    # bad `want`
    rc = lib.pnbx_direct(p(pos), None, None, 4, None, 4, 0, -1, 0, p(out), None, C.byref(o))
    assert rc == r.PNBX_ERR_ARG and b"want" in lib.pnbx_last_error()
    # missing output buffer
    rc = lib.pnbx_direct(p(pos), None, None, 4, None, 4, 0, -1, 1, None, None, C.byref(o))
    assert rc == r.PNBX_ERR_ARG and b"out_pot" in lib.pnbx_last_error()
    # shard outside [0, N)
    rc = lib.pnbx_direct(p(pos), None, None, 4, None, 3, 2, -1, 1, p(out), None, C.byref(o))
    assert rc == r.PNBX_ERR_ARG and b"shard" in lib.pnbx_last_error()
    # softenings without kernel / bad kernel code (gravity.rs:71-73, 480-484)
    rc = lib.pnbx_direct(p(pos), None, p(out), 4, None, 4, 0, -1, 1, p(out), None, C.byref(o))
    assert rc == r.PNBX_ERR_ARG and b"softenings require an explicit kernel" in lib.pnbx_last_error()
    rc = lib.pnbx_direct(p(pos), None, None, 4, None, 4, 0, 7, 1, p(out), None, C.byref(o))
    assert rc == r.PNBX_ERR_ARG and b"kernel must be 0 (Plummer) or 1 (CubicSplineW2)" in lib.pnbx_last_error()
    # tree: bad kernel, NULL handle
    h = C.c_void_p()
    rc = lib.pnbx_tree_create(C.byref(h), p(pos), None, None, 4, 8, 0, 5, C.byref(o))
    assert rc == r.PNBX_ERR_ARG and not h.value
    assert lib.pnbx_tree_set_kernel(None, 0) == r.PNBX_ERR_ARG
    lib.pnbx_shard_count.restype = C.c_int64
    lib.pnbx_shard_count.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_int32]
    n = 100_003
    assert sum(lib.pnbx_shard_count(n, 4096, 8, k) for k in range(8)) == n
    assert lib.pnbx_shard_count(n, 4096, 1, 0) == n and lib.pnbx_shard_count(n, 4096, 8, 9) == 0
