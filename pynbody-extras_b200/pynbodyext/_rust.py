"""Drop-in replacement for the PyO3 module ``pynbodyext._rust`` (reference:
crates/pynbodyext-rust/src/lib.rs:10-27, gravity.rs).

Exports the same five names with the same signatures, defaults and error strings —
``Octree`` and the four ``direct_*_py`` functions — but backed by ``libpnbx_gravity.so``
(hand-written sm_100a CUDA, C-ABI in include/pnbx_gravity.h) through ctypes. ctypes releases
the GIL for the duration of every call, like ``py.allow_threads`` in gravity.rs:103-111.

There is no CPU fallback: if the library or a CUDA device is missing the calls raise.
``threads`` is accepted for signature compatibility and ignored (it sized the rayon pool).

Additive, keyword-only extensions (not in the reference): ``precision=`` and ``device=`` on every call; the
device-resident variants live in ``pynbodyext.gravity.device``.

Precision. The reference is float64 throughout; BASELINE's north_star asks for fp32 interaction arithmetic
(accumulated in float64, RMS error ~1e-7 at the tested softenings). ``precision`` selects it per call:
``"f32"``, ``"f64"`` (every pair in float64: the validation mode), or ``None`` / ``"auto"`` = the value of
``$PNBX_PRECISION`` if set, else fp32 — except for UNSOFTENED direct sums of at most 2^30 pairs (the reference's
documented small-N validation use of method="direct"), which run in float64: without a softening length a close pair's
separation can be below the fp32 resolution of box-centred coordinates (6e-8 of the box).

Devices. ``device=None`` is the current CUDA device; with ``$PNBX_DEVICES=all`` (or "0,1,...") host-array calls
without an explicit device are spread over those GPUs (csrc/multi.cu), results identical.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

__all__ = [
    "Octree",
    "direct_accelerations_py",
    "direct_potentials_py",
    "direct_accelerations_at_points_py",
    "direct_potentials_at_points_py",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_CANDIDATES = [
    os.environ.get("PNBX_GRAVITY_LIB", ""),
    os.path.join(_HERE, "..", "lib", "libpnbx_gravity.so"),
]

PNBX_OK, PNBX_ERR_ARG, PNBX_ERR_CUDA, PNBX_ERR_STATE, PNBX_ERR_DEPTH = 0, 1, 2, 3, 4
WANT_POT, WANT_ACC = 1, 2
MEM_HOST, MEM_DEVICE = 0, 1
PREC_F32, PREC_F64 = 0, 1


class pnbx_opts(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("mem_space", C.c_int32),
        ("precision", C.c_int32),
        ("flags", C.c_int32),
        ("stream", C.c_void_p),
        ("shard_rank", C.c_int32),
        ("shard_world", C.c_int32),
        ("shard_block", C.c_int64),
    ]


class pnbx_tree_info(C.Structure):
    _fields_ = [
        ("n_particles", C.c_int64),
        ("n_nodes", C.c_int64),
        ("n_leaves", C.c_int64),
        ("depth", C.c_int32),
        ("multipole_order", C.c_int32),
        ("n_moments", C.c_int32),
        ("has_payload", C.c_int32),
        ("has_hmax", C.c_int32),
        ("kernel", C.c_int32),
        ("leaf_capacity", C.c_int64),
    ]


_vp = C.c_void_p
_lib = None


def _load():
    """Load libpnbx_gravity.so once; fail loudly if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = next((os.path.abspath(p) for p in _LIB_CANDIDATES if p and os.path.exists(p)), None)
    if path is None:
        raise ImportError(
            "libpnbx_gravity.so not found (expected pynbody-extras_b200/lib/libpnbx_gravity.so); "
            "build it with `python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback."
        )
    L = C.CDLL(path)
    L.pnbx_last_error.restype = C.c_char_p
    L.pnbx_abi_version.restype = C.c_int
    L.pnbx_device_count.restype = C.c_int
    L.pnbx_direct.argtypes = [_vp, _vp, _vp, C.c_int64, _vp, C.c_int64, C.c_int64, C.c_int, C.c_int, _vp, _vp,
                              C.POINTER(pnbx_opts)]
    if hasattr(L, "pnbx_tree_create"):
        L.pnbx_tree_create.argtypes = [C.POINTER(_vp), _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                       C.POINTER(pnbx_opts)]
        L.pnbx_tree_build_mass.argtypes = [_vp, _vp]
        L.pnbx_tree_set_softenings.argtypes = [_vp, _vp]
        L.pnbx_tree_get_order.argtypes = [_vp, C.c_int64, C.c_int64, _vp, C.POINTER(pnbx_opts)]
        if hasattr(L, "pnbx_tree_build_mass_ex"):  # ABI additions of this round (older builds are used for A/B timing)
            L.pnbx_tree_build_mass_ex.argtypes = [_vp, _vp, C.POINTER(pnbx_opts)]
            L.pnbx_tree_set_softenings_ex.argtypes = [_vp, _vp, C.POINTER(pnbx_opts)]
            L.pnbx_trim_memory.restype = None
        L.pnbx_tree_set_kernel.argtypes = [_vp, C.c_int]
        L.pnbx_tree_eval.argtypes = [_vp, _vp, C.c_int64, C.c_int64, C.c_double, C.c_int, _vp, _vp,
                                     C.POINTER(pnbx_opts)]
        L.pnbx_tree_destroy.argtypes = [_vp]
        L.pnbx_tree_destroy.restype = None
        L.pnbx_tree_get_info.argtypes = [_vp, C.POINTER(pnbx_tree_info)]
        L.pnbx_tree_dump_topology.argtypes = [_vp] + [_vp] * 10
        L.pnbx_tree_dump_payload.argtypes = [_vp] + [_vp] * 4
        L.pnbx_tree_dump_keys.argtypes = [_vp, _vp, _vp]
        L.pnbx_tree_walk_counters.argtypes = [_vp, _vp, C.c_int64, C.c_int64, C.c_double, _vp, C.POINTER(pnbx_opts)]
    L.pnbx_last_timings.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.c_int]
    _lib = L
    return L


def library_path() -> str:
    _load()
    return _lib._name


def last_timings() -> dict:
    """(label -> ms) of the last call on this thread when GRAVITY_TIMING is set."""
    L = _load()
    labels = (C.c_char_p * 32)()
    ms = (C.c_double * 32)()
    k = L.pnbx_last_timings(labels, ms, 32)
    return {labels[i].decode(): ms[i] for i in range(k)}


def _check(rc: int) -> None:
    if rc == PNBX_OK:
        return
    msg = _load().pnbx_last_error().decode()
    if rc in (PNBX_ERR_ARG, PNBX_ERR_STATE, PNBX_ERR_DEPTH):
        raise ValueError(msg)
    raise RuntimeError(msg)


AUTO_F64_MAX_PAIRS = 1 << 30


def resolve_precision(precision, unsoftened_pairs=None):
    """'f32' | 'f64' for a call (module docstring). unsoftened_pairs: N*M of a direct sum without softening."""
    if precision in (None, "auto"):
        env = os.environ.get("PNBX_PRECISION", "").strip().lower()
        if env in ("f32", "f64"):
            return env
        if env not in ("", "auto"):
            raise ValueError("PNBX_PRECISION must be 'f32', 'f64' or 'auto'")
        return "f64" if (unsoftened_pairs is not None and unsoftened_pairs <= AUTO_F64_MAX_PAIRS) else "f32"
    if precision in ("f32", PREC_F32):
        return "f32"
    if precision in ("f64", PREC_F64):
        return "f64"
    raise ValueError("precision must be 'f32', 'f64' or 'auto'")


def _opts(device=None, precision=None, mem_space=MEM_HOST, stream=None, unsoftened_pairs=None):
    o = pnbx_opts()
    o.device = -1 if device is None else int(device)
    o.mem_space = mem_space
    o.precision = PREC_F64 if resolve_precision(precision, unsoftened_pairs) == "f64" else PREC_F32
    o.flags = 0
    o.stream = stream
    o.shard_rank, o.shard_world, o.shard_block = 0, 1, 0
    return o


# ---- argument extraction with the binding's checks and messages (gravity.rs:33-65, 154-189) ----
def _vec3(arr, name):
    a = np.asarray(arr)
    if a.dtype != np.float64:
        raise TypeError(f"{name} must be a float64 array (got {a.dtype})")
    if a.ndim != 2 or a.shape[1] != 3:
        # The reference reinterprets any contiguous buffer whose length is divisible by 3
        # (SURVEY F15); we require the documented (N,3) shape instead.
        raise ValueError(f"{name} must be (N,3) float64 array")
    return np.ascontiguousarray(a)


def _vec1(arr, n, what):
    if arr is None:
        return None
    a = np.asarray(arr)
    if a.dtype != np.float64:
        raise TypeError(f"{what} must be a float64 array (got {a.dtype})")
    if a.ndim != 1 or not a.flags.c_contiguous:
        raise ValueError(f"{what} must be a contiguous 1-D float64 array")
    if a.shape[0] != n:
        raise ValueError(f"{what} must be length N")
    return a


def _kernel_code(kernel, none_value):
    if kernel is None:
        return none_value
    k = int(kernel)
    if k not in (0, 1):
        raise ValueError("kernel must be 0 (Plummer) or 1 (CubicSplineW2)")
    return k


def _ptr(a):
    return None if a is None else a.ctypes.data


def _direct(positions, targets, masses, softenings, kernel, want, device, precision, tname="targets"):
    pos = _vec3(positions, "positions")
    n = pos.shape[0]
    m_arr = _vec1(masses, n, "masses")
    h_arr = _vec1(softenings, n, "softenings")
    tgt = None if targets is None else _vec3(targets, tname)
    if kernel is None and h_arr is not None:
        raise ValueError("softenings require an explicit kernel; pass kernel=0/1 (or omit softenings)")
    k = _kernel_code(kernel, -1)
    m = n if tgt is None else tgt.shape[0]
    pot = np.empty(m, dtype=np.float64) if want & WANT_POT else None
    acc = np.empty((m, 3), dtype=np.float64) if want & WANT_ACC else None
    o = _opts(device, precision, unsoftened_pairs=n * m if k == -1 else None)
    rc = _load().pnbx_direct(_ptr(pos), _ptr(m_arr), _ptr(h_arr), n, _ptr(tgt), m, 0, k, want, _ptr(pot), _ptr(acc),
                             C.byref(o))
    _check(rc)
    return pot, acc


def direct_accelerations_py(positions, masses=None, threads=0, softenings=None, kernel=None, *, device=None,
                            precision=None):
    """gravity.rs:448-512 -> direct.rs:115-185 / 443-524. Returns (N,3) float64."""
    return _direct(positions, None, masses, softenings, kernel, WANT_ACC, device, precision)[1]


def direct_potentials_py(positions, masses=None, threads=0, softenings=None, kernel=None, *, device=None,
                         precision=None):
    """gravity.rs:585-644 -> direct.rs:255-313 / 370-441. Returns (N,) float64."""
    return _direct(positions, None, masses, softenings, kernel, WANT_POT, device, precision)[0]


def direct_accelerations_at_points_py(positions, targets, masses=None, threads=0, softenings=None, kernel=None, *,
                                      device=None, precision=None):
    """gravity.rs:514-582 -> direct.rs:187-251 / 587-658. Returns (M,3) float64."""
    return _direct(positions, targets, masses, softenings, kernel, WANT_ACC, device, precision)[1]


def direct_potentials_at_points_py(positions, targets, masses=None, threads=0, softenings=None, kernel=None, *,
                                   device=None, precision=None):
    """gravity.rs:646-709 -> direct.rs:315-368 / 526-585. Returns (M,) float64."""
    return _direct(positions, targets, masses, softenings, kernel, WANT_POT, device, precision)[0]


class Octree:
    """GPU octree with the interface of the reference pyclass (gravity.rs:113-445).

    ``Octree(positions, masses=None, leaf_capacity=32, multipole_order=0, softenings=None, kernel=None)``
    builds the topology (and the mass / hmax / multipole payloads iff ``masses`` is given) on the
    device; the object owns device copies of the sources for its lifetime.
    """

    def __init__(self, positions, masses=None, leaf_capacity=32, multipole_order=0, softenings=None, kernel=None, *,
                 device=None, precision=None):
        pos = _vec3(positions, "positions")
        n = pos.shape[0]
        m_arr = _vec1(masses, n, "masses")
        h_arr = _vec1(softenings, n, "softenings")
        if kernel is None and h_arr is not None:
            raise ValueError("softenings require an explicit kernel; pass kernel=0/1 (or omit softenings)")
        k = _kernel_code(kernel, 0)  # None -> Plummer (gravity.rs:77-82)
        lc = int(leaf_capacity)
        mo = int(multipole_order)
        if lc < 0:
            raise OverflowError("can't convert negative int to unsigned")
        if not 0 <= mo <= 255:
            raise OverflowError("multipole_order out of range for u8")
        self._n = n
        self._device = device
        self._precision = precision
        self._h = _vp()
        o = _opts(device, precision)
        _check(_load().pnbx_tree_create(C.byref(self._h), _ptr(pos), _ptr(m_arr), _ptr(h_arr), n, lc, mo, k,
                                        C.byref(o)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and _lib is not None:
            _lib.pnbx_tree_destroy(h)
            self._h = _vp()

    # -- setters (gravity.rs:228-265)
    def build_mass(self, masses=None):
        m_arr = _vec1(masses, self._n, "masses")
        _check(_load().pnbx_tree_build_mass(self._h, _ptr(m_arr)))

    def set_softenings(self, softenings=None):
        h_arr = _vec1(softenings, self._n, "softenings")
        _check(_load().pnbx_tree_set_softenings(self._h, _ptr(h_arr)))

    def set_kernel(self, kernel=None):
        _check(_load().pnbx_tree_set_kernel(self._h, _kernel_code(kernel, 0)))

    # -- compute (gravity.rs:267-444)
    def _eval(self, points, theta, want, pname="points", tgt_begin=0, count=None, precision=None, method="compute"):
        if not self.info()["has_payload"]:  # gravity.rs:274-278, 327-331, 356-360, 417-421
            raise ValueError(f"mass payload not built; call build_mass() before {method}")
        tgt = None if points is None else _vec3(points, pname)
        m = (self._n if count is None else int(count)) if tgt is None else tgt.shape[0]
        pot = np.empty(m, dtype=np.float64) if want & WANT_POT else None
        acc = np.empty((m, 3), dtype=np.float64) if want & WANT_ACC else None
        o = _opts(self._device, self._precision if precision is None else precision)
        _check(_load().pnbx_tree_eval(self._h, _ptr(tgt), m, int(tgt_begin), float(theta), want, _ptr(pot), _ptr(acc),
                                      C.byref(o)))
        return pot, acc

    def compute_accelerations(self, theta, threads=0, *, precision=None):
        return self._eval(None, theta, WANT_ACC, method="compute_accelerations", precision=precision)[1]

    def compute_potentials(self, theta, threads=0, *, precision=None):
        return self._eval(None, theta, WANT_POT, method="compute_potentials", precision=precision)[0]

    def accelerations_at_points(self, points, theta, threads=0, *, precision=None):
        return self._eval(points, theta, WANT_ACC, method="accelerations_at_points", precision=precision)[1]

    def potentials_at_points(self, points, theta, threads=0, *, precision=None):
        return self._eval(points, theta, WANT_POT, method="potentials_at_points", precision=precision)[0]

    def walk_counters(self, theta, points=None, tgt_begin=0, count=None) -> dict:
        """Totals over the targets of node visits / accepts / leaf visits / leaf particles of the walk
        (decisions-only pass; equals the oracle's counters). Not in the reference."""
        tgt = None if points is None else _vec3(points, "points")
        m = (self._n if count is None else int(count)) if tgt is None else tgt.shape[0]
        out = np.zeros(5, dtype=np.int64)
        o = _opts(self._device, None)
        _check(_load().pnbx_tree_walk_counters(self._h, _ptr(tgt), m, int(tgt_begin), float(theta), _ptr(out), C.byref(o)))
        d = dict(zip(("visits", "accepts", "leaf_visits", "leaf_particles", "warp_visits"), out.tolist()))
        return d

    # -- introspection used by the parity tests (not in the reference)
    def info(self) -> dict:
        inf = pnbx_tree_info()
        _check(_load().pnbx_tree_get_info(self._h, C.byref(inf)))
        return {f: getattr(inf, f) for f, _ in pnbx_tree_info._fields_}

    def topology(self) -> dict:
        inf = self.info()
        nn, n = inf["n_nodes"], inf["n_particles"]
        out = dict(
            center=np.empty((nn, 3)), half=np.empty(nn), depth=np.empty(nn, np.int32),
            first_subnode=np.empty(nn, np.int64), next_branch=np.empty(nn, np.int64),
            leaf_start=np.empty(nn, np.int64), leaf_count=np.empty(nn, np.int64),
            leaf_particles=np.empty(n, np.int64), path_hi=np.empty(nn, np.uint64), path_lo=np.empty(nn, np.uint64),
        )
        order = ["center", "half", "depth", "first_subnode", "next_branch", "leaf_start", "leaf_count",
                 "leaf_particles", "path_hi", "path_lo"]
        _check(_load().pnbx_tree_dump_topology(self._h, *[_ptr(out[k]) for k in order]))
        return out

    def payload(self) -> dict:
        inf = self.info()
        nn, k = inf["n_nodes"], inf["n_moments"]
        mass = np.empty(nn)
        com = np.empty((nn, 3))
        hmax = np.empty(nn) if inf["has_hmax"] else None
        mom = np.zeros((nn, max(k, 1)))
        _check(_load().pnbx_tree_dump_payload(self._h, _ptr(mass), _ptr(com), _ptr(hmax), _ptr(mom)))
        return dict(mass=mass, com=com, hmax=hmax, moments=mom)

    def keys(self):
        hi = np.empty(self._n, np.uint64)
        lo = np.empty(self._n, np.uint64)
        _check(_load().pnbx_tree_dump_keys(self._h, _ptr(hi), _ptr(lo)))
        return hi, lo
