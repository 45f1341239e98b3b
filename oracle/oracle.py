"""ctypes loader for the CPU oracle (oracle/gravity_oracle.cpp).

TEST INFRASTRUCTURE ONLY (parity unpinned, see the .cpp header): imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs. Never by the
product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpnbx_oracle.so")
_lib = None

_d = C.POINTER(C.c_double)
_i64 = C.POINTER(C.c_int64)
_i32 = C.POINTER(C.c_int32)
_u64 = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gravity_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpnbx_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.pnbx_oracle_last_error.restype = C.c_char_p
        L.pnbx_oracle_num_threads.restype = C.c_int
        L.pnbx_oracle_direct.argtypes = [_d, _d, _d, C.c_int64, _d, C.c_int64, C.c_int, C.c_int, _d, _d]
        L.pnbx_oracle_kernel_potential.restype = C.c_double
        L.pnbx_oracle_kernel_potential.argtypes = [C.c_int, C.c_double, C.c_double]
        L.pnbx_oracle_kernel_accel_factor.restype = C.c_double
        L.pnbx_oracle_kernel_accel_factor.argtypes = [C.c_int, C.c_double, C.c_double]
        L.pnbx_oracle_tree_create.restype = C.c_void_p
        L.pnbx_oracle_tree_create.argtypes = [_d, _d, _d, C.c_int64, C.c_int64, C.c_int, C.c_int]
        L.pnbx_oracle_tree_build_mass.argtypes = [C.c_void_p, _d]
        L.pnbx_oracle_tree_set_softenings.argtypes = [C.c_void_p, _d]
        L.pnbx_oracle_tree_set_kernel.argtypes = [C.c_void_p, C.c_int]
        L.pnbx_oracle_tree_destroy.argtypes = [C.c_void_p]
        L.pnbx_oracle_tree_destroy.restype = None
        L.pnbx_oracle_tree_eval.argtypes = [C.c_void_p, _d, C.c_int64, C.c_double, C.c_int, _d, _d, _i64]
        L.pnbx_oracle_tree_eval_range.argtypes = [C.c_void_p, _d, C.c_int64, C.c_int64, C.c_double, C.c_int, _d, _d, _i64]
        L.pnbx_oracle_tree_info.argtypes = [C.c_void_p, _i64]
        L.pnbx_oracle_tree_dump_topology.argtypes = [C.c_void_p, _d, _d, _i32, _i64, _i64, _i64, _i64, _i64, _u64, _u64]
        L.pnbx_oracle_tree_dump_payload.argtypes = [C.c_void_p, _d, _d, _d, _d]
        L.pnbx_oracle_p2m.argtypes = [_d, _d, C.c_int64, _d, C.c_int, _d]
        L.pnbx_oracle_p2m.restype = None
        L.pnbx_oracle_m2m.argtypes = [_d, _d, C.c_int, _d]
        L.pnbx_oracle_m2m.restype = None
        L.pnbx_oracle_m2p.argtypes = [_d, _d, C.c_double, C.c_int, _d, _d]
        L.pnbx_oracle_m2p.restype = None
        _lib = L
    return _lib


def _p(a, t=_d):
    return None if a is None else a.ctypes.data_as(t)


def _f64(a, shape_last=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape_last is not None:
        assert a.ndim == 2 and a.shape[1] == shape_last
    return a


def _kcode(kernel):
    return -1 if kernel is None else int(kernel)


def num_threads() -> int:
    return lib().pnbx_oracle_num_threads()


def set_num_threads(n: int) -> None:
    lib().pnbx_oracle_set_num_threads(int(n))


def direct(pos, mass=None, h=None, targets=None, kernel=None, want=3):
    """Reference direct summation (direct.rs). kernel: None | 0 | 1. Returns (pot|None, acc|None)."""
    pos = _f64(pos, 3)
    mass = _f64(mass)
    h = _f64(h)
    targets = _f64(targets, 3)
    n = pos.shape[0]
    m = n if targets is None else targets.shape[0]
    pot = np.empty(m) if want & 1 else None
    acc = np.empty((m, 3)) if want & 2 else None
    rc = lib().pnbx_oracle_direct(_p(pos), _p(mass), _p(h), n, _p(targets), m, _kcode(kernel), want, _p(pot), _p(acc))
    if rc:
        raise ValueError(lib().pnbx_oracle_last_error().decode())
    return pot, acc


def kernel_potential(kind, r, h):
    return lib().pnbx_oracle_kernel_potential(int(kind), float(r), float(h))


def kernel_accel_factor(kind, r, h):
    return lib().pnbx_oracle_kernel_accel_factor(int(kind), float(r), float(h))


class Tree:
    """Reference Octree (tree.rs) as driven by the PyO3 class (gravity.rs:113-445)."""

    def __init__(self, pos, mass=None, leaf_capacity=32, multipole_order=0, h=None, kernel=None):
        self.pos = _f64(pos, 3)
        mass = _f64(mass)
        h = _f64(h)
        if kernel is None and h is not None:
            raise ValueError("softenings require an explicit kernel; pass kernel=0/1 (or omit softenings)")
        k = 0 if kernel is None else int(kernel)
        self.n = self.pos.shape[0]
        self._h = lib().pnbx_oracle_tree_create(_p(self.pos), _p(mass), _p(h), self.n, int(leaf_capacity), int(multipole_order), k)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().pnbx_oracle_tree_destroy(self._h)
            self._h = None

    def build_mass(self, mass=None):
        lib().pnbx_oracle_tree_build_mass(self._h, _p(_f64(mass)))

    def set_softenings(self, h=None):
        lib().pnbx_oracle_tree_set_softenings(self._h, _p(_f64(h)))

    def set_kernel(self, kernel=None):
        lib().pnbx_oracle_tree_set_kernel(self._h, 0 if kernel is None else int(kernel))

    def eval(self, theta, targets=None, want=3, counters=False, begin=0, count=None):
        """compute_* (targets=None: own particles [begin, begin+count), default all) / *_at_points."""
        targets = _f64(targets, 3)
        m = (self.n - begin if count is None else int(count)) if targets is None else targets.shape[0]
        pot = np.empty(m) if want & 1 else None
        acc = np.empty((m, 3)) if want & 2 else None
        cnt = np.zeros(4, dtype=np.int64) if counters else None
        rc = lib().pnbx_oracle_tree_eval_range(self._h, _p(targets), int(begin), m, float(theta), want, _p(pot), _p(acc),
                                               _p(cnt, _i64))
        if rc:
            raise ValueError(lib().pnbx_oracle_last_error().decode())
        if counters:
            return pot, acc, dict(zip(("visits", "accepts", "leaf_visits", "leaf_particles"), cnt.tolist()))
        return pot, acc

    def info(self):
        a = np.zeros(7, dtype=np.int64)
        lib().pnbx_oracle_tree_info(self._h, _p(a, _i64))
        return dict(zip(("n_particles", "n_nodes", "n_leaves", "depth", "n_moments", "has_payload", "has_hmax"), a.tolist()))

    def topology(self):
        nn = self.info()["n_nodes"]
        out = dict(
            center=np.empty((nn, 3)), half=np.empty(nn), depth=np.empty(nn, np.int32),
            first_subnode=np.empty(nn, np.int64), next_branch=np.empty(nn, np.int64),
            leaf_start=np.empty(nn, np.int64), leaf_count=np.empty(nn, np.int64),
            leaf_particles=np.empty(self.n, np.int64), path_hi=np.empty(nn, np.uint64), path_lo=np.empty(nn, np.uint64),
        )
        lib().pnbx_oracle_tree_dump_topology(
            self._h, _p(out["center"]), _p(out["half"]), _p(out["depth"], _i32), _p(out["first_subnode"], _i64),
            _p(out["next_branch"], _i64), _p(out["leaf_start"], _i64), _p(out["leaf_count"], _i64),
            _p(out["leaf_particles"], _i64), _p(out["path_hi"], _u64), _p(out["path_lo"], _u64))
        return out

    def payload(self):
        inf = self.info()
        nn, k = inf["n_nodes"], inf["n_moments"]
        mass = np.empty(nn)
        com = np.empty((nn, 3))
        hmax = np.empty(nn) if inf["has_hmax"] else None
        mom = np.zeros((nn, k))
        rc = lib().pnbx_oracle_tree_dump_payload(self._h, _p(mass), _p(com), _p(hmax), _p(mom))
        if rc:
            raise ValueError(lib().pnbx_oracle_last_error().decode())
        if k == 1:
            mom[:, 0] = mass  # order 0: no multipole array in the reference, the monopole is the BH mass
        return dict(mass=mass, com=com, hmax=hmax, moments=mom)


def p2m(pos, mass, center, order):
    pos = _f64(pos, 3)
    mass = _f64(mass)
    center = _f64(center)
    out = np.zeros(56)
    lib().pnbx_oracle_p2m(_p(pos), _p(mass), pos.shape[0], _p(center), int(order), _p(out))
    return out


def m2m(child, shift, order):
    child = _f64(child)
    shift = _f64(shift)
    out = np.zeros(56)
    lib().pnbx_oracle_m2m(_p(child), _p(shift), int(order), _p(out))
    return out


def m2p(moments, dxyz, order, eps2=0.0):
    moments = _f64(moments)
    dxyz = _f64(dxyz)
    pot = np.zeros(1)
    acc = np.zeros(3)
    lib().pnbx_oracle_m2p(_p(moments), _p(dxyz), float(eps2), int(order), _p(pot), _p(acc))
    return float(pot[0]), acc
