"""Device-resident handoff (not in the reference): the same C-ABI calls with
``mem_space = PNBX_MEM_DEVICE``, taking and returning ``torch`` CUDA tensors (float64) on the
caller's current stream. torch is used for device memory, streams and torch.distributed only.
"""
from __future__ import annotations

import ctypes as C

import pynbodyext._rust as _b

FLAG_KERNEL_EVENTS = 1
FLAG_TREE_ORDER = 2
FLAG_BLOCK_CYCLIC = 4
SHARD_BLOCK = 4096


def shard_count(n, world, rank, block=SHARD_BLOCK) -> int:
    """Targets owned by `rank` under block-cyclic tree-order sharding (pnbx_shard_count)."""
    L = _b._load()
    L.pnbx_shard_count.restype = C.c_int64
    L.pnbx_shard_count.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_int32]
    return int(L.pnbx_shard_count(int(n), int(block), int(world), int(rank)))


def _torch():
    import torch
    return torch


def _dptr(t, shape_tail=None):
    torch = _torch()
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise ValueError("device arrays must be contiguous float64 CUDA tensors")
    if shape_tail is not None and (t.ndim != 2 or t.shape[1] != shape_tail):
        raise ValueError("positions must be (N,3) float64 array")
    return t.data_ptr()


def _dev_opts(t, precision, kernel_events):
    torch = _torch()
    o = _b._opts(t.device.index, precision, mem_space=_b.MEM_DEVICE,
                 stream=torch.cuda.current_stream(t.device).cuda_stream)
    if kernel_events:
        o.flags |= FLAG_KERNEL_EVENTS
    return o


def direct_device(pos, mass=None, h=None, kernel=None, want=_b.WANT_ACC, targets=None, tgt_begin=0, count=None,
                  precision=None, kernel_events=False):
    """Direct summation on device tensors. Self mode evaluates sources [tgt_begin, tgt_begin+count).
    Returns (pot | None, acc | None) as float64 CUDA tensors (stream-ordered, not synchronised)."""
    torch = _torch()
    n = pos.shape[0]
    if kernel is None and h is not None:
        raise ValueError("softenings require an explicit kernel; pass kernel=0/1 (or omit softenings)")
    k = _b._kernel_code(kernel, -1)
    if targets is None:
        m = n - tgt_begin if count is None else int(count)
    else:
        m = targets.shape[0]
    pot = torch.empty(m, dtype=torch.float64, device=pos.device) if want & _b.WANT_POT else None
    acc = torch.empty((m, 3), dtype=torch.float64, device=pos.device) if want & _b.WANT_ACC else None
    o = _dev_opts(pos, precision, kernel_events)
    rc = _b._load().pnbx_direct(_dptr(pos, 3), _dptr(mass), _dptr(h), n, _dptr(targets, 3), m, int(tgt_begin), k, want,
                                _dptr(pot), _dptr(acc), C.byref(o))
    _b._check(rc)
    return pot, acc


class OctreeDevice:
    """Octree built from device-resident float64 tensors (no host copies); evaluates into torch tensors on the
    caller's current stream. Same semantics as ``pynbodyext._rust.Octree`` (reference gravity.rs:113-445)."""

    def __init__(self, pos, mass=None, leaf_capacity=32, multipole_order=0, h=None, kernel=None, precision=None):
        if kernel is None and h is not None:
            raise ValueError("softenings require an explicit kernel; pass kernel=0/1 (or omit softenings)")
        self._n = int(pos.shape[0])
        self._dev = pos.device
        self._precision = precision
        self._h = C.c_void_p()
        o = _dev_opts(pos, precision, False)
        _b._check(_b._load().pnbx_tree_create(C.byref(self._h), _dptr(pos, 3), _dptr(mass), _dptr(h), self._n,
                                              int(leaf_capacity), int(multipole_order), _b._kernel_code(kernel, 0),
                                              C.byref(o)))

    def __del__(self):
        h = getattr(self, "_h", None)
        lib = getattr(_b, "_lib", None) if _b is not None else None  # module globals may be gone at interpreter exit
        if h is not None and h.value and lib is not None:
            lib.pnbx_tree_destroy(h)
            self._h = None

    # -- setters on device tensors (gravity.rs:228-265): stream-ordered on the caller's current stream
    def build_mass(self, mass=None):
        torch = _torch()
        o = _b._opts(self._dev.index, None, mem_space=_b.MEM_DEVICE, stream=torch.cuda.current_stream(self._dev).cuda_stream)
        if mass is not None and mass.shape != (self._n,):
            raise ValueError("masses must be length N")
        _b._check(_b._load().pnbx_tree_build_mass_ex(self._h, _dptr(mass), C.byref(o)))

    def set_softenings(self, h=None):
        torch = _torch()
        o = _b._opts(self._dev.index, None, mem_space=_b.MEM_DEVICE, stream=torch.cuda.current_stream(self._dev).cuda_stream)
        if h is not None and h.shape != (self._n,):
            raise ValueError("softenings must be length N")
        _b._check(_b._load().pnbx_tree_set_softenings_ex(self._h, _dptr(h), C.byref(o)))

    def set_kernel(self, kernel=None):
        _b._check(_b._load().pnbx_tree_set_kernel(self._h, _b._kernel_code(kernel, 0)))

    def info(self):
        inf = _b.pnbx_tree_info()
        _b._check(_b._load().pnbx_tree_get_info(self._h, C.byref(inf)))
        return {f: getattr(inf, f) for f, _ in _b.pnbx_tree_info._fields_}

    def _shard(self, o, shard):
        if shard is not None:
            rank, world = shard
            o.flags |= FLAG_TREE_ORDER | FLAG_BLOCK_CYCLIC
            o.shard_rank, o.shard_world, o.shard_block = int(rank), int(world), SHARD_BLOCK
            return shard_count(self._n, world, rank)
        return None

    def order(self, begin=0, count=None, shard=None):
        """Original particle index (int64 CUDA tensor) of tree-order positions [begin, begin+count), or of the
        block-cyclic shard ``shard=(rank, world)``."""
        torch = _torch()
        o = _b._opts(self._dev.index, None, mem_space=_b.MEM_DEVICE, stream=torch.cuda.current_stream(self._dev).cuda_stream)
        ms = self._shard(o, shard)
        m = ms if ms is not None else (self._n - begin if count is None else int(count))
        out = torch.empty(m, dtype=torch.int64, device=self._dev)
        L = _b._load()
        _b._check(L.pnbx_tree_get_order(self._h, int(begin), m, out.data_ptr(), C.byref(o)))
        return out

    def eval(self, theta, want=_b.WANT_POT, targets=None, tgt_begin=0, count=None, kernel_events=False,
             tree_order=False, shard=None, precision=None):
        """(pot | None, acc | None) for own particles [tgt_begin, tgt_begin+count) or for `targets` (M,3).
        tree_order=True: the range selects tree-order positions and results come back in that order
        (use .order(tgt_begin, count) to scatter them) — coherent warps, the right sharding for multi-GPU.
        shard=(rank, world): block-cyclic tree-order shard of that rank (equal cost per rank); .order(shard=...)."""
        torch = _torch()
        if targets is None:
            m = shard_count(self._n, shard[1], shard[0]) if shard is not None else (
                self._n - tgt_begin if count is None else int(count))
            ref = torch.empty(0, device=self._dev, dtype=torch.float64)
        else:
            m = targets.shape[0]
            ref = targets
        pot = torch.empty(m, dtype=torch.float64, device=self._dev) if want & _b.WANT_POT else None
        acc = torch.empty((m, 3), dtype=torch.float64, device=self._dev) if want & _b.WANT_ACC else None
        o = _dev_opts(ref, self._precision if precision is None else precision, kernel_events)
        if tree_order:
            o.flags |= FLAG_TREE_ORDER
        self._shard(o, shard)
        _b._check(_b._load().pnbx_tree_eval(self._h, _dptr(targets, 3), m, int(tgt_begin), float(theta), want,
                                            _dptr(pot), _dptr(acc), C.byref(o)))
        return pot, acc

    def walk_counters(self, theta, tgt_begin=0, count=None, tree_order=False, shard=None):
        import numpy as np
        out = np.zeros(5, dtype=np.int64)
        o = _b._opts(self._dev.index, None)
        if tree_order:
            o.flags |= FLAG_TREE_ORDER
        ms = self._shard(o, shard)
        m = ms if ms is not None else (self._n - tgt_begin if count is None else int(count))
        _b._check(_b._load().pnbx_tree_walk_counters(self._h, None, m, int(tgt_begin), float(theta), out.ctypes.data,
                                                     C.byref(o)))
        d = dict(zip(("visits", "accepts", "leaf_visits", "leaf_particles", "warp_visits"), out.tolist()))
        return d


def last_kernel_ms() -> float:
    ms = C.c_double(0.0)
    L = _b._load()
    L.pnbx_last_kernel_ms.argtypes = [C.POINTER(C.c_double)]
    _b._check(L.pnbx_last_kernel_ms(C.byref(ms)))
    return ms.value


def trim_memory() -> None:
    """Return the library's cached device memory blocks that are not in use (pnbx_trim_memory)."""
    _b._load().pnbx_trim_memory()


def launch_count() -> int:
    L = _b._load()
    L.pnbx_launch_count.restype = C.c_int64
    return int(L.pnbx_launch_count())


def measure_fp32_peak(device=0, variant=0) -> float:
    """TFLOP/s of the FMA-chain microbenchmark (variant 0 scalar FFMA, 1 packed FFMA2)."""
    t = C.c_double(0.0)
    L = _b._load()
    L.pnbx_measure_fp32_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    _b._check(L.pnbx_measure_fp32_peak(int(device), int(variant), C.byref(t)))
    return t.value
