"""Device-resident handoff (not in the reference): the same C-ABI calls with
``mem_space = PNBX_MEM_DEVICE``, taking and returning ``torch`` CUDA tensors (float64) on the
caller's current stream. torch is used for device memory, streams and torch.distributed only.
"""
from __future__ import annotations

import ctypes as C

import pynbodyext._rust as _b

FLAG_KERNEL_EVENTS = 1


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


def last_kernel_ms() -> float:
    ms = C.c_double(0.0)
    L = _b._load()
    L.pnbx_last_kernel_ms.argtypes = [C.POINTER(C.c_double)]
    _b._check(L.pnbx_last_kernel_ms(C.byref(ms)))
    return ms.value


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
