"""Multi-GPU gravity: one process per GPU (torchrun), targets sharded across ranks, sources
replicated with ONE all-gather over NCCL/NVLink; no reduction step because every target's sum
is complete on its owner rank. Not in the reference (it is single-process rayon).

Host-side logic (shard bounds, padded pack, gather + un-pad) is backend-agnostic and is covered by
world_size-2 gloo tests on CPU; the compute step calls the CUDA C-ABI and has no CPU fallback.
"""
from __future__ import annotations

import numpy as np

ROW = 5  # packed source row: x, y, z, mass, softening (float64)


def shard_bounds(n: int, world: int):
    """Contiguous, balanced target/source shards: rank r owns [b[r], b[r+1])."""
    return [(n * r) // world for r in range(world + 1)]


def pack_shard(pos, mass, h, lo, hi, per):
    """Rows (x,y,z,m,h) of the rank's shard, zero-padded to `per` rows (equal-size all-gather)."""
    out = np.zeros((per, ROW), dtype=np.float64)
    c = hi - lo
    out[:c, 0:3] = pos[lo:hi]
    out[:c, 3] = 1.0 if mass is None else mass[lo:hi]
    if h is not None:
        out[:c, 4] = h[lo:hi]
    return out


def replicate_sources(shard_rows, bounds, group=None):
    """All-gather the padded shards (torch tensor on any device/backend) and drop the padding.
    Returns the (n, ROW) tensor of all sources in original order."""
    import torch
    import torch.distributed as dist

    world = len(bounds) - 1
    per = shard_rows.shape[0]
    if world == 1:
        return shard_rows[: bounds[1] - bounds[0]]
    gathered = torch.empty((world * per, ROW), dtype=shard_rows.dtype, device=shard_rows.device)
    dist.all_gather_into_tensor(gathered, shard_rows.contiguous(), group=group)
    if per * world == bounds[-1]:
        return gathered
    return torch.cat([gathered[r * per: r * per + (bounds[r + 1] - bounds[r])] for r in range(world)])


def direct_sharded(pos, mass, h, kernel, want, rank, world, device, targets=None, group=None, precision=None):
    """Sharded direct summation. Every rank passes the same host arrays (or at least its own shard
    of them filled in); rank r uploads only rows [lo, hi), the all-gather replicates the sources on
    every GPU, and rank r evaluates its target shard. Returns this rank's (pot, acc) numpy shards
    and the shard bounds via attributes of the tuple order: (pot | None, acc | None, (lo, hi))."""
    import torch

    from . import device as gdev

    n = pos.shape[0]
    bounds = shard_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    per = max(bounds[r + 1] - bounds[r] for r in range(world))
    dev = torch.device("cuda", device)
    rows_h = torch.from_numpy(pack_shard(pos, mass, h, lo, hi, per))
    rows = rows_h.to(dev, non_blocking=True)
    allrows = replicate_sources(rows, bounds, group)
    d_pos = allrows[:, 0:3].contiguous()
    d_mass = allrows[:, 3].contiguous()
    d_h = allrows[:, 4].contiguous() if h is not None else None
    if targets is None:
        pot, acc = gdev.direct_device(d_pos, d_mass, d_h, kernel=kernel, want=want, tgt_begin=lo, count=hi - lo,
                                      precision=precision)
        tl, th = lo, hi
    else:
        tb = shard_bounds(targets.shape[0], world)
        tl, th = tb[rank], tb[rank + 1]
        d_t = torch.from_numpy(np.ascontiguousarray(targets[tl:th])).to(dev)
        pot, acc = gdev.direct_device(d_pos, d_mass, d_h, kernel=kernel, want=want, targets=d_t, precision=precision)
    pot_h = pot.cpu().numpy() if pot is not None else None
    acc_h = acc.cpu().numpy() if acc is not None else None
    return pot_h, acc_h, (tl, th)


def tree_sharded(pos, mass, h, kernel, want, theta, rank, world, device, leaf_capacity=8, multipole_order=3,
                 targets=None, group=None, precision=None):
    """Sharded tree gravity: shard upload + one all-gather (as direct_sharded), then every rank builds the SAME
    tree from the replicated sources (deterministic kernels => identical topology on every GPU; the build is
    N-linear and ~10 ms per 1e7 particles) and walks only its own target shard.

    Self mode shards are block-cyclic in TREE order (blocks of 4096 consecutive tree-order particles dealt round-robin:
    coherent warps, and every rank samples every region of the tree, so ranks cost the same), so a rank's results
    belong to scattered particles: returns (pot, acc, idx) with idx the original particle indices (int64) of this
    rank's results. At-points mode returns (pot, acc, (lo, hi)) for the contiguous target slice."""
    import torch

    from . import device as gdev

    n = pos.shape[0]
    bounds = shard_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    per = max(bounds[r + 1] - bounds[r] for r in range(world))
    dev = torch.device("cuda", device)
    rows = torch.from_numpy(pack_shard(pos, mass, h, lo, hi, per)).to(dev, non_blocking=True)
    allrows = replicate_sources(rows, bounds, group)
    d_pos = allrows[:, 0:3].contiguous()
    d_mass = allrows[:, 3].contiguous()
    d_h = allrows[:, 4].contiguous() if h is not None else None
    tree = gdev.OctreeDevice(d_pos, d_mass, leaf_capacity, multipole_order, d_h, kernel, precision=precision)
    if targets is None:
        pot, acc = tree.eval(theta, want, shard=(rank, world))
        idx = tree.order(shard=(rank, world)).cpu().numpy()
        return (pot.cpu().numpy() if pot is not None else None, acc.cpu().numpy() if acc is not None else None, idx)
    else:
        tb = shard_bounds(targets.shape[0], world)
        tl, th = tb[rank], tb[rank + 1]
        d_t = torch.from_numpy(np.ascontiguousarray(targets[tl:th])).to(dev)
        pot, acc = tree.eval(theta, want, targets=d_t)
    pot_h = pot.cpu().numpy() if pot is not None else None
    acc_h = acc.cpu().numpy() if acc is not None else None
    return pot_h, acc_h, (tl, th)
