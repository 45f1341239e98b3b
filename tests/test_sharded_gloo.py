"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: shard bounds, padded pack and the
single all-gather that replicates the sources (pynbodyext/gravity/sharded.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pynbodyext.gravity.sharded import pack_shard, replicate_sources, shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 7, 1000, 1_000_003):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and len(b) == w + 1
            sizes = np.diff(b)
            assert sizes.min() >= 0 and sizes.max() - sizes.min() <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    pos = rng.random((n, 3))
    mass = rng.random(n)
    h = rng.random(n)
    b = shard_bounds(n, world)
    per = max(np.diff(b))
    # each rank only fills in its own shard; the rest of its view is garbage on purpose
    mine = pack_shard(pos, mass, h, b[rank], b[rank + 1], per)
    rows = replicate_sources(torch.from_numpy(mine), b)
    ok = (np.array_equal(rows[:, 0:3].numpy(), pos) and np.array_equal(rows[:, 3].numpy(), mass)
          and np.array_equal(rows[:, 4].numpy(), h))
    q.put((rank, bool(ok), tuple(rows.shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 11, 1001])
def test_all_gather_replicates_sources_world2(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (n, 5) for _, _, shape in res)
