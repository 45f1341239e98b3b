"""torchrun worker for tests/test_sharded_nccl.py: one process per GPU, NCCL all-gather of the source shards
(pynbodyext.gravity.sharded), every rank checks its own target shard against the CPU oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "pynbody-extras_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from benchmarks.synthetic import nfw_disc, rz_grid_targets  # noqa: E402
from oracle import oracle as O  # noqa: E402
from pynbodyext.gravity.sharded import direct_sharded, tree_sharded  # noqa: E402


def rms_rel(p, ref):
    return float(np.sqrt((((p - ref) / ref) ** 2).mean()))


def rms_rel_vec(a, ref):
    return float(np.sqrt((((a - ref) ** 2).sum(1) / (ref ** 2).sum(1)).mean()))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 30_011  # not divisible by the world size: ragged shards, padded all-gather
    pos, m, h = nfw_disc(n, seed=17)
    # direct, self and at-points
    p_o, a_o = O.direct(pos, m, h, kernel=1)
    pot, acc, (lo, hi) = direct_sharded(pos, m, h, kernel=1, want=3, rank=rank, world=world, device=local)
    assert (lo, hi) == ((n * rank) // world, (n * (rank + 1)) // world)
    assert rms_rel(pot, p_o[lo:hi]) < 1e-5 and rms_rel_vec(acc, a_o[lo:hi]) < 1e-5
    q = rz_grid_targets(1001, seed=5, rmax=1.0)
    pq_o, _ = O.direct(pos, m, h, targets=q, kernel=1, want=1)
    pq, _, (tl, th) = direct_sharded(pos, m, h, kernel=1, want=1, rank=rank, world=world, device=local, targets=q)
    assert rms_rel(pq, pq_o[tl:th]) < 1e-5
    # tree: block-cyclic tree-order shards with their scatter map
    ot = O.Tree(pos, m, 8, 3, h, 1)
    tp_o, ta_o = ot.eval(0.7)
    tp, ta, idx = tree_sharded(pos, m, h, kernel=1, want=3, theta=0.7, rank=rank, world=world, device=local)
    assert rms_rel(tp, tp_o[idx]) < 1e-5 and rms_rel_vec(ta, ta_o[idx]) < 1e-5
    # the shards of all ranks are disjoint and complete
    seen = torch.zeros(n, dtype=torch.int32, device=f"cuda:{local}")
    seen[torch.from_numpy(idx).to(seen.device)] += 1
    dist.all_reduce(seen)
    assert bool((seen == 1).all())
    tq_o, _ = ot.eval(0.7, targets=q, want=1)
    tq, _, (tl, th) = tree_sharded(pos, m, h, kernel=1, want=1, theta=0.7, rank=rank, world=world, device=local, targets=q)
    assert rms_rel(tq, tq_o[tl:th]) < 1e-5
    dist.barrier()
    if rank == 0:
        print("NCCL_SHARDED_OK world=%d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
