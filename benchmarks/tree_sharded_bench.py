"""BASELINE.json configs[3]/[4]: zoom-shaped dm/gas/star set, N up to 1e8, tree gravity sharded across the GPUs of
one box. One process per GPU (torchrun); every rank owns 1/world of the snapshot (its index range of the deterministic
zoom set, `zoom_range`: the same global set at any world size, no 1e8-particle host array per rank), uploads it, ONE all-gather
replicates the sources, every rank builds the identical tree and walks its own target shard.

  torchrun --nproc-per-node 8 benchmarks/tree_sharded_bench.py --particles 100000000
  python benchmarks/tree_sharded_bench.py --particles 100000000          (single GPU)

Prints one JSON line with per-stage device times (max over ranks)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "pynbody-extras_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", dest="n", type=int, default=100_000_000)
    ap.add_argument("--theta", type=float, default=0.7)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--leaf", type=int, default=8)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--grid-targets", type=int, default=0, help="config 5: also evaluate this many (R,z) grid points")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from benchmarks.synthetic import rz_grid_targets, zoom_range
    from pynbodyext.gravity import device as gdev
    from pynbodyext.gravity.sharded import pack_shard, replicate_sources, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.n
    b = shard_bounds(n, world)
    lo, hi = b[rank], b[rank + 1]
    per = max(b[r + 1] - b[r] for r in range(world))
    t0 = time.perf_counter()
    pos, mass, h, _fam = zoom_range(n, lo, hi, seed=4)
    rows_h = torch.from_numpy(pack_shard(pos, mass, h, 0, hi - lo, per)).pin_memory()
    gen_s = time.perf_counter() - t0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def stage(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), r

    m_r = gdev.shard_count(n, world, rank)                            # block-cyclic tree-order shard of this rank
    cap = per + gdev.SHARD_BLOCK
    pot_h = torch.empty(cap, dtype=torch.float64).pin_memory()        # caller-owned pinned result buffers
    acc_h = torch.empty((cap, 3), dtype=torch.float64).pin_memory()
    idx_h = torch.empty(cap, dtype=torch.int64).pin_memory()          # original index of every result row
    best = None
    for rep in range(args.reps):
        t_h2d, rows = stage(lambda: rows_h.to(dev, non_blocking=True))
        t_gather, allrows = stage(lambda: replicate_sources(rows, b))
        t_split, (d_pos, d_mass, d_h) = stage(lambda: (allrows[:, 0:3].contiguous(), allrows[:, 3].contiguous(),
                                                       allrows[:, 4].contiguous()))
        del allrows, rows
        t_build, tree = stage(lambda: gdev.OctreeDevice(d_pos, d_mass, args.leaf, args.order, d_h, 1))
        t_pot, pot = stage(lambda: tree.eval(args.theta, 1, shard=(rank, world))[0])
        t_acc, acc = stage(lambda: tree.eval(args.theta, 2, shard=(rank, world))[1])
        t_d2h, _ = stage(lambda: (pot_h[:m_r].copy_(pot, non_blocking=True), acc_h[:m_r].copy_(acc, non_blocking=True),
                                  idx_h[:m_r].copy_(tree.order(shard=(rank, world)), non_blocking=True)))
        res = {"h2d_ms": t_h2d, "allgather_ms": t_gather, "unpack_ms": t_split, "build_ms": t_build, "walk_pot_ms": t_pot,
               "walk_acc_ms": t_acc, "d2h_ms": t_d2h}
        res["total_pot_ms"] = t_h2d + t_gather + t_split + t_build + t_pot
        res["total_pot_acc_ms"] = res["total_pot_ms"] + t_acc + t_d2h
        if best is None or res["total_pot_acc_ms"] < best["total_pot_acc_ms"]:
            best = res
        info = tree.info()
        finite = bool(torch.isfinite(pot).all() and torch.isfinite(acc).all())
        grid = None
        if args.grid_targets and rep == args.reps - 1:
            tg = rz_grid_targets(args.grid_targets, seed=5)
            tb = shard_bounds(tg.shape[0], world)
            d_t = torch.from_numpy(np.ascontiguousarray(tg[tb[rank]:tb[rank + 1]])).to(dev)
            t_grid, gp = stage(lambda: tree.eval(args.theta, 1, targets=d_t)[0])
            grid = {"targets": int(tg.shape[0]), "walk_pot_ms": t_grid, "finite": bool(torch.isfinite(gp).all())}
        del tree, pot, acc, d_pos, d_mass, d_h
    if rank == 0:
        out = {"config": f"zoom dm/gas/star N={n}, theta={args.theta}, order={args.order}, leaf={args.leaf}, spline softening",
               "n_gpus": world, "n": n, "nodes": info["n_nodes"], "depth": info["depth"], "finite": finite,
               "host_generate_s": gen_s, **best,
               "particles_per_s_pot": n / (best["total_pot_ms"] * 1e-3),
               "particles_per_s_pot_acc": n / (best["total_pot_acc_ms"] * 1e-3), "grid": grid}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
