"""Tree-gravity timings on synthetic sets (BASELINE.json configs 3/4 shapes): construct, walk, end to end.
Usage: python benchmarks/tree_timing.py [--n 1000000] [--theta 0.7] [--order 3] [--set nfw|hernquist|plummer|zoom]"""
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

from benchmarks import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--theta", type=float, default=0.7)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--leaf", type=int, default=8)
    ap.add_argument("--set", default="nfw")
    ap.add_argument("--kernel", type=int, default=1)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--oracle", type=int, default=0, help="also time the CPU oracle on this many targets")
    args = ap.parse_args()
    import pynbodyext._rust as r

    if args.set == "nfw":
        pos, m, h = synthetic.nfw_disc(args.n, seed=3)
    elif args.set == "zoom":
        pos, m, h = synthetic.zoom_set(args.n, seed=4)
    elif args.set == "hernquist":
        pos, m = synthetic.hernquist(args.n, seed=2)
        h = np.full(args.n, 0.01)
    else:
        pos, m = synthetic.plummer(args.n, seed=1)
        h = None
    kern = args.kernel if h is not None else None
    out = {"n": args.n, "set": args.set, "theta": args.theta, "order": args.order, "leaf": args.leaf}
    r.Octree(pos[:1000], m[:1000], args.leaf, args.order)  # warm up context
    for _ in range(2):  # warm the stream-ordered memory pool, then free again
        tree = r.Octree(pos, m, args.leaf, args.order, h, kern)
        del tree
    print("---- steady-state construct ----", file=sys.stderr)
    t0 = time.perf_counter()
    tree = r.Octree(pos, m, args.leaf, args.order, h, kern)
    out["construct_s"] = time.perf_counter() - t0
    print("---- end construct ----", file=sys.stderr)
    out.update({k: v for k, v in tree.info().items() if k in ("n_nodes", "n_leaves", "depth")})
    for name, want in (("pot", 1), ("acc", 2)):
        ts = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            res = tree._eval(None, args.theta, want)
            ts.append(time.perf_counter() - t0)
        out[f"walk_{name}_s"] = min(ts)
        out[f"particles_per_s_{name}"] = args.n / min(ts)
    if args.oracle:
        from oracle import oracle as O
        t0 = time.perf_counter()
        ot = O.Tree(pos, m, args.leaf, args.order, h, kern)
        out["oracle_construct_s"] = time.perf_counter() - t0
        idx = np.linspace(0, args.n - 1, args.oracle).astype(np.int64)
        q = np.ascontiguousarray(pos[idx])
        t0 = time.perf_counter()
        p_o, a_o, cnt = ot.eval(args.theta, targets=q, want=3, counters=True)
        dt = time.perf_counter() - t0
        out["oracle_targets_per_s"] = 2 * args.oracle / dt  # pot + acc passes
        out["oracle_counters_per_target"] = {k: v / (2 * args.oracle) for k, v in cnt.items()}
        p_g, a_g = tree._eval(q, args.theta, 3)
        out["rms_rel_pot_vs_oracle"] = float(np.sqrt((((p_g - p_o) / p_o) ** 2).mean()))
        out["rms_rel_acc_vs_oracle"] = float(np.sqrt((((a_g - a_o) ** 2).sum(1) / (a_o ** 2).sum(1)).mean()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
