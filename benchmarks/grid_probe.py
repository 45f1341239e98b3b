"""Config-5 shaped probe: potentials at M log-spaced (R,z) grid points from an N-particle zoom set on one GPU, with
the lane-per-target kernel, the warp-per-target kernel (PNBX_WPT_MAX_TARGETS) and the hybrid split
(PNBX_WALK_HYBRID_COST, argv[3]). Prints kernel ms for each."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "pynbody-extras_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from benchmarks.synthetic import rz_grid_targets, zoom_set  # noqa: E402
from pynbodyext.gravity import device as gdev  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ms = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1000, 10000, 125000]
HYB = sys.argv[3] if len(sys.argv) > 3 else "48000"
pos, m, h = zoom_set(n, seed=4)
d = torch.device("cuda", 0)
dp, dm, dh = (torch.from_numpy(a).to(d) for a in (pos, m, h))
t = gdev.OctreeDevice(dp, dm, 8, 3, dh, 1)
for M in ms:
    q = torch.from_numpy(rz_grid_targets(M, seed=5)).to(d)
    res = {}
    for mode, env, pairs in (("lane", "0", "0"), ("warp", "4000000000", "0"), ("hybrid", "0", HYB)):
        os.environ["PNBX_WPT_MAX_TARGETS"] = env
        os.environ["PNBX_WALK_HYBRID_COST"] = pairs
        best = 1e9
        for _ in range(3):
            p = t.eval(0.7, 1, targets=q, kernel_events=True)[0]
            torch.cuda.synchronize()
            best = min(best, gdev.last_kernel_ms())
        res[mode] = (best, p)
    rel = ((res["lane"][1] - res["warp"][1]) / res["lane"][1]).abs().max().item()
    print(f"N={n} M={q.shape[0]}: lane-per-target {res['lane'][0]:.3f} ms, warp-per-target {res['warp'][0]:.3f} ms, "
          f"hybrid(cost>{HYB}) {res['hybrid'][0]:.3f} ms, max rel diff {rel:.2e}", flush=True)
