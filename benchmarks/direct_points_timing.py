"""Kernel-only timing of direct sums AT POINTS (M query points, N sources) for the softened kernels.
Usage: python benchmarks/direct_points_timing.py [N] [M] [plummer_pair|spline]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "pynbody-extras_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from benchmarks.synthetic import hernquist  # noqa: E402
from pynbodyext.gravity import device as gdev  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
mode = sys.argv[3] if len(sys.argv) > 3 else "spline"
pos, mass = hernquist(n, seed=2)
q, _ = hernquist(m, seed=3)
d = torch.device("cuda", 0)
hv = np.random.default_rng(0).uniform(0.005, 0.02, n)
dp, dm, dh, dq = (torch.from_numpy(a).to(d) for a in (pos, mass, hv, q))
kern = {"plummer_pair": 0, "spline": 1}[mode]
best = 1e30
for i in range(4):
    gdev.direct_device(dp, dm, dh, kernel=kern, want=2, targets=dq, kernel_events=True)
    torch.cuda.synchronize()
    if i:
        best = min(best, gdev.last_kernel_ms())
print(mode, "N", n, "M", m, f"{best:.2f} ms", f"{n * m / best / 1e6:.1f} Ginteractions/s", "sort_min", os.environ.get("PNBX_DIRECT_SORT_MIN", "default"))
