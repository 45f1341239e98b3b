"""Kernel-only timing of the direct-sum kernel (CUDA events around the launch) for the library selected by
$PNBX_GRAVITY_LIB; used to compare tuning variants. Prints: tag kernel_ms Ginteractions/s."""
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
want = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = sys.argv[3] if len(sys.argv) > 3 else "plummer_const"  # newton | plummer_const | plummer_pair | spline
pos, mass = hernquist(n, seed=2)
d = torch.device("cuda", 0)
hv = np.full(n, 0.01) if mode == "plummer_const" else np.random.default_rng(0).uniform(0.005, 0.02, n)
dp, dm, dh = (torch.from_numpy(a).to(d) for a in (pos, mass, hv))
kern = {"newton": None, "plummer_const": 0, "plummer_pair": 0, "spline": 1}[mode]
if mode == "newton":
    dh = None
best = 1e30
for i in range(4):
    gdev.direct_device(dp, dm, dh, kernel=kern, want=want, kernel_events=True)
    torch.cuda.synchronize()
    if i:
        best = min(best, gdev.last_kernel_ms())
print(os.path.basename(os.environ.get("PNBX_GRAVITY_LIB", "default")), mode, "want", want, f"{best:.2f} ms", f"{n * (n - 1) / best / 1e6:.1f} Ginteractions/s")
