"""Walk statistics + kernel time at a given N (NFW+disc set): per-target counters, warp-union visits, lane efficiency."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "pynbody-extras_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402
from benchmarks.synthetic import nfw_disc, zoom_set  # noqa: E402
from pynbodyext.gravity import device as gdev  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
which = sys.argv[2] if len(sys.argv) > 2 else "nfw"
orders = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [3]
pos, m, h = nfw_disc(n, seed=3) if which == "nfw" else zoom_set(n, seed=4)
d = torch.device("cuda", 0)
dp, dm, dh = (torch.from_numpy(a).to(d) for a in (pos, m, h))
for order in orders:
    t = gdev.OctreeDevice(dp, dm, 8, order, dh, 1)
    for want in (1, 2):
        best = 1e9
        for _ in range(int(os.environ.get("REPS", "3"))):
            t.eval(0.7, want, kernel_events=True)
            torch.cuda.synchronize()
            best = min(best, gdev.last_kernel_ms())
        print("order", order, "want", want, "kernel ms", round(best, 3), "targets/s", n / best * 1e3)
c = t.walk_counters(0.7)
print({k: v / n for k, v in c.items() if k != "warp_visits"}, "warp visits/warp", c["warp_visits"] / (n / 32),
      "lane eff", c["visits"] / (32 * c["warp_visits"]))
