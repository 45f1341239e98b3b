"""Breakdown of the host-API tree path (construct + potentials) at N = 1e7 with pageable vs pinned host arrays."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "pynbody-extras_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from benchmarks.synthetic import nfw_disc  # noqa: E402
from pynbodyext.gravity import Gravity, KernelKind  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
pos, mass, h = nfw_disc(n, seed=3)


def step(tag, p, m, hh):
    t0 = time.perf_counter()
    g = Gravity(p, m, softening=hh, kernel=KernelKind.Spline)
    t1 = time.perf_counter()
    tree = g.tree
    t2 = time.perf_counter()
    out = g.tree_potentials(theta=0.7)
    t3 = time.perf_counter()
    del g, tree
    t4 = time.perf_counter()
    print(tag, "init %.1f build %.1f eval %.1f del %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3), flush=True)


for i in range(3):
    step(f"pageable {i}", pos, mass, h)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
pp, pm, ph = pin(pos), pin(mass), pin(h)
for i in range(3):
    step(f"pinned {i}", pp, pm, ph)

# bench-like conditions: a device-resident tree built on torch's stream stays alive while the host path runs
from pynbodyext.gravity import device as gdev  # noqa: E402
d = torch.device("cuda", 0)
dt = gdev.OctreeDevice(*(torch.from_numpy(a).to(d) for a in (pos, mass)), 8, 3, torch.from_numpy(h).to(d), 1)
dt.eval(0.7, 1)
torch.cuda.synchronize()
for i in range(3):
    step(f"pinned+live device tree {i}", pp, pm, ph)
dt.walk_counters(0.7)
for i in range(3):
    step(f"after counters {i}", pp, pm, ph)
