"""Small end-to-end case for compute-sanitizer (memcheck): direct (all softening modes, ragged sizes) + tree
(build, payload, walk self / points / shards). Exits non-zero on any mismatch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "pynbody-extras_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402

import pynbodyext._rust as r  # noqa: E402
from benchmarks.synthetic import plummer  # noqa: E402

pos, m = plummer(1537, seed=3)
hv = np.random.default_rng(1).uniform(0.01, 0.1, 1537)
q = pos[:301] * 1.1
for kern, h in ((None, None), (0, np.full(1537, 0.02)), (0, hv), (1, hv)):
    a = r.direct_accelerations_py(pos, m, 0, h, kern)
    p = r.direct_potentials_at_points_py(pos, q, m, 0, h, kern)
    assert np.isfinite(a).all() and np.isfinite(p).all()
for order in (0, 3, 5):
    t = r.Octree(pos, m, 8, order, hv, 1)
    p, a = t._eval(None, 0.7, 3)
    p2, a2 = t._eval(q, 0.7, 3)
    p3, _ = t._eval(None, 0.7, 1, tgt_begin=100, count=700)
    assert np.array_equal(p3, p[100:800]) and np.isfinite(a2).all()
    t.walk_counters(0.7)
print("sanitizer case ok")
