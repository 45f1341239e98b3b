"""Staged host<->device copies (csrc/staging.cu): large PAGEABLE numpy arrays are moved in 4 MB chunks by several host
threads through pinned double buffers. The bytes that arrive must be exactly the bytes a plain copy delivers, for sizes
that are not multiples of the chunk, in both directions — checked by running the same call in a child process with
staging disabled (PNBX_STAGING_THREADS=0; the setting is read once per process) and comparing bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import os, sys
sys.path.insert(0, os.path.join({root!r}, "pynbody-extras_b200")); sys.path.insert(0, {root!r})
import numpy as np
import pynbodyext._rust as r
rng = np.random.default_rng(77)
n = 2_200_003                                   # pos 52.8 MB, masses 17.6 MB, potentials 17.6 MB: all above the 16 MB threshold
pos = rng.uniform(-1.0, 1.0, (n, 3))
m = rng.uniform(0.5, 1.5, n) / n
t = r.Octree(pos, m, 16, 2)
pot = t.compute_potentials(0.8)
acc = t.compute_accelerations(0.8)
q = rng.uniform(-1.0, 1.0, (64, 3))
dp = r.direct_potentials_at_points_py(pos, q, m)   # staged H2D only, tiny result
np.savez({out!r}, pot=pot, acc=acc, dp=dp)
"""


def run_child(tmp_path, name, threads):
    out = str(tmp_path / (name + ".npz"))
    env = dict(os.environ)
    env["PNBX_STAGING_THREADS"] = str(threads)
    subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT, out=out)], check=True, env=env, timeout=600)
    return np.load(out)


def test_staged_copies_deliver_the_same_bytes_as_plain_copies(tmp_path):
    plain = run_child(tmp_path, "plain", 0)
    for threads in (1, 3, 8):
        staged = run_child(tmp_path, "staged%d" % threads, threads)
        for k in ("pot", "acc", "dp"):
            assert np.array_equal(staged[k], plain[k]), (threads, k)
    assert np.isfinite(plain["pot"]).all() and (plain["pot"] < 0).all()
