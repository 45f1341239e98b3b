"""Generates tests/golden/*.npz from the CPU oracle (NOT from the reference: it cannot be built here).
The fixtures freeze the oracle's outputs on small seeded inputs so that (a) later edits of the oracle are caught and
(b) the GPU tests have oracle-free expected values. Re-run only when the oracle is deliberately changed:

    python tests/golden/make_golden.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from benchmarks.synthetic import hernquist, plummer, uniform_cube  # noqa: E402
from oracle import oracle as O  # noqa: E402


def direct_cases():
    out = {}
    pos, m = plummer(700, seed=11)
    hv = np.random.default_rng(12).uniform(0.01, 0.1, 700)
    q, _ = uniform_cube(64, 13, with_masses=False)
    out["pos"], out["mass"], out["h"], out["q"] = pos, m, hv, q
    for name, kern, h in (("newton", None, None), ("plummer", 0, hv), ("spline", 1, hv)):
        p, a = O.direct(pos, m, h, kernel=kern)
        out[f"{name}_self_pot"], out[f"{name}_self_acc"] = p, a
        p, a = O.direct(pos, m, h, targets=q, kernel=kern)
        out[f"{name}_pts_pot"], out[f"{name}_pts_acc"] = p, a
    pos_s, m_s = uniform_cube(200, 14)  # n < 512: the symmetric pair-loop path of direct.rs:130-156
    p, a = O.direct(pos_s, m_s)
    out["small_pos"], out["small_mass"], out["small_pot"], out["small_acc"] = pos_s, m_s, p, a
    return out


def tree_cases():
    out = {}
    pos, m = hernquist(3000, seed=21)
    h = np.random.default_rng(22).uniform(0.0, 0.05, 3000)
    q, _ = plummer(100, seed=23, a=2.0)
    out["pos"], out["mass"], out["h"], out["q"] = pos, m, h, q
    for order in (0, 2, 3, 5):
        t = O.Tree(pos, m, 8, order, h, 1)
        p, a = t.eval(0.7)
        out[f"o{order}_self_pot"], out[f"o{order}_self_acc"] = p, a
        p, a = t.eval(0.7, targets=q)
        out[f"o{order}_pts_pot"], out[f"o{order}_pts_acc"] = p, a
    t = O.Tree(pos, m, 8, 3, h, 1)
    topo, pay = t.topology(), t.payload()
    for k in ("center", "half", "depth", "first_subnode", "next_branch", "leaf_count", "path_hi", "path_lo"):
        out[f"topo_{k}"] = topo[k]
    leaves = np.nonzero(topo["leaf_count"] >= 0)[0]
    out["topo_leaf_particles_in_node_order"] = np.concatenate(
        [topo["leaf_particles"][topo["leaf_start"][i]: topo["leaf_start"][i] + topo["leaf_count"][i]] for i in leaves])
    out["pay_mass"], out["pay_com"], out["pay_hmax"], out["pay_moments"] = pay["mass"], pay["com"], pay["hmax"], pay["moments"]
    _, _, c = t.eval(0.7, want=1, counters=True)
    out["counters"] = np.array([c["visits"], c["accepts"], c["leaf_visits"], c["leaf_particles"]], dtype=np.int64)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "direct_golden.npz"), **direct_cases())
    np.savez_compressed(os.path.join(HERE, "tree_golden.npz"), **tree_cases())
    print("wrote", os.listdir(HERE))
