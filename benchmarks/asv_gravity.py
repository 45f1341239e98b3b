"""The reference's gravity benchmark (benchmarks/bench_gravity.py:74-188) against the B200 backend.

Same ASV classes, parameter grids and timed calls as the reference — `TimeTreeConstruct`, `TimeTreeGravityTheta`,
`TimeTreeGravityOrder`, `TimeTreeGravityFull` on halo 0 of pynbody's `gadget3/data/subhalos_103/subhalo_103`
(15 682 particles), `TimeGravityRealSnapData` on the full file (382 909) — so `asv run` or
`python benchmarks/asv_gravity.py` reproduces `bench_gravity.main()` literally when pynbody and its test data are
present. They are NOT in this image: the loader then falls back to a synthetic stand-in of the same
particle counts (Hernquist halo in kpc-like units, seed 103) and says so; timings on the stand-in are indicative of the
small-N latency of the path, not a reproduction of the reference's data set.

    python benchmarks/asv_gravity.py            # the reference's main(): N=... repeat=10 avg=...s
    python benchmarks/asv_gravity.py --all      # every ASV case once, JSON lines
"""
from __future__ import annotations

import json
import os
import pathlib
import sys
import time
from functools import lru_cache

_REPO_ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (str(_REPO_ROOT / "pynbody-extras_b200"), str(_REPO_ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

from pynbodyext.gravity import Gravity  # noqa: E402

N_HALO, N_SNAP = 15_682, 382_909  # reference tests/conftest.py:50-52
DATA_SOURCE = {}


@lru_cache(maxsize=2)
def _load_pos_mass(only_halo: bool = False):
    """(pos, mass) float64 C-order: the pynbody test snapshot if available, else the synthetic stand-in."""
    try:
        import pynbody
        import pynbody.test_utils
        cwd = os.getcwd()
        try:
            os.chdir(_REPO_ROOT)
            pynbody.test_utils.ensure_test_data_available("gadget", "arepo")
        finally:
            os.chdir(cwd)
        data = pynbody.load(str(_REPO_ROOT / "testdata" / "gadget3" / "data" / "subhalos_103" / "subhalo_103"))
        if only_halo:
            data = data.halos()[0].load_copy()
        DATA_SOURCE[only_halo] = "pynbody testdata subhalo_103"
        return np.asarray(data["pos"], dtype=np.float64, order="C"), np.asarray(data["mass"], dtype=np.float64, order="C")
    except Exception as exc:  # pynbody or its test data missing (this image): same sizes, synthetic positions
        from benchmarks.synthetic import hernquist
        n = N_HALO if only_halo else N_SNAP
        pos, mass = hernquist(n, seed=103, a=30.0, rmax=100.0)
        DATA_SOURCE[only_halo] = f"synthetic stand-in (Hernquist a=30, N={n}): {type(exc).__name__}"
        return pos, mass * 1.0e12


class TimeGravityRealBaseData:
    """Gravity(...).tree / .tree_potentials() on halo 0, as the reference times it."""

    pos, mass = _load_pos_mass(only_halo=True)

    def _gravity(self, **kw):
        return Gravity(self.pos, self.mass, softening=kw.get("softening"), kernel=kw.get("kernel"),
                       leaf_capacity=kw.get("leaf_capacity", 8), multipole_order=kw.get("multipole_order", 0))

    def _construct_tree(self, **kw):
        self._gravity(**kw).tree

    def _construct_and_tree_potentials(self, **kw):
        return self._gravity(**kw).tree_potentials(theta=kw.get("theta", 0.7), leaf_capacity=kw.get("leaf_capacity", 8),
                                                   multipole_order=kw.get("multipole_order", 0), kernel=kw.get("kernel"))


class TimeTreeConstruct(TimeGravityRealBaseData):
    params = [[8, 32, 128], [None, 0.288], [0, 3, 5]]
    param_names = ["leaf_capacity", "softening", "multipole_order"]

    def time_construct_tree(self, leaf_capacity, softening, multipole_order):
        self._construct_tree(leaf_capacity=leaf_capacity, softening=softening, kernel=1, multipole_order=multipole_order)


class TimeTreeGravityTheta(TimeGravityRealBaseData):
    params = [0.5, 0.7, 1.0]
    param_names = ["theta"]

    def time_construct_and_tree_potentials(self, theta):
        self._construct_and_tree_potentials(theta=theta)


class TimeTreeGravityOrder(TimeGravityRealBaseData):
    params = [2, 3, 4, 5]
    param_names = ["multipole_order"]

    def time_construct_and_tree_potentials(self, multipole_order):
        self._construct_and_tree_potentials(multipole_order=multipole_order)


class TimeTreeGravityFull(TimeGravityRealBaseData):
    def time_construct_and_tree_potentials(self):
        self._construct_and_tree_potentials(theta=0.7, softening=0.001, kernel=1, multipole_order=3)


class TimeGravityRealSnapData(TimeGravityRealBaseData):
    pos, mass = _load_pos_mass(only_halo=False)


def _timeit(fn, repeat=10):
    fn()  # warm-up (the reference's ASV run does the same through its setup / first sample)
    t0 = time.perf_counter()
    for _ in range(repeat):
        fn()
    return (time.perf_counter() - t0) / repeat


def run_all():
    """Every ASV case once (mean of 10 calls after one warm-up), one JSON line each."""
    import itertools
    out = []
    b = TimeTreeConstruct()
    for lc, soft, order in itertools.product(*TimeTreeConstruct.params):
        out.append({"case": "TimeTreeConstruct.time_construct_tree", "leaf_capacity": lc, "softening": soft,
                    "multipole_order": order, "s": _timeit(lambda: b.time_construct_tree(lc, soft, order))})
    b = TimeTreeGravityTheta()
    for theta in TimeTreeGravityTheta.params:
        out.append({"case": "TimeTreeGravityTheta.time_construct_and_tree_potentials", "theta": theta,
                    "s": _timeit(lambda: b.time_construct_and_tree_potentials(theta))})
    b = TimeTreeGravityOrder()
    for order in TimeTreeGravityOrder.params:
        out.append({"case": "TimeTreeGravityOrder.time_construct_and_tree_potentials", "multipole_order": order,
                    "s": _timeit(lambda: b.time_construct_and_tree_potentials(order))})
    b = TimeTreeGravityFull()
    out.append({"case": "TimeTreeGravityFull.time_construct_and_tree_potentials", "s": _timeit(b.time_construct_and_tree_potentials)})
    b = TimeGravityRealSnapData()
    out.append({"case": "TimeGravityRealSnapData (theta 0.7, softening 0.001, kernel 1, order 3)", "n": len(b.pos),
                "s": _timeit(lambda: b._construct_and_tree_potentials(theta=0.7, softening=0.001, kernel=1, multipole_order=3))})
    for line in out:
        line["n"] = line.get("n", N_HALO)
        line["data"] = DATA_SOURCE.get(line["n"] == N_HALO, "?")
        print(json.dumps(line))
    return out


def main():
    bench = TimeGravityRealBaseData()
    n_repeat = 10
    bench._construct_and_tree_potentials(theta=0.7, softening=0.001, kernel=1, multipole_order=3)  # context warm-up
    t0 = time.perf_counter()
    for _ in range(n_repeat):
        bench._construct_and_tree_potentials(theta=0.7, softening=0.001, kernel=1, multipole_order=3)
    t1 = time.perf_counter()
    print(f"N={len(bench.pos)} repeat={n_repeat} avg={(t1 - t0) / n_repeat:.6f}s  [{DATA_SOURCE.get(True)}]")


if __name__ == "__main__":
    run_all() if "--all" in sys.argv else main()
