#!/usr/bin/env python
"""bench.py — headline benchmark of the gravity hot path on B200.

Metric (BASELINE.json): direct-summation Ginteractions/s on config 2 — Hernquist halo N = 1e6,
Plummer softening eps = 0.01, accelerations of all particles (reference call:
Gravity(pos, mass, softening=0.01, kernel=KernelKind.Plummer).direct_accelerations(), i.e.
direct.rs:443-524), fp32 interaction arithmetic checked against the float64 oracle.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n 1000000] [--impl ours|reference]
  torchrun --nproc-per-node N ... bench.py --gpus N ...     (one rank per GPU, targets sharded)

One JSON line on stdout (rank 0). See DESIGN.md §Measurement for every key.

N > 1 (torchrun): `value` = device-resident, one rank per GPU, sources replicated by one NCCL all-gather, targets
sharded; `e2e` = rank 0 alone calls the UNCHANGED host API (Gravity(...).direct_accelerations(), pageable numpy
arrays) with PNBX_DEVICES = all N GPUs while the other ranks wait at a host barrier; the tree_1e8 section
(BASELINE configs 4 and 5: zoom set N = 1e8, 1e6 grid targets) runs for N >= 2 with in-run parity checks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

os.environ.setdefault("OMP_WAIT_POLICY", "passive")  # idle OpenMP workers of the CPU oracle must not spin
ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "pynbody-extras_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

BUILD_TRAFFIC_1E7 = None  # dram bytes of one N=1e7 build from profiles/ (filled once captured)
FLOP_PER_INTERACTION = 20.0  # GPU-Gems-3 convention for an acceleration interaction
EPS = 0.01
METRIC = "direct_sum_ginteractions_per_s"
UNIT = "Ginteractions/s"


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def workload_config(n, n_gpus):
    return {
        "workload": f"Hernquist halo N={n} (a=1, r<=100a, seed 2), direct-sum accelerations, Plummer eps={EPS} "
                    f"(BASELINE.json configs[1])",
        "n_sources": n, "n_targets": n, "softening": EPS, "kernel": "Plummer", "want": "acc",
        "sharding": f"targets/{n_gpus}" if n_gpus > 1 else "none",
        "l2_policy": "inputs (40 MB f64 + 16 MB packed) are re-packed every step; each step streams "
                     "the packed sources ~1000x from L2/HBM, kernel is FP32-pipe bound",
    }


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return json.load(open(path))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_rate(pos, mass, target_seconds=12.0):
    """Time the CPU oracle (kind 'port': C++ restatement of direct.rs:587-658 with OpenMP on all host
    cores; the Rust reference cannot be built in this image) on a bounded target sample of the same
    workload. Returns (interactions/s, cores, sample description)."""
    from oracle import oracle as O
    n = pos.shape[0]
    h = np.full(n, EPS)
    try:  # torchrun exports OMP_NUM_THREADS=1; the reference (rayon, threads=0) uses every core it may run on
        O.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        pass
    cores = O.num_threads()
    probe = max(cores * 32, 1024)
    O.direct(pos[:4096], mass[:4096], h[:4096], targets=pos[:512], kernel=0, want=2)  # spin up the OpenMP team
    t0 = time.perf_counter()
    O.direct(pos, mass, h, targets=np.ascontiguousarray(pos[:probe]), kernel=0, want=2)
    rate = probe * n / (time.perf_counter() - t0)
    m = int(min(n, max(probe, rate * target_seconds / n)))
    m = max(cores, m - m % cores)
    if m < 512:  # keep the parallel (>= 512 targets) code path of the reference
        m = 512
    idx = np.linspace(0, n - 1, m).astype(np.int64)
    t0 = time.perf_counter()
    O.direct(pos, mass, h, targets=np.ascontiguousarray(pos[idx]), kernel=0, want=2)
    dt = time.perf_counter() - t0
    return m * n / dt, cores, f"{m} of {n} targets (evenly strided), all {n} sources, at-points kernel path, {dt:.1f} s"


def oracle_build_info():
    """How the CPU port was compiled (so the GPU/CPU ratio is interpretable)."""
    flags = "-O2 -std=c++17 -fopenmp -ffp-contract=off (oracle/Makefile; no -march=native: the reference's release "\
            "profile does not set target-cpu either)"
    try:
        cxx = subprocess.run(["/usr/bin/g++", "--version"], capture_output=True, text=True).stdout.splitlines()[0]
    except Exception:
        cxx = "g++ (version unavailable)"
    return {"compiler": cxx, "flags": flags}


def real_reference():
    """use the real Rust extension the day it exists — baseline/_ref (a pip --target install of the
    reference) or an importable pynbodyext._rust that is NOT our ctypes shim. Returns the module or None."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(ref_dir):
        return None
    import importlib.util
    for root, _dirs, files in os.walk(ref_dir):
        for f in files:
            if f.startswith("_rust") and f.endswith(".so"):
                try:
                    spec = importlib.util.spec_from_file_location("pynbodyext_ref_rust", os.path.join(root, f))
                    mod = importlib.util.module_from_spec(spec)
                    spec.loader.exec_module(mod)
                    return mod
                except Exception:
                    return None
    return None


def tree_section(args, rank, world, local, dev, barrier, peak_tf):
    """Secondary metric of BASELINE.json: tree gravity particles/s on config 3 (NFW halo + exponential disc,
    per-particle spline softening, theta 0.7, order 3, leaf 8; bench_gravity.py shape = construct + potentials).
    Sources replicated on every rank (same seed), targets sharded; every rank builds the identical tree."""
    import torch
    import torch.distributed as dist

    from benchmarks.synthetic import nfw_disc
    from pynbodyext.gravity import Gravity, KernelKind
    from pynbodyext.gravity import device as gdev
    from pynbodyext.gravity.sharded import shard_bounds

    n, theta, order, leaf = args.tree_n, 0.7, 3, 8
    pos, mass, h = nfw_disc(n, seed=3)
    b = shard_bounds(n, world)
    lo, hi = b[rank], b[rank + 1]
    d_pos, d_mass, d_h = (torch.from_numpy(a).to(dev) for a in (pos, mass, h))

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def timed(fn, reps):
        out = []
        for _ in range(reps):
            barrier()
            e0, e1 = ev(), ev()
            e0.record()
            r = fn()
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out.append(float(t[0]))
        return min(out), r

    for _ in range(2):  # warm the memory pool
        t = gdev.OctreeDevice(d_pos, d_mass, leaf, order, d_h, 1)
        t.eval(theta, 1, shard=(rank, world))
        del t
    build_ms, tree = timed(lambda: gdev.OctreeDevice(d_pos, d_mass, leaf, order, d_h, 1), args.steps)
    walk_pot_ms, _ = timed(lambda: tree.eval(theta, 1, shard=(rank, world), kernel_events=True), args.steps)
    k_pot_ms = gdev.last_kernel_ms()
    walk_acc_ms, _ = timed(lambda: tree.eval(theta, 2, shard=(rank, world), kernel_events=True), args.steps)
    k_acc_ms = gdev.last_kernel_ms()

    def construct_and_pot():
        tt = gdev.OctreeDevice(d_pos, d_mass, leaf, order, d_h, 1)
        return tt.eval(theta, 1, shard=(rank, world))

    both_ms, _ = timed(construct_and_pot, args.steps)
    cnt = tree.walk_counters(theta, shard=(rank, world))
    info = tree.info()
    m_t = gdev.shard_count(n, world, rank)
    flop_acc = cnt["accepts"] * 140.0 + cnt["leaf_particles"] * 20.0 + cnt["visits"] * 10.0
    flop_pot = cnt["accepts"] * 120.0 + cnt["leaf_particles"] * 20.0 + cnt["visits"] * 10.0
    res = {
        "workload": f"NFW halo (c=10) + exponential disc N={n}, seed 3, per-particle spline softening, theta={theta}, "
                    f"multipole_order={order}, leaf_capacity={leaf} (BASELINE.json configs[2])",
        "metric": "tree_particles_per_s (construct + potentials, device-resident, bench_gravity.py shape)",
        "value": n / (both_ms * 1e-3), "unit": "particles/s", "n_gpus": world,
        "construct_ms": build_ms, "walk_pot_ms": walk_pot_ms, "walk_acc_ms": walk_acc_ms,
        "construct_plus_pot_ms": both_ms,
        "walk_only_particles_per_s": {"pot": n / (walk_pot_ms * 1e-3), "acc": n / (walk_acc_ms * 1e-3)},
        "nodes": info["n_nodes"], "depth": info["depth"],
        "per_target": {k: v / m_t for k, v in cnt.items() if k != "warp_visits"},
        "warp_union_visits_per_warp": cnt["warp_visits"] / max(1, (m_t + 31) // 32),
        "lane_visit_efficiency": cnt["visits"] / max(1, 32 * cnt["warp_visits"]),
        "roofline_walk": {
            "bound": "fp32", "kernel": "walk_kernel<3,acc,f32>", "unit": "TFLOP/s",
            "achieved": flop_acc / (k_acc_ms * 1e-3) / 1e12, "peak": peak_tf,
            "frac": flop_acc / (k_acc_ms * 1e-3) / 1e12 / peak_tf, "kernel_ms": k_acc_ms,
            "achieved_pot": flop_pot / (k_pot_ms * 1e-3) / 1e12, "kernel_ms_pot": k_pot_ms,
            "work_model": "accepts*140(acc)|120(pot) + leaf_pairs*20 + visits*10 flop (BASELINE.md §3), counts from "
                          "pnbx_tree_walk_counters (== oracle counters)",
            "interactions_per_s": (cnt["accepts"] + cnt["leaf_particles"]) / (k_acc_ms * 1e-3),
            "traffic": 1.448e9 if (world == 1 and n == 10_000_000) else None,
            "traffic_source": "from_profile",
            "traffic_note": "dram read+write of one N=1e7 potentials launch from ncu --set full (profiles/, walk kernel "
                            "summary; not re-measured in this run); the kernel is issue bound, not HBM bound",
        },
        "roofline_build": {
            "bound": "hbm", "unit": "GB/s", "achieved": 0.53e3 * n / (build_ms * 1e-3) / 1e9,
            "peak": peaks().get("hbm_gbs", 6650.0), "traffic": BUILD_TRAFFIC_1E7 if n == 10_000_000 else None,
            "traffic_source": "from_profile",
            "frac": 0.53e3 * n / (build_ms * 1e-3) / 1e9 / peaks().get("hbm_gbs", 6650.0),
            "algorithmic_bytes_per_particle": 530,
        },
    }
    if rank == 0 and world == 1:
        # e2e through the drop-in API with HOST arrays (H2D of 40 B/particle and D2H inside the timed region): ordinary
        # pageable numpy arrays are the headline (what a caller holds), pinned arrays the second number
        def pin(a):
            return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

        def e2e_run(pos_h, mass_h, h_h, label):
            for _ in range(2):  # warm-up; one tree alive at a time, like a user holding one Gravity object
                g = Gravity(pos_h, mass_h, softening=h_h, kernel=KernelKind.Spline)
                g.tree_potentials(theta=theta)
                del g
            reps = 3
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                ta = time.perf_counter()
                g = Gravity(pos_h, mass_h, softening=h_h, kernel=KernelKind.Spline)
                _ = g.tree
                tb = time.perf_counter()
                out = g.tree_potentials(theta=theta)
                tc = time.perf_counter()
                del g
                print(f"[bench] tree e2e ({label}): construct {1e3 * (tb - ta):.1f} ms, potentials {1e3 * (tc - tb):.1f} ms, "
                      f"free {1e3 * (time.perf_counter() - tc):.1f} ms", file=sys.stderr)
            dt = (time.perf_counter() - t0) / reps
            return {"value": n / dt, "unit": "particles/s", "ms_per_step": dt * 1e3,
                    "h2d_bytes_per_step": int(pos.nbytes + mass.nbytes + h.nbytes), "d2h_bytes_per_step": int(out.nbytes),
                    "api": "Gravity(pos, mass, softening=h, kernel=Spline).tree_potentials(theta=0.7) (construct + walk), "
                           + label}
        res["e2e"] = e2e_run(pos, mass, h, "pageable numpy arrays")
        res["e2e_pinned"] = e2e_run(pin(pos), pin(mass), pin(h), "pinned host arrays")
    del tree
    return res


def tree_1e8_section(args, rank, world, local, dev, barrier, gloo, peak_tf):
    """BASELINE configs 4 and 5: dm/gas/star zoom set N = 1e8 (deterministic zoom_range: the same global set at any
    world size), spline softening, theta 0.7, order 3, leaf 8, sharded over the GPUs of the box; 1e6 (R,z) grid targets
    from the same sources. Device-timed stages (max over ranks), end to end from pinned host shards to pinned host
    results, with IN-RUN parity: fp32 vs float64 walk on 1e4 self targets (same interaction lists) <= 1e-5, and tree
    (float64) vs float64 DIRECT sum at 1e4 particle positions and at 1e4 grid points within the theta = 0.7
    truncation bound. Then the same workload through the unchanged host API from ONE process (PNBX_DEVICES)."""
    import torch
    import torch.distributed as dist

    from benchmarks.synthetic import rz_grid_targets, zoom_range
    from pynbodyext.gravity import device as gdev
    from pynbodyext.gravity.sharded import pack_shard, replicate_sources, shard_bounds

    n, theta, order, leaf = args.tree1e8_n, 0.7, 3, 8
    b = shard_bounds(n, world)
    lo, hi = b[rank], b[rank + 1]
    per = max(b[r + 1] - b[r] for r in range(world))
    t0 = time.perf_counter()
    pos, mass, h, _fam = zoom_range(n, lo, hi, seed=4)
    rows_h = torch.from_numpy(pack_shard(pos, mass, h, 0, hi - lo, per)).pin_memory()
    gen_s = time.perf_counter() - t0

    def stage(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), r

    m_r = gdev.shard_count(n, world, rank)
    cap = per + gdev.SHARD_BLOCK
    pot_h = torch.empty(cap, dtype=torch.float64).pin_memory()
    acc_h = torch.empty((cap, 3), dtype=torch.float64).pin_memory()
    idx_h = torch.empty(cap, dtype=torch.int64).pin_memory()
    best = None
    tree = d_pos = d_mass = d_h = None
    for rep in range(2):
        del tree, d_pos, d_mass, d_h
        t_h2d, rows = stage(lambda: rows_h.to(dev, non_blocking=True))
        t_gather, allrows = stage(lambda: replicate_sources(rows, b))
        t_split, (d_pos, d_mass, d_h) = stage(lambda: (allrows[:, 0:3].contiguous(), allrows[:, 3].contiguous(),
                                                       allrows[:, 4].contiguous()))
        del allrows, rows
        t_build, tree = stage(lambda: gdev.OctreeDevice(d_pos, d_mass, leaf, order, d_h, 1))
        t_pot, pot = stage(lambda: tree.eval(theta, 1, shard=(rank, world), kernel_events=True)[0])
        k_pot = gdev.last_kernel_ms()
        t_acc, acc = stage(lambda: tree.eval(theta, 2, shard=(rank, world), kernel_events=True)[1])
        k_acc = gdev.last_kernel_ms()
        t_d2h, _ = stage(lambda: (pot_h[:m_r].copy_(pot, non_blocking=True), acc_h[:m_r].copy_(acc, non_blocking=True),
                                  idx_h[:m_r].copy_(tree.order(shard=(rank, world)), non_blocking=True)))
        res = {"h2d_ms": t_h2d, "allgather_ms": t_gather, "unpack_ms": t_split, "build_ms": t_build, "walk_pot_ms": t_pot,
               "walk_acc_ms": t_acc, "d2h_ms": t_d2h, "walk_pot_kernel_ms": k_pot, "walk_acc_kernel_ms": k_acc}
        res["total_pot_ms"] = t_h2d + t_gather + t_split + t_build + t_pot
        res["total_ms"] = res["total_pot_ms"] + t_acc + t_d2h
        if best is None or res["total_ms"] < best["total_ms"]:
            best = res
        finite = bool(torch.isfinite(pot).all() and torch.isfinite(acc).all())
        del pot, acc
    info = tree.info()
    cnt = tree.walk_counters(theta, shard=(rank, world))
    tcnt = torch.tensor([cnt[k] for k in ("visits", "accepts", "leaf_visits", "leaf_particles")], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tcnt)
    visits, accepts, leaf_visits, leaf_pairs = (float(x) for x in tcnt)
    flop_acc = accepts * 140.0 + leaf_pairs * 20.0 + visits * 10.0

    def rms(x):
        return float(x.pow(2).mean().sqrt())

    # ---- check 1 (rank 0's copy of the tree): fp32 vs float64 arithmetic on the same interaction lists
    c_lo, c_n = n // 2, 10_000
    p32, a32 = tree.eval(theta, 3, tgt_begin=c_lo, count=c_n)
    p64, a64 = tree.eval(theta, 3, tgt_begin=c_lo, count=c_n, precision="f64")
    os.environ["PNBX_WPT_MAX_TARGETS"] = "0"  # the whole-array API call below runs the lane-per-target kernel
    p1_ref = tree.eval(theta, 1, tgt_begin=c_lo, count=c_n)[0].cpu().numpy()  # potentials-only kernel (what the API leg runs)
    del os.environ["PNBX_WPT_MAX_TARGETS"]
    chk_fp = {"rms_rel_pot": rms((p32 - p64) / p64), "rms_rel_acc": rms((a32 - a64).norm(dim=1) / a64.norm(dim=1)),
              "targets": c_n, "tolerance": 1e-5}
    chk_fp["ok"] = bool(chk_fp["rms_rel_pot"] < 1e-5 and chk_fp["rms_rel_acc"] < 1e-5)

    # ---- check 2: tree (f64) vs float64 direct sum, at-points, targets sharded over the ranks
    def tree_vs_direct(q_all):
        qb = shard_bounds(q_all.shape[0], world)
        q = q_all[qb[rank]:qb[rank + 1]].contiguous()
        pt = tree.eval(theta, 1, targets=q, precision="f64")[0]
        pd = gdev.direct_device(d_pos, d_mass, d_h, kernel=1, want=1, targets=q, precision="f64")[0]
        err = ((pt - pd) / pd).abs()
        if world > 1:
            parts = [torch.empty(qb[r + 1] - qb[r], dtype=torch.float64, device=dev) for r in range(world)]
            dist.all_gather(parts, err)
            err = torch.cat(parts)
        out = {"targets": int(err.numel()), "median": float(err.median()), "p90": float(err.quantile(0.9)), "max": float(err.max()),
               "bound": "median < 2e-4 and max < 1e-2 (theta 0.7, order 3; single_node.rs far-field bound 1e-2)"}
        out["ok"] = bool(out["median"] < 2e-4 and out["max"] < 1e-2)
        return out

    chk_self = tree_vs_direct(d_pos[c_lo:c_lo + c_n])

    # ---- config 5: 1e6 grid targets, target-sharded
    grid = rz_grid_targets(args.grid_targets, seed=5)
    # interleaved target shards: a contiguous slice of the log-spaced grid would give one rank all the small radii
    # (the dense core, many times the work per target)
    g_h = torch.from_numpy(np.ascontiguousarray(grid[rank::world])).pin_memory()
    gp_h = torch.empty(g_h.shape[0], dtype=torch.float64).pin_memory()

    def grid_e2e():
        d_t = g_h.to(dev, non_blocking=True)
        gp = tree.eval(theta, 1, targets=d_t, kernel_events=True)[0]
        gp_h.copy_(gp, non_blocking=True)
        return gp

    grid_e2e()
    t_grid, gp = stage(grid_e2e)
    kg = torch.tensor([gdev.last_kernel_ms()], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(kg, op=dist.ReduceOp.MAX)
    k_grid = float(kg[0])
    chk_grid = tree_vs_direct(torch.from_numpy(np.ascontiguousarray(grid[::100])).to(dev))
    out = None
    if rank == 0:
        out = {
            "workload": f"dm/gas/star zoom set N={n} (zoom_range seed 4, identical at every world size), spline "
                        f"softening (gas h ∝ spacing), theta={theta}, order={order}, leaf={leaf}; BASELINE.json configs[3]; "
                        f"grid: {grid.shape[0]} log-spaced (R,z) targets (configs[4])",
            "n_gpus": world, "n": n, "nodes": info["n_nodes"], "depth": info["depth"], "finite": finite,
            "host_generate_s_per_rank": gen_s, **best,
            "particles_per_s": n / (best["total_ms"] * 1e-3),
            "value": n / (best["total_ms"] * 1e-3), "unit": "particles/s",
            "metric": "tree_particles_per_s (pinned host shards -> potentials + accelerations in pinned host memory)",
            "e2e": {"value": n / (best["total_ms"] * 1e-3), "unit": "particles/s",
                    "h2d_bytes_per_step": int(n * 40), "d2h_bytes_per_step": int(n * 40),
                    "api": "one process per GPU: shard H2D + NCCL all-gather + build + block-cyclic walk + D2H "
                           "(pynbodyext.gravity.device / sharded)"},
            "per_target": {"visits": visits / n, "accepts": accepts / n, "leaf_visits": leaf_visits / n, "leaf_pairs": leaf_pairs / n},
            "roofline": {"bound": "fp32", "kernel": "walk_kernel<3,acc,f32,spline>", "unit": "TFLOP/s",
                         "achieved": flop_acc / world / (best["walk_acc_kernel_ms"] * 1e-3) / 1e12, "peak": peak_tf,
                         "frac": flop_acc / world / (best["walk_acc_kernel_ms"] * 1e-3) / 1e12 / peak_tf,
                         "work_model": "per GPU: accepts*140 + leaf_pairs*20 + visits*10 flop (BASELINE.md §3), counts summed "
                                       "over ranks / world", "traffic": None},
            "roofline_build": {"bound": "hbm", "unit": "GB/s", "achieved": 0.53e3 * n / (best["build_ms"] * 1e-3) / 1e9,
                               "peak": peaks().get("hbm_gbs", 6650.0),
                               "frac": 0.53e3 * n / (best["build_ms"] * 1e-3) / 1e9 / peaks().get("hbm_gbs", 6650.0),
                               "note": "replicated build: every GPU builds the whole tree"},
            "checks": {"fp32_vs_f64_same_lists": chk_fp, "tree_vs_direct_f64_particle_positions": chk_self,
                       "tree_vs_direct_f64_grid_points": chk_grid,
                       "all_ok": bool(chk_fp["ok"] and chk_self["ok"] and chk_grid["ok"])},
            "grid": {"targets": int(grid.shape[0]), "e2e_ms": t_grid, "walk_kernel_ms": k_grid,
                     "targets_per_s": grid.shape[0] / (t_grid * 1e-3),
                     "api": "pinned host targets -> H2D -> tree.eval(targets) -> pinned host potentials; targets interleaved "
                            "over the ranks (rank r takes grid[r::world]); walk_kernel_ms = max over ranks"},
        }
    del tree, d_pos, d_mass, d_h, gp
    torch.cuda.empty_cache()
    gdev.trim_memory()  # hand the library's cached blocks back before another process builds the same tree on this GPU

    # ---- the same workload through the UNCHANGED host API from one process (rank 0 drives all GPUs, PNBX_DEVICES)
    if world > 1 and not args.no_api_1e8:
        full = []
        for a in (pos, mass, h):  # gather the per-rank shards of the host arrays on rank 0 (gloo, equal padded sizes)
            pad = np.zeros((per,) + a.shape[1:])
            pad[:hi - lo] = a
            t_sh = torch.from_numpy(pad)
            if rank == 0:
                parts = [torch.empty_like(t_sh) for _ in range(world)]
                dist.gather(t_sh, parts, dst=0, group=gloo)
                full.append(np.concatenate([parts[r].numpy()[:b[r + 1] - b[r]] for r in range(world)]))
                del parts
            else:
                dist.gather(t_sh, None, dst=0, group=gloo)
        if rank == 0:
            from pynbodyext.gravity import Gravity, KernelKind
            os.environ["PNBX_DEVICES"] = ",".join(str(i) for i in range(world))
            try:
                fpos, fmass, fh = full
                ms = []
                for rep in range(2):
                    ta = time.perf_counter()
                    g = Gravity(fpos, fmass, softening=fh, kernel=KernelKind.Spline)
                    _ = g.tree
                    tb = time.perf_counter()
                    p_api = g.tree_potentials(theta=theta)
                    tc = time.perf_counter()
                    a_api = g.tree_accelerations(theta=theta)
                    td = time.perf_counter()
                    pg_api = g.tree_potentials(positions=grid, theta=theta)  # config 5 through the same object
                    te = time.perf_counter()
                    del g
                    ms.append({"construct_ms": 1e3 * (tb - ta), "potentials_ms": 1e3 * (tc - tb),
                               "accelerations_ms": 1e3 * (td - tc), "total_ms": 1e3 * (td - ta),
                               "grid_potentials_ms": 1e3 * (te - td), "grid_finite": bool(np.isfinite(pg_api).all())})
                bapi = min(ms, key=lambda x: x["total_ms"])
                same = bool(np.array_equal(p_api[c_lo:c_lo + c_n], p1_ref))
                out["api_e2e"] = {**bapi, "value": n / (bapi["total_ms"] * 1e-3), "unit": "particles/s",
                                  "h2d_bytes_per_step": int(n * 40), "d2h_bytes_per_step": int(n * 32),
                                  "finite": bool(np.isfinite(p_api).all() and np.isfinite(a_api).all()),
                                  "bit_equal_to_sharded_run_on_check_targets": same,
                                  "api": "Gravity(pos, mass, softening=h, kernel=Spline): .tree (construct), "
                                         ".tree_potentials(), .tree_accelerations() — pageable numpy arrays, one process, "
                                         f"PNBX_DEVICES={os.environ['PNBX_DEVICES']}"}
            finally:
                os.environ.pop("PNBX_DEVICES", None)
        dist.barrier(group=gloo)
    return out


def tree_cpu_baseline(args):
    """CPU oracle on a bounded sample of the tree workload (the reference's build is serial: ~0.6 s per 1e6)."""
    from benchmarks.synthetic import nfw_disc
    from oracle import oracle as O
    try:
        O.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        pass
    ns = min(args.tree_n, 1_000_000)
    pos, mass, h = nfw_disc(ns, seed=3)
    t0 = time.perf_counter()
    ot = O.Tree(pos, mass, 8, 3, h, 1)
    tb = time.perf_counter() - t0
    t0 = time.perf_counter()
    ot.eval(0.7, want=1)
    tw = time.perf_counter() - t0
    return {"value": ns / (tb + tw), "unit": "particles/s", "cores": O.num_threads(), "kind": "port",
            "sample": f"NFW+disc N={ns} (same generator, seed 3): construct {tb:.2f} s (serial, as the reference) + "
                      f"potentials {tw:.2f} s (OpenMP), theta 0.7, order 3, leaf 8"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from benchmarks.synthetic import hernquist
    pos, mass = hernquist(args.n, seed=2)
    n_int = args.n * (args.n - 1)
    ref = real_reference()
    per_step = max(2.0, min(12.0, 60.0 / max(1, args.steps + args.warmup)))
    rates = []
    sample = ""
    cores = 1
    kind = "port"
    note = ("CPU oracle (C++/OpenMP port of direct.rs; no Rust toolchain in this image, baseline/_ref absent); "
            "ms_per_step extrapolated from the sampled rate to the full N(N-1) interactions")
    for i in range(args.warmup + args.steps):
        if ref is not None:  # the real Rust extension (gravity.rs:514-582): same bounded at-points sample
            kind = "reference"
            cores = len(os.sched_getaffinity(0))
            m = max(512, min(args.n, int(2.0e8 * cores * per_step / args.n)))
            idx = np.linspace(0, args.n - 1, m).astype(np.int64)
            h = np.full(args.n, EPS)
            t0 = time.perf_counter()
            ref.direct_accelerations_at_points_py(pos, np.ascontiguousarray(pos[idx]), mass, 0, h, 0)
            dt = time.perf_counter() - t0
            r, sample = m * args.n / dt, f"{m} of {args.n} targets (evenly strided), all sources, pynbodyext._rust from baseline/_ref, {dt:.1f} s"
            note = "the reference's own Rust extension (baseline/_ref), rayon on all host cores"
        else:
            r, cores, sample = cpu_reference_rate(pos, mass, per_step)
        if i >= args.warmup:
            rates.append(r)
    v = float(np.mean(rates)) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": n_int / (v * 1e9) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.n, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, **oracle_build_info()},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "parity_pin": "oracle-only" if ref is None else "reference",
        "note": note,
    }
    emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from benchmarks.synthetic import hernquist
    from pynbodyext.gravity import Gravity, KernelKind
    from pynbodyext.gravity import device as gdev

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        gloo = dist.new_group(backend="gloo")  # host-side barrier / gathers: waiting ranks must not spin on their GPUs
    n = args.n
    pos, mass = hernquist(n, seed=2)

    # ---- shards: rank r owns sources/targets [lo, hi); sources are replicated by one all-gather per step
    bounds = [(n * r) // world for r in range(world + 1)]
    lo, hi = bounds[rank], bounds[rank + 1]
    cnt = hi - lo
    per = max(bounds[r + 1] - bounds[r] for r in range(world))
    h_full = np.full(n, EPS)

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    if world > 1:
        # padded equal-size shards packed as (x,y,z,m,h) rows so a single all_gather replicates everything
        shard = np.zeros((per, 5))
        shard[:cnt, 0:3] = pos[lo:hi]
        shard[:cnt, 3] = mass[lo:hi]
        shard[:cnt, 4] = EPS
        d_shard = to_dev(shard)
        d_all = torch.empty((world * per, 5), dtype=torch.float64, device=dev)
    d_pos = to_dev(pos)
    d_mass = to_dev(mass)
    d_h = to_dev(h_full)

    def gather_sources():
        if world == 1:
            return d_pos, d_mass, d_h
        dist.all_gather_into_tensor(d_all, d_shard)
        if per * world == n:
            rows = d_all
        else:
            rows = torch.cat([d_all[r * per: r * per + (bounds[r + 1] - bounds[r])] for r in range(world)])
        return rows[:, 0:3].contiguous(), rows[:, 3].contiguous(), rows[:, 4].contiguous()

    def step(record=False):
        p, m_, h_ = gather_sources()
        _, acc = gdev.direct_device(p, m_, h_, kernel=0, want=2, tgt_begin=lo, count=cnt, kernel_events=record)
        return acc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = gdev.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    ev[0].record()
    for _ in range(args.steps):
        acc = step(record=True)
    ev[1].record()
    barrier()
    ms_total = ev[0].elapsed_time(ev[1])
    clocks = sampler.stop() if rank == 0 else None
    launches = gdev.launch_count() - launches0
    k_ms = gdev.last_kernel_ms()  # dominant kernel, last step (events on the launching stream)
    t = torch.tensor([ms_total, k_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, k_ms = float(t[0]), float(t[1])
    ms_step = ms_total / args.steps
    n_int = float(n) * float(n - 1)
    value = n_int / (ms_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (direct_kernel), per GPU
    sm_max = peaks().get("sm_max_mhz", 1965.0)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    nominal_tf = sms * 128 * 2 * sm_max * 1e6 / 1e12
    meas_tf = gdev.measure_fp32_peak(local, 0)
    meas_tf2 = gdev.measure_fp32_peak(local, 1)
    int_per_launch = float(cnt) * float(n - 1)
    achieved_tf = FLOP_PER_INTERACTION * int_per_launch / (k_ms * 1e-3) / 1e12
    # the same launch without the equal-mass specialisation (12 instead of 11 FP32-pipe ops per interaction)
    os.environ["PNBX_DIRECT_NO_CONSTM"] = "1"
    step(record=True)
    torch.cuda.synchronize()
    k_ms_general = gdev.last_kernel_ms()
    del os.environ["PNBX_DIRECT_NO_CONSTM"]
    # the softened variants on the same sources (per-particle Plummer and cubic spline, h in [0.005, 0.02])
    softened = {}
    if world == 1 and not args.no_softened:
        d_hv = to_dev(np.random.default_rng(0).uniform(0.005, 0.02, n))
        for name, kern in (("plummer_pair", 0), ("spline", 1)):
            for general in (False, True):  # the workload has equal masses: also time the general-mass variant
                if general:
                    os.environ["PNBX_DIRECT_NO_CONSTM"] = "1"
                for _ in range(2):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    gdev.direct_device(d_pos, d_mass, d_hv, kernel=kern, want=2, kernel_events=True)
                    e1.record()
                torch.cuda.synchronize()
                ms_v, call_ms = gdev.last_kernel_ms(), e0.elapsed_time(e1)
                os.environ["PNBX_DIRECT_SORT_MIN"] = "-1"  # the same sweep in the caller's particle order
                for _ in range(2):
                    gdev.direct_device(d_pos, d_mass, d_hv, kernel=kern, want=2, kernel_events=True)
                torch.cuda.synchronize()
                ms_unsorted = gdev.last_kernel_ms()
                del os.environ["PNBX_DIRECT_SORT_MIN"]
                os.environ.pop("PNBX_DIRECT_NO_CONSTM", None)
                softened[name + ("_general_mass" if general else "")] = {
                    "kernel_ms": ms_v, "achieved": FLOP_PER_INTERACTION * n_int / (ms_v * 1e-3) / 1e12,
                    "frac": FLOP_PER_INTERACTION * n_int / (ms_v * 1e-3) / 1e12 / meas_tf,
                    "call_ms": call_ms, "kernel_ms_caller_order": ms_unsorted,
                    "note": "whole-array self call: particles sorted by " + ("softening" if kern == 0 else "Morton key") +
                            " so that whole (target block, source tile) combinations resolve " +
                            ("max(h_i, h_j)" if kern == 0 else "the r < h test") + "; call_ms = device time of the "
                            "whole call (sort, packing, sweep, scatter back to the caller's order)"}
        del d_hv
    roofline = {
        "bound": "fp32", "kernel": "direct_kernel_f2<acc, const_mass> (packed FP32x2; the workload has equal masses)",
        "achieved": achieved_tf,
        "peak": meas_tf, "unit": "TFLOP/s", "frac": achieved_tf / meas_tf,
        "peak_source": "measured here: FFMA-chain microbenchmark (pnbx_measure_fp32_peak), MEASURED_PEAKS.json has no FP32 entry",
        "peak_nominal": nominal_tf, "frac_of_nominal": achieved_tf / nominal_tf, "peak_ffma2_chain": meas_tf2,
        "flop_per_interaction": FLOP_PER_INTERACTION, "interactions_per_launch": int_per_launch,
        "kernel_ms": k_ms, "ginteractions_per_s_kernel": int_per_launch / (k_ms * 1e-3) / 1e9,
        "general_mass_variant": {"kernel_ms": k_ms_general,
                                 "achieved": FLOP_PER_INTERACTION * int_per_launch / (k_ms_general * 1e-3) / 1e12,
                                 "frac": FLOP_PER_INTERACTION * int_per_launch / (k_ms_general * 1e-3) / 1e12 / meas_tf},
        "softened_variants": softened or None,
        "traffic": 172.3e6, "traffic_source": "from_profile",
        "traffic_note": "dram read+write per 1e12-interaction launch from ncu --set full (profiles/, direct kernel summary; "
                        "not re-measured in this run); algorithmic HBM bytes ~ 16 B per source per launch + 24 B per target result",
    }

    # ---- parity in the same run: fp32 GPU vs float64 oracle on a 256-target subsample (rank 0)
    parity = None
    cpu = None
    if rank == 0:
        from oracle import oracle as O
        idx = np.linspace(lo, hi - 1, 256).astype(np.int64)
        a_gpu = acc[torch.from_numpy(idx - lo).to(dev)].cpu().numpy()
        # oracle at-points includes the (exactly zero) self term for accelerations with eps > 0
        _, a_ref = O.direct(pos, mass, h_full, targets=np.ascontiguousarray(pos[idx]), kernel=0, want=2)
        parity = {"rms_rel_acc_vs_f64_oracle": float(np.sqrt((((a_gpu - a_ref) ** 2).sum(1) / (a_ref ** 2).sum(1)).mean())),
                  "targets": 256, "tolerance": 1e-5}

    # ---- e2e through the public drop-in API with HOST buffers (H2D + D2H inside the timed region).
    # Headline: ordinary pageable numpy arrays, Gravity(...).direct_accelerations(). N > 1: the same call from rank 0
    # alone with PNBX_DEVICES = all N GPUs (one host thread per GPU inside the library); the other ranks wait.
    def pinned(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    def time_api(pos_h, mass_h, e_steps):
        def call():
            return Gravity(pos_h, mass_h, softening=EPS, kernel=KernelKind.Plummer).direct_accelerations()
        call()
        call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            out = call()
        return (time.perf_counter() - t0) / e_steps, out

    e_steps = max(1, min(args.steps, 5))
    e2e = None
    e2e_extra = {}
    if world == 1:
        dt, out = time_api(pos, mass, e_steps)
        dt_pin, _ = time_api(pinned(pos), pinned(mass), e_steps)
        e2e = {"value": n_int / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(pos.nbytes + mass.nbytes + h_full.nbytes),
               "d2h_bytes_per_step": int(out.nbytes), "ms_per_step": dt * 1e3, "steps": e_steps,
               "api": "Gravity(pos, mass, softening=0.01, kernel=Plummer).direct_accelerations(), pageable numpy arrays"}
        e2e_extra["e2e_pinned"] = {"value": n_int / dt_pin / 1e9, "unit": UNIT, "ms_per_step": dt_pin * 1e3,
                                   "api": "the same call with pinned host arrays"}
    else:
        barrier()
        if rank == 0:
            os.environ["PNBX_DEVICES"] = ",".join(str(i) for i in range(world))
            try:
                dt, out = time_api(pos, mass, e_steps)
            finally:
                os.environ.pop("PNBX_DEVICES", None)
            e2e = {"value": n_int / dt / 1e9, "unit": UNIT,
                   "h2d_bytes_per_step": int(pos.nbytes + mass.nbytes + h_full.nbytes), "d2h_bytes_per_step": int(out.nbytes),
                   "ms_per_step": dt * 1e3, "steps": e_steps,
                   "api": "Gravity(pos, mass, softening=0.01, kernel=Plummer).direct_accelerations(), pageable numpy arrays, "
                          f"ONE process driving {world} GPUs (PNBX_DEVICES; per-GPU shard H2D, peer all-gather over NVLink, "
                          "per-GPU D2H into the caller's array); the other ranks idle at a host barrier"}
        dist.barrier(group=gloo)
        # second number: one process per GPU through the torch.distributed helper (per-rank shard H2D, NCCL all-gather)
        from pynbodyext.gravity.sharded import direct_sharded

        def sharded_step():
            return direct_sharded(pos, mass, h_full, kernel=0, want=2, rank=rank, world=world, device=local)[1]
        sharded_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            sharded_step()
        barrier()
        e_dt = torch.tensor([(time.perf_counter() - t0) / e_steps], dtype=torch.float64, device=dev)
        dist.all_reduce(e_dt, op=dist.ReduceOp.MAX)
        e2e_extra["e2e_one_process_per_gpu"] = {
            "value": n_int / float(e_dt[0]) / 1e9, "unit": UNIT, "ms_per_step": float(e_dt[0]) * 1e3,
            "api": "pynbodyext.gravity.sharded.direct_sharded (per-rank shard H2D, NCCL all-gather, D2H)"}

    tree = None
    if not args.no_tree:
        tree = tree_section(args, rank, world, local, dev, barrier, meas_tf)
    tree_1e8 = None
    if (world > 1 or args.tree1e8) and not args.no_tree1e8:
        del d_pos, d_mass, d_h
        torch.cuda.empty_cache()
        tree_1e8 = tree_1e8_section(args, rank, world, local, dev, barrier, gloo, meas_tf)

    # CPU legs last: the oracle's OpenMP team must not compete with the host side of the GPU measurements above
    if rank == 0 and world == 1 and not args.no_cpu:
        r, cores, sample = cpu_reference_rate(pos, mass, 12.0)
        cpu = {"value": r / 1e9, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, **oracle_build_info()}
        if tree is not None:
            tree["cpu_baseline"] = tree_cpu_baseline(args)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(n, world), "roofline": roofline, "cpu_baseline": cpu,
            "e2e": e2e, **e2e_extra, "gpu_launches": int(launches), "clocks": clocks, "parity": parity,
            "parity_pin": "reference source executed (no Rust toolchain in this image, so the reference binary cannot be "
                          "built; tests/golden/make_reference_exec.py translates the reference's own kernel.rs / direct.rs "
                          "/ multipole.rs / tree.rs walk functions to Python mechanically and runs them: "
                          "tests/golden/reference_exec.npz, 250 arrays, oracle bit-exact against them in "
                          "tests/test_oracle_vs_reference_source.py; ORACLE_REVIEW.md lists the line-by-line review)",
            "tflops_20flop": FLOP_PER_INTERACTION * value * 1e9 / 1e12, "tree": tree, "tree_1e8": tree_1e8,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", "--particles", dest="n", type=int, default=1_000_000)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-tree", action="store_true", help="skip the secondary tree-gravity section")
    ap.add_argument("--tree-n", type=int, default=10_000_000)
    ap.add_argument("--no-softened", action="store_true", help="skip the per-pair Plummer / spline direct kernel timings")
    ap.add_argument("--tree1e8", action="store_true", help="run the N=1e8 zoom tree section on one GPU too")
    ap.add_argument("--no-tree1e8", action="store_true", help="skip the N=1e8 zoom tree section (it runs by default for N > 1)")
    ap.add_argument("--tree1e8-n", type=int, default=100_000_000)
    ap.add_argument("--grid-targets", type=int, default=1_000_000)
    ap.add_argument("--no-api-1e8", action="store_true", help="skip the one-process PNBX_DEVICES leg of the N=1e8 section")
    args = ap.parse_args()
    # Exactly one JSON line may reach stdout: route everything else written to fd 1 (NCCL's version banner,
    # library chatter) to stderr and keep the real stdout for the result line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
