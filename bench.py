#!/usr/bin/env python
"""bench.py — headline benchmark of the gravity hot path on B200.

Metric (BASELINE.json): direct-summation Ginteractions/s on config 2 — Hernquist halo N = 1e6,
Plummer softening eps = 0.01, accelerations of all particles (reference call:
Gravity(pos, mass, softening=0.01, kernel=KernelKind.Plummer).direct_accelerations(), i.e.
direct.rs:443-524), fp32 interaction arithmetic checked against the float64 oracle.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n 1000000] [--impl ours|reference]
  torchrun --nproc-per-node N ... bench.py --gpus N ...     (one rank per GPU, targets sharded)

One JSON line on stdout (rank 0). See DESIGN.md §Measurement for every key.
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
            "traffic_note": "dram read+write of one N=1e7 potentials launch from ncu --set full "
                            "(profiles/r01_walk_kernel_ncu.md, final capture); the kernel is issue bound, not HBM bound",
        },
        "roofline_build": {
            "bound": "hbm", "unit": "GB/s", "achieved": 0.53e3 * n / (build_ms * 1e-3) / 1e9,
            "peak": peaks().get("hbm_gbs", 6650.0), "traffic": None,
            "frac": 0.53e3 * n / (build_ms * 1e-3) / 1e9 / peaks().get("hbm_gbs", 6650.0),
            "algorithmic_bytes_per_particle": 530,
        },
    }
    if rank == 0 and world == 1:
        # e2e through the drop-in API with host arrays (H2D of 40 B/particle and D2H inside the timed region)
        def pin(a):
            return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        pos_h, mass_h, h_h = pin(pos), pin(mass), pin(h)
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
            print(f"[bench] tree e2e: construct {1e3 * (tb - ta):.1f} ms, potentials {1e3 * (tc - tb):.1f} ms, "
                  f"free {1e3 * (time.perf_counter() - tc):.1f} ms", file=sys.stderr)
        dt = (time.perf_counter() - t0) / reps
        res["e2e"] = {"value": n / dt, "unit": "particles/s", "ms_per_step": dt * 1e3,
                      "h2d_bytes_per_step": int(pos.nbytes + mass.nbytes + h.nbytes), "d2h_bytes_per_step": int(out.nbytes),
                      "api": "Gravity(pos, mass, softening=h, kernel=Spline).tree_potentials(theta=0.7) (construct + walk), "
                             "pinned host arrays"}
    del tree
    return res


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
    rates = []
    per_step = max(2.0, min(12.0, 60.0 / max(1, args.steps + args.warmup)))
    sample = ""
    cores = 1
    for i in range(args.warmup + args.steps):
        r, cores, sample = cpu_reference_rate(pos, mass, per_step)
        if i >= args.warmup:
            rates.append(r)
    v = float(np.mean(rates)) / 1e9
    n_int = args.n * (args.n - 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": n_int / (v * 1e9) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.n, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU oracle (C++/OpenMP port of direct.rs; Rust toolchain absent); ms_per_step extrapolated "
                "from the sampled rate to the full N(N-1) interactions",
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
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

    kernel_ms = []

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
        kernel_ms.append(None)  # filled after the sync (events are read without stalling the loop)
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
        "traffic": 172.3e6,
        "traffic_note": "dram read+write per 1e12-interaction launch from ncu --set full (profiles/r01_direct_kernel_f2_ncu.md); "
                        "algorithmic HBM bytes ~ 16 B per source per launch + 24 B per target result",
    }

    # ---- parity in the same run: fp32 GPU vs float64 oracle on a 256-target subsample (rank 0)
    parity = None
    e2e = None
    cpu = None
    if rank == 0:
        from oracle import oracle as O
        idx = np.linspace(lo, hi - 1, 256).astype(np.int64)
        a_gpu = acc[torch.from_numpy(idx - lo).to(dev)].cpu().numpy()
        # oracle at-points includes the (exactly zero) self term for accelerations with eps > 0
        _, a_ref = O.direct(pos, mass, h_full, targets=np.ascontiguousarray(pos[idx]), kernel=0, want=2)
        parity = {"rms_rel_acc_vs_f64_oracle": float(np.sqrt((((a_gpu - a_ref) ** 2).sum(1) / (a_ref ** 2).sum(1)).mean())),
                  "targets": 256, "tolerance": 1e-5}

    # ---- e2e through the public drop-in API with pinned HOST buffers (H2D + D2H inside the timed region)
    def pinned(a):
        tt = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return tt.numpy()
    pos_h, mass_h = pinned(pos), pinned(mass)
    import pynbodyext._rust as backend

    def e2e_step():
        if world == 1:
            g = Gravity(pos_h, mass_h, softening=EPS, kernel=KernelKind.Plummer)
            return g.direct_accelerations()
        from pynbodyext.gravity.sharded import direct_sharded
        return direct_sharded(pos_h, mass_h, h_full, kernel=0, want=2, rank=rank, world=world, device=local)[1]

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e_steps = max(1, min(args.steps, 3))
    for _ in range(e_steps):
        out = e2e_step()
    barrier()
    e_dt = torch.tensor([(time.perf_counter() - t0) / e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e_dt, op=dist.ReduceOp.MAX)
    e2e = {"value": n_int / float(e_dt[0]) / 1e9, "unit": UNIT,
           "h2d_bytes_per_step": int(pos.nbytes + mass.nbytes + h_full.nbytes) if world == 1 else int(cnt * 40),
           "d2h_bytes_per_step": int(cnt * 24), "ms_per_step": float(e_dt[0]) * 1e3, "steps": e_steps,
           "api": "Gravity(...).direct_accelerations() with pinned host numpy arrays" if world == 1
                  else "pynbodyext.gravity.sharded.direct_sharded (per-rank shard H2D, NCCL all-gather, D2H)"}

    tree = None
    if not args.no_tree:
        tree = tree_section(args, rank, world, local, dev, barrier, meas_tf)

    # CPU legs last: the oracle's OpenMP team must not compete with the host side of the GPU measurements above
    if rank == 0 and world == 1 and not args.no_cpu:
        r, cores, sample = cpu_reference_rate(pos, mass, 12.0)
        cpu = {"value": r / 1e9, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        if tree is not None:
            tree["cpu_baseline"] = tree_cpu_baseline(args)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(n, world), "roofline": roofline, "cpu_baseline": cpu,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "parity": parity,
            "tflops_20flop": FLOP_PER_INTERACTION * value * 1e9 / 1e12, "tree": tree,
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
