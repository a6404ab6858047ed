#!/usr/bin/env python
"""bench.py — checked edge expansions/sec of the KGMT tree-expansion path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3]

A step = one complete KGMT plan (root insertion + every expansion iteration until the goal is reached, the tree is
full or the iteration limit hits) of BASELINE config 2: car, synthetic dense map of 1 000 obstacle AABBs,
2^20-node tree capacity, N=16 / n=32 region grid, step s planned with seed s+1.  One process per GPU; with N > 1 the
seeds (independent planning queries) are sharded over the ranks, no data-path collective (weak scaling).

  value   total expansions of the K timed plans / their device time (CUDA events on the planner's stream around
          reset + root kernel + the cooperative expansion kernel; obstacles, cull grid and tree storage resident in HBM)
  e2e     the same plans through the public C-ABI calls with HOST buffers every step: kgmt_set_obstacles_host
          (H2D of the obstacle set + cull grid) + kgmt_plan (host init/goal) + result block and solution path D2H,
          wall clock between device synchronisations
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement".

--impl reference times the reference's OWN propagate+collision loop (oracle/_ref/libref_host.so: its unmodified
statePropagator.cu + collisionCheck.cu built for the host) on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "checked edge expansions/sec"
UNIT = "expansions/s"
B_EXP, B_INS = 33.5, 77.0          # algorithmic HBM bytes per expansion / per accepted node


def workload(name):
    from cudasbmp_b200 import workloads as w
    if name == "c1":
        return dict(cfg=w.C1, obstacles=w.C1_OBSTACLES, init=w.C1_INIT, goal=w.C1_GOAL,
                    label="config1: reference demo map (5 AABBs), N=16 n=8, maxTree=30000, numDisc=10")
    if name == "c3":
        return dict(cfg=w.C3, obstacles=w.c3_obstacles(10000), init=w.C2_INIT, goal=w.C2_GOAL,
                    label="config3: car, 10000 synthetic AABBs, numDisc=40, N=16 n=32, maxTree=2^20")
    return dict(cfg=w.C2, obstacles=w.c2_obstacles(1000), init=w.C2_INIT, goal=w.C2_GOAL,
                label="config2: car KGMT single query, 1000 synthetic AABBs, N=16 n=32, maxTree=2^20, numDisc=10")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_rate(wl, seconds=12.0, threads=None):
    """The reference's own propagate+collision loop (host build) on a bounded sample of the workload."""
    from oracle import pyoracle as po
    from cudasbmp_b200 import workloads as w
    R = po.ref_host()
    kind = "reference"
    cfg, obs = wl["cfg"], wl["obstacles"]
    cores = threads or (R.ref_host_hw_threads() if R is not None else (os.cpu_count() or 1))
    parents = w.random_parents(4096, obs, seed=7)

    def run(M):
        pof = (np.arange(M, dtype=np.int32) // 32) % len(parents)
        if R is not None:
            sec, _, _, _ = po.ref_host_batch(parents, pof, 12345, 0, cfg["numDisc"], cfg["agentLength"], obs, cfg["width"],
                                             cfg["height"], threads=cores, outputs=False)
        else:
            t0 = time.perf_counter()
            po.propagate_batch(parents, pof, 12345, 0, cfg["numDisc"], cfg["agentLength"], obs, cfg["width"], cfg["height"],
                               po.MATH_HOST, want_margin=False)
            sec = time.perf_counter() - t0
        return sec

    if R is None:
        kind, cores = "port", 1
    M = 4096 * 8
    sec = run(M)
    rate = M / max(sec, 1e-9)
    M2 = int(min(max(rate * seconds, M), 64e6)) // 32 * 32
    sec2 = run(M2)
    return dict(value=M2 / sec2, unit=UNIT, cores=cores, kind=kind,
                sample="%d candidate edges from 4096 free-space parents x 32 children on the %s map, %.1f s wall"
                       % (M2, wl["label"].split(":")[0], sec2)), M2, sec2


def ref_cuda_rate(wl):
    """Baseline A of BASELINE.json: the reference's OWN propagateG kernel (unmodified sources recompiled for sm_100a,
    oracle/_ref/libref_gpu.so, cuRAND Philox states) on the same obstacle map, CUDA-event time of stages 2-5a for
    30 000 parents x 32 children (its largest launch, KGMT.cu:160-173)."""
    try:
        from oracle import pyoracle as po
        from cudasbmp_b200 import workloads as w
        if po.ref_gpu() is None:
            return {"value": None, "unit": UNIT, "kind": "unavailable", "sample": "oracle/_ref/libref_gpu.so missing"}
        cfg, obs = wl["cfg"], wl["obstacles"]
        P = 30000
        parents = w.random_parents(P, obs, seed=7)
        N, n = 16, 8
        c1, c2 = N * N, N * N * n * n
        maps = {k: np.zeros(c1 if k.startswith("R1") else c2, dtype=np.int32)
                for k in ("R1", "R2", "R1Valid", "R2Valid", "R1Invalid", "R2Invalid", "R1Avail", "R2Avail")}
        _, _, _, ms = po.ref_gpu_expand(1, 32, parents, np.arange(P, dtype=np.int32), maps, np.ones(c1, np.float32), N, n,
                                        cfg["width"] / N, cfg["width"] / (N * n), cfg["numDisc"], cfg["agentLength"], obs,
                                        cfg["width"], cfg["height"], 99, reps=5)
        return {"value": P * 32 / ms * 1e3, "unit": UNIT, "kind": "reference CUDA kernel propagateG recompiled for sm_100a",
                "sample": "%d parents x 32 children on the %s map, %.3f ms per launch (best of 5)" % (P, wl["label"].split(":")[0], ms)}
    except Exception as e:
        return {"value": None, "unit": UNIT, "kind": "unavailable", "sample": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collide", default="grid", choices=["grid", "brute"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = workload(args.workload)
    config = {"workload": wl["label"], "step": "one complete plan (all expansion iterations), seed = step index + 1",
              "sharding": "independent queries/seeds per GPU, no data-path collective", "collision": args.collide,
              "l2": "flushed between steps (256 MiB device write)"}

    # ------------------------------------------------------------------ reference arm (CPU; rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        steps, rates = max(args.steps, 1), []
        per_step = max(2.0, min(12.0, 120.0 / (steps + args.warmup)))
        for _ in range(args.warmup):
            cpu_reference_rate(wl, seconds=min(per_step, 2.0))
        tot_M, tot_s, base = 0, 0.0, None
        for _ in range(steps):
            base, M, sec = cpu_reference_rate(wl, seconds=per_step)
            tot_M += M; tot_s += sec
        val = tot_M / tot_s
        base["value"] = val
        print(json.dumps({"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / steps, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "impl": "reference", "config": config, "cpu_baseline": base,
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import cudasbmp_b200 as k
    from cudasbmp_b200 import kgmt as K
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = wl["cfg"]
    mode = K.COLLIDE_GRID if args.collide == "grid" else K.COLLIDE_BRUTE
    plan = k.KGMT(**cfg, seed=1, device=local, collision_mode=mode)
    plan.set_obstacles(wl["obstacles"])
    L = k.load()
    import ctypes as C
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def query(step):
        plan.set_seed(1 + step * world + rank)          # independent query per step and rank
        return wl["init"]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run_resident(step):
        flush.fill_(step & 0xFF)
        torch.cuda.synchronize()
        return plan.plan(query(step), wl["goal"])

    for s in range(args.warmup):
        run_resident(s)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = plan.launch_count
    t0 = time.perf_counter()
    dev_ms, results = 0.0, []
    for s in range(args.steps):
        r = run_resident(args.warmup + s)
        dev_ms += r["device_ms"]
        results.append(r)
    barrier()
    wall = time.perf_counter() - t0
    launches = plan.launch_count - launches0
    expansions = sum(r["expansions"] for r in results)
    accepted = sum(r["tree_size"] - 1 for r in results)

    # ---- end to end through the C ABI with host buffers
    obs_host = np.ascontiguousarray(wl["obstacles"], dtype=np.float32)
    for s in range(min(args.warmup, 3)):
        plan.set_obstacles(obs_host); plan.plan(query(s), wl["goal"])
    barrier()
    e2e_exp, d2h = 0, 0
    t1 = time.perf_counter()
    for s in range(args.steps):
        plan.set_obstacles(obs_host)
        r = plan.plan(query(args.warmup + s), wl["goal"])
        e2e_exp += r["expansions"]
        d2h += 128
        if r["stop"] == 1:
            path = plan.extract_path()
            d2h += 128 + 4 + len(path) * 28          # state block + length + the AoS-7 rows
    barrier()
    e2e_s = time.perf_counter() - t1
    cfgd = plan.config()
    h2d = obs_host.nbytes + cfgd["cull_items"] * 16 + (cfgd["cull_cells"] ** 2 + 4) * 4 + 56 + 128
    clocks = sampler.stop() if rank == 0 else None

    # ---- median time-to-first-solution, BASELINE config 1 (reference demo), 101 seeds: host wall clock from the
    #      kgmt_plan call (state allocated, obstacles resident) to the host holding costToGoal != 0
    ttfs = None
    if rank == 0:
        from cudasbmp_b200 import workloads as w
        p1 = k.KGMT(**w.C1, seed=1, device=local)
        p1.set_obstacles(w.C1_OBSTACLES)
        for s in range(3):
            p1.set_seed(1000 + s); p1.plan(w.C1_INIT, w.C1_GOAL)
        walls, devs, unsolved = [], [], 0
        for s in range(1, 102):
            p1.set_seed(s)
            tq = time.perf_counter()
            r = p1.plan(w.C1_INIT, w.C1_GOAL)
            dtq = time.perf_counter() - tq
            if r["stop"] == 1:
                walls.append(dtq * 1e3); devs.append(r["device_ms"])
            else:
                unsolved += 1
        if walls:
            ws = sorted(walls)
            ttfs = {"config": "config1: reference demo map, init (5,5) goal (2,18), maxTree 30000", "seeds": 101,
                    "solved": len(walls), "median_ms": statistics.median(walls), "p95_ms": ws[int(0.95 * (len(ws) - 1))],
                    "device_median_ms": statistics.median(devs), "clock": "host wall around kgmt_plan"}
        p1.close()

    # ---- max over ranks, sum of work
    t = torch.tensor([dev_ms, e2e_s, wall], dtype=torch.float64, device="cuda")
    c = torch.tensor([expansions, e2e_exp, accepted, launches], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_s_max, wall_max = (float(v) for v in t.tolist())
    exp_all, e2e_all, acc_all, launches_all = (float(v) for v in c.tolist())
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    value = exp_all / (dev_ms_max * 1e-3)
    alpha = acc_all / max(exp_all, 1.0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # dominant kernel = expand_kernel (one cooperative launch per plan); its duration ~ the plan's device time
    # minus the reset memsets and the root kernel, measured live below with its own events
    per_launch_exp = exp_all / world / args.steps
    kern_ms = dev_ms / args.steps
    achieved = per_launch_exp * (B_EXP + B_INS * alpha) / (kern_ms * 1e-3) / 1e9
    solved = [r for r in results if r["stop"] == 1]
    # ncu-derived per-launch figures of the dominant kernel (DRAM bytes, warp instructions per expansion)
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_inputs.json")))
        if prof.get("workload") != args.workload or prof.get("collision") != args.collide:
            prof = {}
    except Exception:
        prof = {}
    sm_clock_hz = 1e6 * float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
    issue_peak = cfgd["sms"] * 4 * sm_clock_hz                 # one warp instruction per scheduler per clock
    ipe = prof.get("warp_instructions_per_expansion")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "e2e": {"value": e2e_all / e2e_s_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h / max(args.steps, 1)),
                "ms_per_step": 1e3 * e2e_s_max / args.steps},
        "gpu_launches": int(launches_all),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": prof.get("dram_bytes_per_launch"), "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                     "kernel": "kgmt::expand_kernel<grid|brute, LOOP> (cooperative, one launch per plan)",
                     "algorithmic_bytes_per_expansion": B_EXP + B_INS * alpha, "accept_ratio": alpha,
                     "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, %s)" % prof.get("source"),
                     "algorithmic_bytes_per_launch": per_launch_exp * (B_EXP + B_INS * alpha),
                     "note": "stages 2-4 are FP32-issue bound, not HBM bound (DESIGN.md); see roofline_issue and profiles/"},
        "roofline_issue": {"bound": "warp-instruction issue", "unit": "G warp-inst/s",
                           "achieved": (per_launch_exp * ipe / (kern_ms * 1e-3) / 1e9) if ipe else None,
                           "peak": issue_peak / 1e9, "frac": (per_launch_exp * ipe / (kern_ms * 1e-3) / issue_peak) if ipe else None,
                           "warp_instructions_per_expansion": ipe, "avg_active_lanes": prof.get("avg_active_lanes"),
                           "peak_source": "SMs x 4 schedulers x SM clock under load (nvidia-smi during the timed region)",
                           "source": prof.get("source")},
        "plan": {"expansions_per_plan": exp_all / world / args.steps, "tree_size_mean": acc_all / world / args.steps + 1,
                 "iterations_mean": statistics.mean(r["iterations"] for r in results),
                 "solved": len(solved), "stops": sorted(set(r["stop"] for r in results)),
                 "time_to_first_solution_ms_median": statistics.median(r["device_ms"] for r in solved) if solved else None,
                 "host_wall_ms_per_plan": 1e3 * wall_max / args.steps},
        "collide_backend": cfgd,
    }
    line["ttfs"] = ttfs
    if not args.no_cpu_baseline:
        line["ref_cuda_baseline"] = ref_cuda_rate(wl)
    if not args.no_cpu_baseline:
        try:
            base, _, _ = cpu_reference_rate(wl, seconds=12.0)
            line["cpu_baseline"] = base
        except Exception as e:       # the checker's absence must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
