#!/usr/bin/env python
"""bench.py — checked edge expansions/sec of the KGMT tree-expansion path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--skip c3,c4,c5,ttfs,...]

Headline (every N): BASELINE config 2 — car, synthetic dense map of 1 000 obstacle AABBs, 2^20-node tree capacity,
N=16 / n=32 region grid.  A STEP = one batch of PLANS_PER_STEP complete KGMT plans (root insertion + every expansion
iteration until the goal is reached, the tree is full or the iteration limit hits), plan j of step s planned with its
own seed.  One process per GPU; with N > 1 the seeds (independent planning queries) are sharded over the ranks, no
data-path collective (weak scaling).

  value   total expansions of the timed plans / their device time (CUDA events on the planner's stream around
          reset + root kernel + the cooperative expansion kernel; obstacles, cull grid and tree storage resident in
          HBM; L2 flushed before every plan, outside the events)
  e2e     the same plans through the public C-ABI calls with HOST buffers: kgmt_set_obstacles_host (H2D of the
          obstacle set + cull grid) + kgmt_plan (host init/goal) + result block and solution path D2H, wall clock
          between device synchronisations
  roofline / roofline_issue / roofline_fp32 / cpu_baseline / clocks / gpu_launches: DESIGN.md "Measurement".

Further keys of the same JSON line measure the other named configurations (VERDICT r01 "next" 2):
  ttfs        median time-to-first-solution on config 1 and config 2 (101 seeds), and — N = 1 — the reference's own
              plan() on config 1 (ttfs.reference_ms)
  ttfs_multi  N > 1: the same as a portfolio race over peer memory (one seed per GPU, first solution stops the rest)
  c3          N = 1: config 3 (10 000 obstacles, numDisc 40) with the culled and the TMA-streamed exhaustive back end
  c4          config 4: the FIXED batch of 1 024 queries sharded over the N ranks (strong scaling), kgmt_plan_batch
  c5          config 5: one sharded iteration of M = 2^20 .. 2^26 candidates, compute and exchange time
  same_population   stages 2-4 on IDENTICAL inputs (frontier nodes sampled from a real config-2 plan x 32 children,
              same Philox streams): this library, the reference's CUDA kernel, the reference's host loop

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
STATE_BYTES = 512                  # sizeof(kgmt::DevState): the planner's scalar block read back after every plan
PLANS_PER_STEP = 64                # a step = one batch of this many complete plans (>= 1 s timed at the driver's 20 steps)
SAMPLE = os.path.join(ROOT, "bench_data", "c2_frontier_sample.npz")


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


def sample_parents(wl):
    """Parents of the bounded samples: frontier nodes of a REAL config-2 plan (bench_data/, made on a B200 by
    scripts/dump_frontier_sample.py) — the same candidate population the GPU arm expands; free-space random parents
    only when the fixture is missing or for another map."""
    from cudasbmp_b200 import workloads as w
    if wl["label"].startswith("config2") and os.path.exists(SAMPLE):
        d = np.load(SAMPLE)
        return np.ascontiguousarray(d["parents"], dtype=np.float32), "frontier nodes of a config-2 plan (seed %d, iterations %s)" % (
            int(d["seed"]), ",".join(str(int(i)) for i in d["iterations"]))
    return w.random_parents(4096, wl["obstacles"], seed=7), "random free-space parents"


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
        sm, smax, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); pw.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_mhz_min": min(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_cpu():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return model, os.cpu_count() or 1


def cpu_reference_rate(wl, seconds=12.0, threads=None):
    """The reference's own propagate+collision loop (host build) on a bounded sample of the workload."""
    from oracle import pyoracle as po
    R = po.ref_host()
    kind = "reference"
    cfg, obs = wl["cfg"], wl["obstacles"]
    cores = threads or (R.ref_host_hw_threads() if R is not None else (os.cpu_count() or 1))
    parents, what = sample_parents(wl)

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
    model, ncpu = host_cpu()
    return dict(value=M2 / sec2, unit=UNIT, cores=cores, kind=kind, host_cpu=model, host_logical_cpus=ncpu,
                sample="%d candidate edges = %s x 32 children on the %s map, %.1f s wall"
                       % (M2, what, wl["label"].split(":")[0], sec2)), M2, sec2


def same_population(wl, plan_mod, local):
    """Stages 2-4 (sample + integrate + collide + region index) on IDENTICAL inputs: the same parents x 32 children and
    the same Philox streams through (1) this library (kgmt_stage_propagate), (2) the reference's own propagateG kernel
    recompiled for sm_100a (baseline A of BASELINE.json; includes its map atomics), (3) the reference's host loop."""
    out = {"parents": None}
    try:
        import cudasbmp_b200 as k
        from oracle import pyoracle as po
        cfg, obs = wl["cfg"], wl["obstacles"]
        parents, what = sample_parents(wl)
        P = min(len(parents), 30000)
        parents = parents[:P]
        M = P * 32
        out["parents"] = "%d %s x 32 children, Philox key 99" % (P, what)
        p = k.KGMT(**dict(cfg, N=16, n=8, maxTreeSize=M), device=local, record_candidates=True)
        p.set_obstacles(obs)
        best = min(p.stage_propagate(parents, 32, 99, 0) for _ in range(5))
        out["ours"] = {"value": M / best * 1e3, "unit": UNIT, "ms": best, "what": "kgmt_stage_propagate (stages 2-4, candidate records written)"}
        valid_ours = float(p.export(plan_mod.ARR_U_VALID)[:M].mean())
        p.close()
        if po.ref_gpu() is not None:
            N, n = 16, 8
            c1, c2 = N * N, N * N * n * n
            maps = {kk: np.zeros(c1 if kk.startswith("R1") else c2, dtype=np.int32)
                    for kk in ("R1", "R2", "R1Valid", "R2Valid", "R1Invalid", "R2Invalid", "R1Avail", "R2Avail")}
            _, _, gnew, ms = po.ref_gpu_expand(1, 32, parents, np.arange(P, dtype=np.int32), maps, np.ones(c1, np.float32), N, n,
                                               cfg["width"] / N, cfg["width"] / (N * n), cfg["numDisc"], cfg["agentLength"], obs,
                                               cfg["width"], cfg["height"], 99, reps=5)
            out["reference_cuda"] = {"value": M / ms * 1e3, "unit": UNIT, "ms": ms,
                                     "what": "reference propagateG (KGMT.cu:341-414) unmodified, recompiled for sm_100a, cuRAND Philox"}
        R = po.ref_host()
        if R is not None:
            cores = R.ref_host_hw_threads()
            pof = (np.arange(M, dtype=np.int32) // 32)
            sec, _, _, _ = po.ref_host_batch(parents, pof, 99, 0, cfg["numDisc"], cfg["agentLength"], obs, cfg["width"], cfg["height"],
                                             threads=cores, outputs=False)
            out["reference_host"] = {"value": M / sec, "unit": UNIT, "cores": cores, "s": sec,
                                     "what": "reference statePropagator.cu + collisionCheck.cu unmodified, host build"}
        out["valid_fraction"] = valid_ours
    except Exception as e:
        out["error"] = repr(e)
    return out


def reference_plan_ms(runs=7):
    """The reference's OWN planner end to end on config 1 (its KGMT::plan, XORWOW, recompiled for sm_100a): wall clock
    around plan() as its 'time inside KGMT' (KGMT.cu:294-295), the constructor's allocations excluded."""
    try:
        from oracle import pyoracle as po
        from cudasbmp_b200 import workloads as w
        if po.ref_gpu() is None:
            return None
        ms, wall = [], []
        for i in range(runs + 2):
            r = po.ref_gpu_plan(w.C1, w.C1_INIT, w.C1_GOAL, w.C1_OBSTACLES)
            if i >= 2:
                ms.append(r["inside_ms"]); wall.append(r["plan_wall_ms"])
            time.sleep(1.01)                                     # the reference seeds cuRAND from time(NULL): one seed per second
        return {"median_ms": statistics.median(ms), "min_ms": min(ms), "runs": runs,
                "plan_wall_incl_csv_median_ms": statistics.median(wall),
                "what": "reference KGMT::plan() on config 1, unmodified sources recompiled for sm_100a: its own 'time inside KGMT' "
                        "(std::clock around its loop, KGMT.cu:82,294-295), seeds = time(NULL)"}
    except Exception as e:
        return {"error": repr(e)}


def ttfs_single(k, w, local):
    """Median time-to-first-solution, 101 seeds, host wall clock from the kgmt_plan call (state allocated, obstacles
    resident) to the host holding costToGoal != 0."""
    out = {}
    for name, cfg, obs, init, goal in (("c1", w.C1, w.C1_OBSTACLES, w.C1_INIT, w.C1_GOAL),
                                       ("c2", w.C2, w.c2_obstacles(1000), w.C2_INIT, w.C2_GOAL)):
        p = k.KGMT(**cfg, seed=1, device=local)
        p.set_obstacles(obs)
        for s in range(3):
            p.set_seed(1000 + s); p.plan(init, goal)
        walls, devs = [], []
        for s in range(1, 102):
            p.set_seed(s)
            tq = time.perf_counter()
            r = p.plan(init, goal)
            dtq = time.perf_counter() - tq
            if r["stop"] == 1:
                walls.append(dtq * 1e3); devs.append(r["device_ms"])
        if walls:
            ws = sorted(walls)
            out[name] = {"seeds": 101, "solved": len(walls), "median_ms": statistics.median(walls),
                         "p95_ms": ws[int(0.95 * (len(ws) - 1))], "device_median_ms": statistics.median(devs)}
        p.close()
    out["clock"] = "host wall around kgmt_plan"
    return out


def comm_id(k, rank, dist):
    """The NCCL unique id of the library's communicator: made by rank 0, handed to the others (here through
    torch.distributed, which the driver's launcher has set up; a C++ host uses a pipe or a file: demos/kgmt_multi_demo.cu)."""
    box = [k.KGMT.comm_unique_id() if rank == 0 else None]
    if dist is not None:
        dist.broadcast_object_list(box, src=0)
    return box[0]


def ttfs_multi(k, w, local, rank, world, dist, torch, races=101):
    """Portfolio over the N GPUs through the C ABI (kgmt_plan_portfolio): every rank plans the same query with its own
    seed in ONE launch; the first rank to reach the goal stops the others through a word in their memory; the best
    solution's result block and path are broadcast to every rank (NCCL).  Time-to-first-solution = host wall clock from a
    common barrier until the LAST rank holds the winner's solution."""
    out = {}
    race_id = 0
    for name, cfg, obs, init, goal in (("c1", w.C1, w.C1_OBSTACLES, w.C1_INIT, w.C1_GOAL),
                                       ("c2", w.C2, w.c2_obstacles(1000), w.C2_INIT, w.C2_GOAL)):
        p = k.KGMT(**cfg, seed=1, device=local)
        p.set_obstacles(obs)
        p.comm_init(rank, world, comm_id(k, rank, dist))
        rows = []
        for q in range(races + 3):
            race_id += 1
            p.comm_barrier()
            t0 = time.perf_counter()
            win, r, path = p.plan_portfolio(init, goal, 1000 * q + 1, race_id)
            dt = (time.perf_counter() - t0) * 1e3
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if q >= 3:
                rows.append((float(t[0]), win, r["device_ms"], len(path)))
        solved = [x for x in rows if x[1] >= 0]
        if solved:
            tt = sorted(x[0] for x in solved)
            out[name] = {"races": races, "solved": len(solved), "median_ms": statistics.median(tt), "p95_ms": tt[int(0.95 * (len(tt) - 1))],
                         "winner_device_ms_median": statistics.median(x[2] for x in solved),
                         "path_nodes_median": statistics.median(x[3] for x in solved)}
        # the race alone (kgmt_peer_race, no result exchange): when does the FIRST rank hold a solution?  All ranks are
        # processes of one host, so time.perf_counter (CLOCK_MONOTONIC) is one clock: first = min over the solving ranks
        # of their return time - the latest barrier exit.
        firsts = []
        for q in range(races + 3):
            race_id += 1
            p.set_seed(1000 * q + 1 + rank)
            p.comm_barrier()
            t0 = time.perf_counter()
            r = p.peer_race(init, goal, race_id)
            t1 = time.perf_counter()
            row = torch.tensor([t0, t1 if r["stop"] == 1 else float("inf")], dtype=torch.float64, device="cuda")
            allr = [torch.zeros_like(row) for _ in range(world)]
            dist.all_gather(allr, row)
            start = max(float(a[0]) for a in allr)
            done = min(float(a[1]) for a in allr)
            if q >= 3 and done != float("inf"):
                firsts.append((done - start) * 1e3)
        if firsts and name in out:
            fs = sorted(firsts)
            out[name]["first_solution_median_ms"] = statistics.median(fs)
            out[name]["first_solution_p95_ms"] = fs[int(0.95 * (len(fs) - 1))]
        p.comm_destroy(); p.close()
    out["clock"] = ("median_ms: host wall from a common barrier until the slowest rank holds the winning solution and path "
                    "(kgmt_plan_portfolio: peer-memory race + NCCL min-reduce + broadcast); first_solution_*: host wall from "
                    "the latest barrier exit until the first rank's kgmt_peer_race returns solved (one CLOCK_MONOTONIC for "
                    "all ranks of the box); seed = 1000*race + 1 + rank")
    return out


def bench_c3(k, plan_mod, wl3, local):
    """Config 3 on one GPU: culled back end (median of 5 plans) and the exhaustive TMA-streamed back end (one plan)."""
    out = {"workload": wl3["label"]}
    for name, mode, reps in (("grid", plan_mod.COLLIDE_GRID, 5), ("streamed_exhaustive", plan_mod.COLLIDE_BRUTE, 1)):
        p = k.KGMT(**wl3["cfg"], seed=1, device=local, collision_mode=mode)
        p.set_obstacles(wl3["obstacles"])
        if name == "grid":
            p.plan(wl3["init"], wl3["goal"])
        ms, exp = [], 0
        for s in range(reps):
            p.set_seed(1 + s)
            r = p.plan(wl3["init"], wl3["goal"])
            ms.append(r["device_ms"]); exp += r["expansions"]
        out[name] = {"plans": reps, "median_ms": statistics.median(ms), "expansions_per_s": exp / (sum(ms) * 1e-3),
                     "expansions_per_plan": exp / reps, "backend": p.config()["collide_backend"], "tree_size": r["tree_size"], "stop": r["stop"]}
        p.close()
    return out


def bench_c4(k, w, local, rank, world, dist, torch, Q=1024, reps=5):
    """Config 4: ONE fixed batch of Q queries on the reference demo map, sharded contiguously over the ranks
    (kgmt_plan_batch_sharded: every rank plans its shard in one launch, a thread-block cluster per query; the fixed-size
    result rows are all-gathered by NCCL inside the library).  Strong scaling: the same Q at every N.  Reported:
    queries/s and aggregate expansions/s on the slowest rank's device time and on wall time (barrier to every rank holding
    all Q results), median / p95 time-to-solution of a query (device clock from the start of its rank's launch to the
    query's last iteration)."""
    inits, goals = w.random_queries(Q, w.C1_OBSTACLES)
    seeds = np.arange(Q, dtype=np.uint32)
    p = k.KGMT(**w.C1, seed=1, device=local)
    p.set_obstacles(w.C1_OBSTACLES)
    p.comm_init(rank, world, comm_id(k, rank, dist))
    per = (Q + world - 1) // world
    cs = p.batch_cluster_size(per)
    p.plan_batch_sharded(inits, goals, seeds, cluster_size=cs)                          # warm-up: allocations, code upload
    best = None
    for _ in range(reps):
        p.comm_barrier()
        t0 = time.perf_counter()
        res, dev_ms = p.plan_batch_sharded(inits, goals, seeds, cluster_size=cs, as_array=True)   # kgmt_result rows, as the C ABI fills them
        wall = time.perf_counter() - t0
        tw = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        wall = float(tw[0])
        if best is None or wall < best["wall_ms"] * 1e-3:
            ok = res["stop"] == 1
            tt = np.sort(res["done_ms"][ok])
            exp = float(res["expansions"].sum())
            best = {"queries": Q, "gpus": world, "cluster_size": cs, "solved": int(ok.sum()),
                    "device_ms_max": dev_ms, "wall_ms": wall * 1e3, "queries_per_s_device": Q / dev_ms * 1e3,
                    "queries_per_s_wall": Q / wall, "expansions": exp,
                    "expansions_per_s_device": exp / dev_ms * 1e3, "expansions_per_s_wall": exp / wall,
                    "time_to_solution_median_ms": float(np.median(tt)) if len(tt) else None,
                    "time_to_solution_p95_ms": float(tt[int(0.95 * (len(tt) - 1))]) if len(tt) else None,
                    "service_ms_median": float(np.median(res["service_ms"][ok])) if ok.any() else None}
    p.comm_destroy(); p.close()
    return best


def bench_c5(k, K, w, local, rank, world, dist, torch, logs=(20, 22, 24, 26), reps=3, P=32768):
    """Config 5: ONE iteration of M = 2^log candidates (P parents in free space of the config-2 map, each expanded M/P
    times), sharded over the ranks through kgmt_expand_sharded with its three exchanges: fused into the persistent kernel
    (peer memory), the multi-launch peer sequence, NCCL all-gather / all-reduce (compute and exchange time separately).
    At N = 1 also the single-GPU cooperative kernel.  CUDA-event times, max over ranks, best of `reps`.
    Plus whole config-2 PLANS with sharded iterations (kgmt_plan_sharded), median of 5."""
    obs = w.c2_obstacles(1000)
    parents = w.random_parents(P, obs, seed=7)
    Mmax = 1 << max(logs)
    p = k.KGMT(**dict(w.C1, maxTreeSize=Mmax + P, numIterations=4), seed=5, device=local, max_candidates=Mmax)
    p.set_obstacles(obs)
    p.comm_init(rank, world, comm_id(k, rank, dist))
    out = {"parents": P, "K": 1000, "gpus": world, "points": []}

    def mx(*v):
        t = torch.tensor(list(v), dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    names = {K.EXCHANGE_FUSED: "fused", K.EXCHANGE_PEER_LAUNCHES: "peer_launches", K.EXCHANGE_NCCL: "nccl"}
    for lg in logs:
        M = 1 << lg
        pt = {"log2M": lg, "M": M}
        for ex, nm in names.items():
            best = None
            for rep in range(reps + 1):
                p.set_seed(5 + rep)
                p.seed_frontier(parents, w.C2_GOAL)
                p.set_children(M // P)
                p.comm_barrier()
                st = p.expand_sharded(ex, timing=True)
                comp, exch = mx(st["compute_ms"], st["exchange_ms"])
                if rep > 0 and (best is None or comp + exch < best[0] + best[1]):
                    best = (comp, exch, st["exchange_bytes"], st["accepted"])
            pt["accepted"] = best[3]
            if ex == K.EXCHANGE_NCCL:
                pt["nccl_compute_ms"], pt["nccl_exchange_ms"], pt["nccl_exchange_bytes"] = best[0], best[1], best[2]
                pt["nccl_expansions_per_s"] = M / (best[0] + best[1]) * 1e3
            else:
                pt[nm + "_ms"] = best[0]
                pt[nm + "_expansions_per_s"] = M / best[0] * 1e3
        if world == 1:
            ms = []
            for rep in range(reps):
                p.set_seed(6 + rep)
                p.seed_frontier(parents, w.C2_GOAL); p.set_children(M // P)
                s = torch.cuda.ExternalStream(p.stream)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s); p.iterate(); e1.record(s); torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            pt["cooperative_kernel_ms"] = min(ms)
            pt["cooperative_expansions_per_s"] = M / min(ms) * 1e3
        out["points"].append(pt)
    p.comm_destroy(); p.close()
    # whole config-2 plans, every iteration's candidates split over the ranks, one persistent kernel per rank
    q = k.KGMT(**w.C2, seed=1, device=local)
    q.set_obstacles(obs)
    q.comm_init(rank, world, comm_id(k, rank, dist))
    q.plan_sharded(w.C2_INIT, w.C2_GOAL)
    ms, exp, tree = [], 0, 0
    for s in range(5):
        q.set_seed(1 + s)
        q.comm_barrier()
        r = q.plan_sharded(w.C2_INIT, w.C2_GOAL)
        ms.append(mx(r["device_ms"])[0]); exp += r["expansions"]; tree = r["tree_size"]
    out["c2_plan_sharded"] = {"plans": 5, "median_ms": statistics.median(ms), "expansions_per_s": exp / (sum(ms) * 1e-3),
                              "tree_size_last": tree, "what": "kgmt_plan_sharded: whole config-2 plans, tree bit-identical to kgmt_plan"}
    q.comm_destroy(); q.close()
    return out


def fp32_roofline(k, plan_mod, wl, local, rates, value, clock_mhz, sms):
    """roofline_fp32 of SURVEY.md §8d: exp/s x F_exp(K_tested, S_mean) against the MEASURED lane rate of the instruction
    class that dominates (compare-class: 4 FSETP per overlap test), with K_tested = K (what the reference executes) and
    with the pairs the culled kernel really tests (device work counters of a recording plan of the same seed)."""
    try:
        p = k.KGMT(**wl["cfg"], seed=21, device=local, record_candidates=True)
        p.set_obstacles(wl["obstacles"])
        p.plan(wl["init"], wl["goal"])
        wc = p.work_counters()
        p.close()
        S = wc["steps"] / max(wc["expansions"], 1)
        pairs = wc["pairs"] / max(wc["steps"], 1)
        Kobs = len(wl["obstacles"])
        F_fixed, F_step = 110.0, 60.0                                  # F_rng + F_ctrl + tanf ; F_trig + F_dyn per step
        f_brute = F_fixed + S * (F_step + 4.0 * Kobs)
        f_culled = F_fixed + S * (F_step + 4.0 * pairs)
        fsetp = (rates or {}).get("fsetp", {}).get("lane_ops_per_clk_per_sm")
        ffma = (rates or {}).get("ffma", {}).get("lane_ops_per_clk_per_sm")
        src = "profiles/issue_rates_b200.json (scripts/native/issue_rate_bench.cu, measured on the box)"
        if not fsetp:
            fsetp, src = 64.0, "ASSUMED 64 compare lanes/clk/SM (no measured profiles/issue_rates_b200.json)"
        peak_cmp = fsetp * sms * clock_mhz * 1e6
        out = {"steps_per_expansion": S, "pairs_tested_per_step": pairs, "K": Kobs,
               "lane_ops_per_expansion_brute_equivalent": f_brute, "lane_ops_per_expansion_culled": f_culled,
               "compare_lane_ops_per_clk_per_sm": fsetp, "ffma_lane_ops_per_clk_per_sm": ffma, "peak_compare_lane_ops_per_s": peak_cmp,
               "frac_brute_equivalent": value * f_brute / peak_cmp, "frac_culled": value * f_culled / peak_cmp,
               "peak_source": src, "sm_mhz": clock_mhz,
               "note": "brute-equivalent > 1 means the culled kernel does less arithmetic than the reference's algorithm needs at this roof"}
        return out
    except Exception as e:
        return {"error": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collide", default="grid", choices=["grid", "brute"])
    ap.add_argument("--plans-per-step", type=int, default=PLANS_PER_STEP)
    ap.add_argument("--skip", default="", help="comma list of sections to skip: ttfs,c3,c4,c5,same,fp32,e2e")
    ap.add_argument("--only-headline", action="store_true", help="headline + e2e only (profiling runs)")
    args = ap.parse_args()
    skip = set(s for s in args.skip.split(",") if s)
    if args.only_headline:
        skip |= {"ttfs", "c3", "c4", "c5", "same", "fp32"}
        args.no_cpu_baseline = True
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = workload(args.workload)
    pps = max(1, args.plans_per_step)
    config = {"workload": wl["label"],
              "step": "one batch of %d complete plans (all expansion iterations each), every plan its own seed" % pps,
              "plans_per_step": pps,
              "sharding": "independent queries/seeds per GPU, no data-path collective", "collision": args.collide,
              "l2": "flushed before every plan (256 MiB device write, outside the CUDA-event region)"}

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
    from cudasbmp_b200 import workloads as w
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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def seed_of(step, j):
        return 1 + ((step * pps + j) * world + rank)          # independent query per plan and rank

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run_resident(step):
        ms, res = 0.0, []
        for j in range(pps):
            flush.fill_((step + j) & 0xFF)
            torch.cuda.synchronize()
            plan.set_seed(seed_of(step, j))
            r = plan.plan(wl["init"], wl["goal"])
            ms += r["device_ms"]
            res.append(r)
        return ms, res

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
        ms, res = run_resident(args.warmup + s)
        dev_ms += ms
        results += res
    barrier()
    wall = time.perf_counter() - t0
    launches = plan.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    expansions = sum(r["expansions"] for r in results)
    accepted = sum(r["tree_size"] - 1 for r in results)

    # ---- end to end through the C ABI with host buffers
    obs_host = np.ascontiguousarray(wl["obstacles"], dtype=np.float32)
    e2e_exp, d2h, e2e_s = 0, 0, 0.0
    if "e2e" not in skip:
        for s in range(3):
            plan.set_seed(seed_of(0, s)); plan.set_obstacles(obs_host); plan.plan(wl["init"], wl["goal"])
        barrier()
        t1 = time.perf_counter()
        for s in range(args.steps):
            for j in range(pps):
                plan.set_obstacles(obs_host)
                plan.set_seed(seed_of(args.warmup + s, j))
                r = plan.plan(wl["init"], wl["goal"])
                e2e_exp += r["expansions"]
                d2h += 4 + STATE_BYTES                        # cull-grid item count + the planner's scalar block
                if r["stop"] == 1:
                    path = plan.extract_path()
                    d2h += 16 + len(path) * 28                    # path header + its AoS-7 rows (one copy)
        barrier()
        e2e_s = time.perf_counter() - t1
    cfgd = plan.config()
    h2d = pps * (obs_host.nbytes + 56)        # the obstacle set + init/goal rows; the cull grid is built on the device

    # ---- the other named configurations (extra keys of the line)
    extra = {}
    t_extra = time.perf_counter()
    if "ttfs" not in skip:
        if world == 1:
            extra["ttfs"] = ttfs_single(k, w, local)
            extra["ttfs"]["reference_ms"] = reference_plan_ms()
        else:
            try:
                extra["ttfs_multi"] = ttfs_multi(k, w, local, rank, world, dist, torch)
            except Exception as e:
                extra["ttfs_multi"] = {"error": repr(e)}
    if "c3" not in skip and world == 1:
        try:
            extra["c3"] = bench_c3(k, K, workload("c3"), local)
        except Exception as e:
            extra["c3"] = {"error": repr(e)}
    if "c4" not in skip:
        try:
            # the named batch (1 024 queries) and one that fills an 8-GPU box (8 192): strong scaling, the same Q at every N
            for key, Q, reps in (("c4", 1024, 5), ("c4_8192", 8192, 3)):
                extra[key] = bench_c4(k, w, local, rank, world, dist, torch, Q=Q, reps=reps)
                if world > 1:
                    # the same fixed batch on ONE GPU of this box (rank 0 alone), so the line carries its own strong-scaling base
                    one = bench_c4(k, w, local, 0, 1, None, torch, Q=Q, reps=reps) if rank == 0 else None
                    torch.cuda.synchronize()
                    dist.barrier()
                    extra[key]["one_gpu_same_box"] = one
        except Exception as e:
            extra["c4"] = {"error": repr(e)}
    if "c5" not in skip:
        try:
            extra["c5"] = bench_c5(k, K, w, local, rank, world, dist, torch)
        except Exception as e:
            extra["c5"] = {"error": repr(e)}
    extra_s = time.perf_counter() - t_extra

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
    peaks, rates = {}, {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    try:
        rates = json.load(open(os.path.join(ROOT, "profiles", "issue_rates_b200.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    n_plans = args.steps * pps
    per_launch_exp = exp_all / world / n_plans                 # dominant kernel: ONE cooperative launch per plan
    kern_ms = dev_ms / n_plans
    achieved = per_launch_exp * (B_EXP + B_INS * alpha) / (kern_ms * 1e-3) / 1e9
    solved = [r for r in results if r["stop"] == 1]
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_inputs.json")))
        if prof.get("workload") != args.workload or prof.get("collision") != args.collide:
            prof = {}
    except Exception:
        prof = {}
    sm_mhz = float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
    sm_clock_hz = 1e6 * sm_mhz
    issue_peak = cfgd["sms"] * 4 * sm_clock_hz                 # one warp instruction per scheduler per clock
    ipe = prof.get("warp_instructions_per_expansion")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "e2e": {"value": (e2e_all / e2e_s_max) if e2e_s_max > 0 else None, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h / max(args.steps, 1)), "ms_per_step": 1e3 * e2e_s_max / args.steps},
        "gpu_launches": int(launches_all),
        "clocks": clocks,
        "timed_region_s": dev_ms_max * 1e-3, "timed_wall_s": wall_max,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": prof.get("dram_bytes_per_launch"), "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                     "kernel": "kgmt::expand_kernel<grid|brute, RECORD, chunks in flight> (cooperative, one launch per plan)",
                     "algorithmic_bytes_per_expansion": B_EXP + B_INS * alpha, "accept_ratio": alpha,
                     "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, %s)" % prof.get("source"),
                     "algorithmic_bytes_per_launch": per_launch_exp * (B_EXP + B_INS * alpha),
                     "launch_ms": kern_ms,
                     "note": "stages 2-4 are instruction-issue bound, not HBM bound (DESIGN.md); see roofline_issue, roofline_fp32 and profiles/"},
        "roofline_issue": {"bound": "warp-instruction issue", "unit": "G warp-inst/s",
                           "achieved": (per_launch_exp * ipe / (kern_ms * 1e-3) / 1e9) if ipe else None,
                           "peak": issue_peak / 1e9, "frac": (per_launch_exp * ipe / (kern_ms * 1e-3) / issue_peak) if ipe else None,
                           "warp_instructions_per_expansion": ipe, "avg_active_lanes": prof.get("avg_active_lanes"),
                           "peak_source": "SMs x 4 schedulers x SM clock under load (nvidia-smi during the timed region); "
                                          "1 warp-inst/clk/scheduler confirmed by profiles/issue_rates_b200.json" if rates else
                                          "SMs x 4 schedulers x SM clock under load (nvidia-smi during the timed region)",
                           "source": prof.get("source")},
        "plan": {"expansions_per_plan": per_launch_exp, "tree_size_mean": acc_all / world / n_plans + 1,
                 "iterations_mean": statistics.mean(r["iterations"] for r in results),
                 "plans_timed": n_plans, "solved": len(solved), "stops": sorted(set(r["stop"] for r in results)),
                 "device_ms_per_plan": kern_ms,
                 "time_to_first_solution_ms_median": statistics.median(r["device_ms"] for r in solved) if solved else None,
                 "host_wall_ms_per_plan": 1e3 * wall_max / n_plans},
        "collide_backend": cfgd,
        "extra_sections_s": extra_s,
    }
    line.update(extra)
    if "fp32" not in skip:
        line["roofline_fp32"] = fp32_roofline(k, K, wl, local, rates, value / world, sm_mhz, cfgd["sms"])
    if "same" not in skip and world == 1:
        line["same_population"] = same_population(wl, K, local)
        sp = line["same_population"]
        if "reference_cuda" in sp:
            line["ref_cuda_baseline"] = dict(sp["reference_cuda"], kind="reference CUDA kernel propagateG recompiled for sm_100a", sample=sp["parents"])
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
