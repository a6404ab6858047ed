"""Baseline A: the reference's own CUDA kernels (unmodified, sm_100a, Philox states) timed with CUDA events on the same
inputs as ours (stages 2-5a: propagateG incl. map atomics), plus the reference's full plan() (XORWOW, its own host loop)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cudasbmp_b200 import kgmt as K, workloads as w
from oracle import pyoracle as po
import ctypes as C

R = po.ref_gpu()
assert R is not None, "oracle/_ref/libref_gpu.so missing"
out = {}
for name, obs, nd in (("c1", w.C1_OBSTACLES, 10), ("c2", w.c2_obstacles(1000), 10)):
    for P in (1024, 8192, 30000):
        parents = w.random_parents(P, obs, seed=7)
        N, n = 16, 8
        c1, c2 = N * N, N * N * n * n
        maps = {k: np.zeros(c1 if k.startswith("R1") else c2, dtype=np.int32)
                for k in ("R1", "R2", "R1Valid", "R2Valid", "R1Invalid", "R2Invalid", "R1Avail", "R2Avail")}
        _, _, _, ms = po.ref_gpu_expand(1, 32, parents, np.arange(P, dtype=np.int32), maps, np.ones(c1, np.float32), N, n,
                                        1.25, 0.15625, nd, 1.0, obs, 20.0, 20.0, 99, reps=5)
        M = P * 32
        plan = K.KGMT(**dict(w.C1, maxTreeSize=M), record_candidates=False); plan.set_obstacles(obs)
        plan.stage_propagate(parents, 32, 99, 0)
        ours = min(plan.stage_propagate(parents, 32, 99, 0) for _ in range(5))
        out["%s_P%d" % (name, P)] = dict(M=M, ref_ms=ms, ref_exp_per_s=M / ms * 1e3, ours_stage24_ms=ours, ours_exp_per_s=M / ours * 1e3)
        print(name, "P", P, "M", M, "reference propagateG %.3f ms = %.3g exp/s | ours (stages 2-4 + records) %.3f ms = %.3g exp/s" % (
            ms, M / ms * 1e3, ours, M / ours * 1e3), flush=True)
# the reference's whole plan() on its demo configuration (C1), wall clock, 11 runs (seed = time(NULL))
f32p = po.f32p
ts, sizes = [], []
for i in range(11):
    tsz = C.c_int(); cost = C.c_float()
    obs = np.ascontiguousarray(w.C1_OBSTACLES)
    t0 = time.perf_counter()
    rc = R.ref_gpu_plan(20.0, 20.0, 16, 8, 100, 30000, 10, 1.0, 0.5, w.C1_INIT.ctypes.data_as(f32p), w.C1_GOAL.ctypes.data_as(f32p),
                        obs.ctypes.data_as(f32p), 5, C.byref(tsz), C.byref(cost), None)
    ts.append(time.perf_counter() - t0); sizes.append((tsz.value, cost.value))
    time.sleep(1.01)   # its seed is time(NULL)
print("reference plan() C1 wall ms (incl. its CSV dump + allocs):", ["%.1f" % (t * 1e3) for t in ts], sizes)
p = K.KGMT(**w.C1, seed=1); p.set_obstacles(w.C1_OBSTACLES)
ours = []
for s in range(11):
    p.set_seed(s + 1); t0 = time.perf_counter(); r = p.plan(w.C1_INIT, w.C1_GOAL); ours.append((time.perf_counter() - t0, r["device_ms"], r["tree_size"], r["stop"]))
print("ours plan() C1 (wall ms, device ms, tree, stop):", [("%.3f" % (a * 1e3), "%.3f" % b, c, d) for a, b, c, d in ours])
out["ref_plan_c1_wall_ms"] = [t * 1e3 for t in ts]
out["ours_plan_c1"] = ours
json.dump(out, open("gpurun_out/baseline_ref_gpu.json", "w"), indent=1)
