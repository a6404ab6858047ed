"""config 4: Q independent car queries on the C1 map, batched in one launch (kgmt_plan_batch) vs one kgmt_plan per query."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cudasbmp_b200 import kgmt as K, workloads as w
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
inits, goals = w.random_queries(Q, w.C1_OBSTACLES)
seeds = np.arange(Q)
p = K.KGMT(**w.C1, seed=1); p.set_obstacles(w.C1_OBSTACLES)
t0 = time.perf_counter(); exp = 0
for q in range(min(Q, 256)):
    p.set_seed(int(seeds[q])); exp += p.plan(inits[q], goals[q])["expansions"]
dt = time.perf_counter() - t0
print(json.dumps(dict(mode="sequential kgmt_plan", queries=min(Q, 256), wall_s=dt, queries_per_s=min(Q, 256) / dt, expansions_per_s=exp / dt)))
for cs in (1, 2, 4, 8):
    p.plan_batch(inits[:64], goals[:64], seeds[:64], cluster_size=cs)
    t0 = time.perf_counter()
    res, ms, _, ws = p.plan_batch(inits, goals, seeds, cluster_size=cs, max_path=64)
    dt = time.perf_counter() - t0
    exp = sum(r["expansions"] for r in res)
    print(json.dumps(dict(mode="batch", cluster=cs, workspaces=ws, queries=Q, device_ms=ms, wall_s=dt, queries_per_s_device=Q / ms * 1e3,
                          expansions_per_s_device=exp / ms * 1e3, solved=sum(r["stop"] == 1 for r in res))))
