"""A/B of phase A with 1..4 chunks in flight per warp (lane refill) on configs 1 and 2: identical trees, device time per
plan (median of `reps` seeds), work counters.   python scripts/ab_groups.py [reps]"""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cudasbmp_b200 import kgmt as K, workloads as w

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 31
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for name, cfg, obs, init, goal in (("c2", w.C2, w.c2_obstacles(1000), w.C2_INIT, w.C2_GOAL),
                                   ("c1", w.C1, w.C1_OBSTACLES, w.C1_INIT, w.C1_GOAL)):
    base = None
    for g in (1, 2, 3, 4):
        for ctas in (0,):
            p = K.KGMT(**cfg, seed=1, chunks_in_flight=g, ctas_per_sm=ctas)
            p.set_obstacles(obs)
            for s in range(3):
                p.set_seed(900 + s); p.plan(init, goal)
            ms, exp, sig = [], 0, []
            for s in range(reps):
                flush.fill_(s & 0xFF); torch.cuda.synchronize()
                p.set_seed(1 + s)
                r = p.plan(init, goal)
                ms.append(r["device_ms"]); exp += r["expansions"]; sig.append((r["tree_size"], r["iterations"], r["stop"], r["expansions"]))
            if base is None:
                base = sig
            cfgd = p.config()
            row = dict(group=g, grid=cfgd["grid"], smem=cfgd["smem_bytes"], median_ms=statistics.median(ms), mean_ms=statistics.mean(ms),
                       gexp_per_s=exp / (sum(ms) * 1e-3) / 1e9, same_trees=(sig == base))
            out["%s_g%d" % (name, g)] = row
            print(name, row, flush=True)
            p.close()
    # work counters of one recorded plan (the recording kernels count)
    p = K.KGMT(**cfg, seed=1, record_candidates=True)
    p.set_obstacles(obs)
    r = p.plan(init, goal)
    wc = p.work_counters()
    wc.update(steps_per_expansion=wc["steps"] / max(wc["expansions"], 1), pairs_per_step=wc["pairs"] / max(wc["steps"], 1),
              pairs_per_expansion=wc["pairs"] / max(wc["expansions"], 1), K=len(obs), num_disc=cfg["numDisc"])
    out[name + "_work"] = wc
    print(name, "work", wc, flush=True)
    p.close()
print(json.dumps(out))
