"""Config-2 plan time against cull-grid resolution and resident CTAs per SM (device ms, median of 9 seeds)."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cudasbmp_b200 import kgmt as K, workloads as w

obs = w.c2_obstacles(1000)
culls = [int(a) for a in sys.argv[1].split(',')] if len(sys.argv) > 1 else (32, 40, 48, 56, 64, 72)
pers = [int(a) for a in sys.argv[2].split(',')] if len(sys.argv) > 2 else (0, 2, 4)
for cull in culls:
    for per in pers:
        try:
            p = K.KGMT(**w.C2, seed=1, cull_cells=cull, stage_limit_bytes=110 * 1024, ctas_per_sm=per)
            p.set_obstacles(obs)
            ms = []
            for s in range(10):
                p.set_seed(1 + s); r = p.plan(w.C2_INIT, w.C2_GOAL); ms.append(r["device_ms"])
            print("cull %3d ctas/sm %d -> %s  exp %d  median ms %.4f" % (cull, per, p.config(), r["expansions"], statistics.median(ms[1:])), flush=True)
            del p
        except Exception as e:
            print("cull", cull, per, "failed", e, flush=True)
