"""Finalizer timeline of the pipelined loop: per iteration, microseconds since the previous publication."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cudasbmp_b200 import kgmt as K, workloads as w
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg, obs, init, goal = (w.C1, w.C1_OBSTACLES, w.C1_INIT, w.C1_GOAL) if name == "c1" else (w.C2, w.c2_obstacles(1000), w.C2_INIT, w.C2_GOAL)
p = K.KGMT(**cfg, seed=1, loop=1); p.set_obstacles(obs); p.iteration_log(True)
for s in (1, 2, 3):
    p.set_seed(s); r = p.plan(init, goal)
print(r)
log = p.iteration_log().astype(np.int64)
prev = None
for i, row in enumerate(log):
    pub = row[5]
    base = prev if prev is not None else row[2]
    cols = [(row[k] - base) / 1e3 for k in (2, 3, 4, 5)]
    print("itr %2d M %8d acc %7d | all CTAs signed off %7.1f chain %7.1f prev-insert+recycle %7.1f published %7.1f us" % (
        i + 1, row[1] >> 32, row[1] & 0xFFFFFFFF, *cols))
    prev = pub
