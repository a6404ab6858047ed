"""A/B of the two planner loops (grid barrier vs barrier-free pipelined): identical trees, timing."""
import sys, os, zlib, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cudasbmp_b200 import kgmt as K, workloads as w

def crc(p, T):
    h = zlib.crc32(p.export(K.ARR_SAMPLES)[:T].tobytes())
    h = zlib.crc32(p.export(K.ARR_PARENT)[:T].tobytes(), h)
    h = zlib.crc32(p.export(K.ARR_COSTS)[:T].tobytes(), h)
    for m in (K.ARR_R1, K.ARR_R1VALID, K.ARR_R1INVALID, K.ARR_R1AVAIL, K.ARR_R2, K.ARR_R2VALID, K.ARR_R2INVALID, K.ARR_R2AVAIL, K.ARR_R1SCORE):
        h = zlib.crc32(p.export(m).tobytes(), h)
    return h

cases = [("c1", w.C1, w.C1_OBSTACLES, w.C1_INIT, w.C1_GOAL), ("c2", w.C2, w.c2_obstacles(1000), w.C2_INIT, w.C2_GOAL)]
which = sys.argv[1:] or ["c1", "c2"]
for name, cfg, obs, init, goal in cases:
    if name not in which: continue
    out = {}
    for loop in (2, 1):
        p = K.KGMT(**cfg, seed=1, loop=loop); p.set_obstacles(obs)
        rows = []
        for s in (1, 2, 3, 4, 5):
            p.set_seed(s); r = p.plan(init, goal)
            rows.append((r["stop"], r["iterations"], r["tree_size"], r["expansions"], r["goal_index"], r["cost_to_goal"], crc(p, r["tree_size"]), round(r["device_ms"], 4)))
        out[loop] = rows
        print(name, "loop", loop, [x[-1] for x in rows], rows[-1][:6], flush=True)
    same = all(a[:-1] == b[:-1] for a, b in zip(out[1], out[2]))
    print(name, "IDENTICAL" if same else "DIFFERENT", flush=True)
    if not same:
        for a, b in zip(out[1], out[2]): print("  pipe", a, "\n  barr", b)
