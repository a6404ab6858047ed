"""Per-iteration device timeline of one plan + a sweep over cull-grid resolution / staging."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cudasbmp_b200 import kgmt as K, workloads as w

def run(cfg, obs, init, goal, reps=5, **kw):
    p = K.KGMT(**cfg, seed=1, **kw); p.set_obstacles(obs)
    p.iteration_log(True)
    ms = []
    for s in range(reps):
        p.set_seed(1 + s); r = p.plan(init, goal); ms.append(r["device_ms"])
    return p, r, ms

if "c1" in sys.argv:
    cfg, obs, INIT, GOAL = w.C1, w.C1_OBSTACLES, w.C1_INIT, w.C1_GOAL
else:
    cfg, obs, INIT, GOAL = w.C2, w.c2_obstacles(1000), w.C2_INIT, w.C2_GOAL
p, r, ms = run(cfg, obs, INIT, GOAL)
log = p.iteration_log()
print("plan", r, "ms", ms)
t = log[:, 0].astype(np.int64); dt = np.diff(t, prepend=t[0])
for i, row in enumerate(log):
    r = row.astype(np.int64)
    ph = [(r[k] - r[2]) / 1e3 if r[k] else float("nan") for k in (3, 4, 5, 0, 6)]
    print("itr %2d  M %8d acc %7d  dt %7.1f us | CTA0: A done %6.1f  bar1 %6.1f  B done %6.1f  finalized %6.1f  bar2 %6.1f" % (
        i + 1, r[1] >> 32, r[1] & 0xFFFFFFFF, dt[i] / 1e3, *ph))
if "timeline" in sys.argv:
    sys.exit(0)
print("---- sweep")
for cull in (16, 32, 48, 64, 96, 128):
    for lim in (48 * 1024, 200 * 1024, 1):
        try:
            p, r, ms = run(cfg, obs, w.C2_INIT, w.C2_GOAL, cull_cells=cull, stage_limit_bytes=lim)
            print("cull %3d lim %6d -> %s  exp %d  ms %s" % (cull, lim, p.config(), r["expansions"], ["%.3f" % m for m in ms]))
        except Exception as e:
            print("cull", cull, lim, "failed", e)
p, r, ms = run(cfg, obs, w.C2_INIT, w.C2_GOAL, collision_mode=K.COLLIDE_BRUTE)
print("brute", p.config(), r["expansions"], ms)
