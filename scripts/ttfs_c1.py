"""Median time-to-first-solution on BASELINE config 1 (reference demo) over 101 seeds, device and host wall time."""
import sys, os, time, statistics, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cudasbmp_b200 import kgmt as K, workloads as w
kw = {}
for a in sys.argv[1:]:
    k, v = a.split("="); kw[k] = int(v)
p = K.KGMT(**w.C1, seed=1, **kw); p.set_obstacles(w.C1_OBSTACLES)
for s in range(5): p.set_seed(900 + s); p.plan(w.C1_INIT, w.C1_GOAL)
dev, wall, its = [], [], []
for s in range(1, 102):
    p.set_seed(s); t0 = time.perf_counter(); r = p.plan(w.C1_INIT, w.C1_GOAL); dt = time.perf_counter() - t0
    if r["stop"] == 1: dev.append(r["device_ms"]); wall.append(dt * 1e3); its.append(r["iterations"])
print(json.dumps(dict(args=kw, solved=len(dev), device_median_ms=statistics.median(dev), wall_median_ms=statistics.median(wall),
                      iterations_median=statistics.median(its), grid=p.config()["grid"])))
