"""Debug build only (make EXTRA=-DKGMT_PIPE_PROF): per-warp event trace of one iteration of the pipelined loop."""
import sys, os
os.environ["KGMT_ITERLOG_ROWS"] = str(256 + 8192 * 4)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cudasbmp_b200 import kgmt as K, workloads as w
p = K.KGMT(**w.C2, seed=1, loop=1); p.set_obstacles(w.c2_obstacles(1000)); p.iteration_log(True)
p.set_seed(1); r = p.plan(w.C2_INIT, w.C2_GOAL)
log = p.iteration_log()
print(r, p.config())
tot = log[255].astype(np.float64)
names = ["parent wait", "P (setup+propagate)", "U (finish, incl. score wait)", "sign-off", "insertion", "spec wait", "epoch wait", "iteration start (ticket)"]
for n, v in zip(names, tot): print("%-32s %6.1f %%" % (n, 100 * v / tot.sum()))
nw = p.config()["grid"] * 8
tr = log[256:].reshape(-1)[: nw * 32].reshape(nw, 32).astype(np.int64)
tr = tr[tr[:, 0] > 0]
t0 = tr[:, 0].min()
def us(col): 
    v = tr[:, col]; v = v[v > 0]; return (v - t0) / 1e3
def show(name, v):
    if len(v): print("%-26s n %5d  min %7.1f p10 %7.1f med %7.1f p90 %7.1f p99 %7.1f max %7.1f" % (name, len(v), v.min(), np.percentile(v, 10), np.median(v), np.percentile(v, 90), np.percentile(v, 99), v.max()))
show("iteration seen", us(0)); show("chunk loop left", us(1)); show("signed off", us(2)); show("insertion left", us(3))
show("spec wait over", us(4)); show("spec P done", us(5)); show("next epoch seen", us(6))
# per-chunk durations
for k in range(8):
    a, b, c = tr[:, 8 + 3 * k], tr[:, 9 + 3 * k], tr[:, 10 + 3 * k]
    m = (a > 0) & (c > 0)
    if m.sum() == 0: break
    pw = np.where(b[m] > 0, (b[m] - a[m]) / 1e3, 0.0); fin = np.where(b[m] > 0, (c[m] - b[m]) / 1e3, (c[m] - a[m]) / 1e3)
    print("chunk %d: n %5d  start med %6.1f max %6.1f | wait+P med %5.1f p99 %5.1f max %5.1f | U med %5.1f p99 %5.1f max %5.1f | end med %6.1f max %6.1f" % (
        k, m.sum(), np.median((a[m] - t0) / 1e3), ((a[m] - t0) / 1e3).max(), np.median(pw), np.percentile(pw, 99), pw.max(), np.median(fin), np.percentile(fin, 99), fin.max(),
        np.median((c[m] - t0) / 1e3), ((c[m] - t0) / 1e3).max()))
