import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cudasbmp_b200 import kgmt as K, workloads as w
from oracle import pyoracle as po
from tests import parity
po.build(ref=False)
plan = K.KGMT(**w.C1, seed=1, record_candidates=True); plan.set_obstacles(w.C1_OBSTACLES)
plan.begin(w.C1_INIT, w.C1_GOAL)
for it in range(8):
    res0 = plan.result(); T0 = res0["tree_size"]
    tree0 = plan.export(K.ARR_SAMPLES).copy()
    st = plan.iterate(); M = st["candidates"]
    cand = plan.export(K.ARR_UNEXPLORED)[:M]; valid = plan.export(K.ARR_U_VALID)[:M]; upar = plan.export(K.ARR_U_PARENT)[:M]
    xo, vo, u3o, margin = po.propagate_batch(tree0, upar, (1 + st["iteration"]) & 0xffffffff, 0, 10, 1.0, w.C1_OBSTACLES, 20., 20., po.MATH_FMA)
    err = np.abs(cand[:, :4].astype(np.float64) - xo[:, :4]) / np.maximum(1.0, np.abs(xo[:, :4]))
    off = ((valid != vo) | (err.max(axis=1) > parity.state_tolerance(xo, tree0[upar, 2], 10))) & (margin > 5e-4)
    print("itr", st["iteration"], "M", M, "off", int(off.sum()), "maxerr(all)", err.max(), "flag mismatches", int((valid != vo).sum()))
    for i in np.nonzero(off)[0][:5]:
        print("  cand", i, "parent", tree0[upar[i]], "\n   gpu", cand[i], valid[i], "\n   cpu", xo[i], vo[i], "margin", margin[i], "tan", np.tan(float(xo[i, 5])))
    if st["stop"]: break
