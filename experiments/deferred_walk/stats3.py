"""Mip-level bit queries: one bit lookup per step at the level where the step's cell rectangle spans <= 2 cells."""
import sys, numpy as np
sys.path.insert(0, '.')
from cudasbmp_b200 import workloads as w

def build(C, ob, W=20.0):
    inv = np.float32(C / W)
    cell = lambda v: np.clip(np.floor(v * inv).astype(np.int64), 0, C - 1)
    T = np.zeros((C, C), bool); I = np.zeros((C, C), bool)
    for o in ob:
        x0, y0, x1, y1 = cell(o[0]), cell(o[1]), cell(o[2]), cell(o[3])
        T[y0:y1 + 1, x0:x1 + 1] = True
        if x1 - x0 >= 2 and y1 - y0 >= 2: I[y0 + 1:y1, x0 + 1:x1] = True
    levels = [T]
    while levels[-1].shape[0] > 1:
        t = levels[-1]; levels.append(t[0::2, 0::2] | t[1::2, 0::2] | t[0::2, 1::2] | t[1::2, 1::2])
    return cell, levels, I

def shapes(t):
    Cl = t.shape[0]
    p = np.zeros((Cl + 1, Cl + 1), bool); p[:Cl, :Cl] = t
    q = {}
    q[(0, 0)] = t
    q[(1, 0)] = p[:Cl, :Cl] | p[:Cl, 1:]
    q[(0, 1)] = p[:Cl, :Cl] | p[1:, :Cl]
    q[(1, 1)] = p[:Cl, :Cl] | p[:Cl, 1:] | p[1:, :Cl] | p[1:, 1:]
    return q

def run(C, ob, P, four_shapes, numDisc=10, seed=1):
    rng = np.random.default_rng(seed); W = H = 20.0
    cell, levels, I = build(C, ob)
    Q = [shapes(t) for t in levels]
    par = np.repeat(P, 32, axis=0); n = len(par)
    x, y, th, v = [par[:, i].astype(np.float32).copy() for i in range(4)]
    a = rng.uniform(-5, 5, n).astype(np.float32); st = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    dur = rng.uniform(0.05, 1.05, n).astype(np.float32); dt = dur / np.float32(numDisc); tanS = np.tan(st)
    live_true = np.ones(n, bool); live_spec = np.ones(n, bool)
    amb_steps = np.zeros(n, np.int32); needed = np.zeros(n, np.int32); lv = []
    cxp, cyp = cell(x), cell(y)
    for i in range(numDisc):
        px, py = x.copy(), y.copy()
        x = (x + dt * v * np.cos(th)).astype(np.float32); y = (y + dt * v * np.sin(th)).astype(np.float32)
        oob = (x <= 0) | (x >= W) | (y <= 0) | (y >= H)
        th = (th + dt * v * tanS).astype(np.float32); v = (v + a * dt).astype(np.float32)
        bnx, bxx, bny, bxy = np.minimum(px, x), np.maximum(px, x), np.minimum(py, y), np.maximum(py, y)
        hit = np.zeros(n, bool)
        for lo in range(0, len(ob), 250):
            o = ob[lo:lo + 250]
            hit |= ((bxx[:, None] > o[None, :, 0]) & (o[None, :, 2] > bnx[:, None]) & (bxy[:, None] > o[None, :, 1]) & (o[None, :, 3] > bny[:, None])).any(1)
        cxn, cyn = cell(x), cell(y)
        x0, x1, y0, y1 = np.minimum(cxp, cxn), np.maximum(cxp, cxn), np.minimum(cyp, cyn), np.maximum(cyp, cyn)
        d = np.maximum(x1 - x0, y1 - y0)
        l = np.where(d <= 1, 0, np.ceil(np.log2(np.maximum(d, 1))).astype(np.int64))
        l = np.minimum(l, len(levels) - 1)
        amb = np.zeros(n, bool)
        for L in np.unique(l):
            m = l == L
            X0, X1, Y0, Y1 = x0[m] >> L, x1[m] >> L, y0[m] >> L, y1[m] >> L
            assert ((X1 - X0) <= 1).all() and ((Y1 - Y0) <= 1).all()
            if four_shapes:
                r = np.zeros(m.sum(), bool)
                for dx in (0, 1):
                    for dy in (0, 1):
                        s = ((X1 - X0) == dx) & ((Y1 - Y0) == dy)
                        r[s] = Q[L][(dx, dy)][Y0[s], X0[s]]
                amb[m] = r
            else:
                amb[m] = Q[L][(1, 1)][Y0, X0]
        inside = I[cyn, cxn]
        amb &= ~inside
        a_now = live_spec & ~oob & amb
        amb_steps += a_now; needed += (a_now & live_true); lv.append(l[live_spec].mean())
        live_true &= ~(oob | hit); live_spec &= ~(oob | inside)
        cxp, cyp = cxn, cyn
    ch = amb_steps.reshape(-1, 32).sum(1)
    print(f"C={C} four_shapes={four_shapes}: amb entries/exp {amb_steps.mean():.3f} (needed {needed.mean():.3f}) per chunk mean {ch.mean():.1f} p95 {np.percentile(ch,95):.0f} mean level {np.mean(lv):.2f}")

if __name__ == '__main__':
    P = np.load('bench_data/c2_frontier_sample.npz')['parents'][:6000]
    ob = w.c2_obstacles()
    run(256, ob, P, True); run(512, ob, P, False); run(512, ob, P, True); run(1024, ob, P, False); run(1024, ob, P, True)
