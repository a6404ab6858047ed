"""Warp-level instruction-cost model of the collision part of phase A on the config-2 frontier sample:
current per-step walk vs deferred exact tests (own-entry loop / flattened queue).  Sizing only."""
import sys, numpy as np
sys.path.insert(0, '.')
from cudasbmp_b200 import workloads as w
sys.path.insert(0, 'experiments/deferred_walk')
from stats3 import build, shapes

FIX_ROW, TRIP, FIX_STEP = 10, 28, 28

def csr(ob, C=48, W=20.0):
    inv = np.float32(C / W)
    cell = lambda v: np.clip(np.floor(v * inv).astype(np.int64), 0, C - 1)
    cnt = np.zeros((C, C), np.int64)
    for o in ob:
        cnt[cell(o[1]):cell(o[3]) + 1, cell(o[0]):cell(o[2]) + 1] += 1
    start = np.concatenate([[0], cnt.ravel().cumsum()])
    return cell, start, C

def walk_cost(act, rows_lo, rows_hi, cx0, cx1, start, C):
    """act: [warps,32] bool; returns warp-level cost [warps] of the row/item loops"""
    cost = np.zeros(act.shape[0])
    nrows = np.where(act, rows_hi - rows_lo + 1, 0)
    for r in range(int(nrows.max()) if nrows.size else 0):
        m = act & (nrows > r)
        row = (rows_lo + r) * C
        k = start[np.where(m, row + cx0, 0)]; e = start[np.where(m, row + cx1 + 1, 0)]
        trips = np.where(m, (e - k + 3) // 4, 0)
        cost += np.where(m.any(1), FIX_ROW + TRIP * trips.max(1), 0)
    return cost

def run(ob, P, Cf=256, four=True, numDisc=10, seed=1):
    rng = np.random.default_rng(seed); W = H = 20.0
    ccell, start, Cc = csr(ob)
    fcell, levels, I = build(Cf, ob)
    Q = [shapes(t) for t in levels]
    par = np.repeat(P, 32, axis=0); n = len(par); nw = n // 32
    x, y, th, v = [par[:, i].astype(np.float32).copy() for i in range(4)]
    a = rng.uniform(-5, 5, n).astype(np.float32); st = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    dur = rng.uniform(0.05, 1.05, n).astype(np.float32); dt = dur / np.float32(numDisc); tanS = np.tan(st)
    live_true = np.ones(n, bool); live_spec = np.ones(n, bool)
    cur_cost = np.zeros(nw); cur_trips = np.zeros(nw); spec_trips = np.zeros(nw)
    ent = []   # per step: (amb mask incl. speculation, needed mask, rows_lo, rows_hi, cx0, cx1)
    fxp, fyp = fcell(x), fcell(y)
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
        c0x, c1x, c0y, c1y = ccell(bnx), ccell(bxx), ccell(bny), ccell(bxy)
        # current scheme: lanes live (true semantics) and in bounds walk their ranges
        act = (live_true & ~oob).reshape(nw, 32)
        anylive = live_true.reshape(nw, 32).any(1)
        cur_trips += anylive
        cur_cost += np.where(anylive, FIX_STEP, 0) + walk_cost(act, c0y.reshape(nw, 32), c1y.reshape(nw, 32), c0x.reshape(nw, 32), c1x.reshape(nw, 32), start, Cc)
        # deferred: fine query
        fxn, fyn = fcell(x), fcell(y)
        x0, x1, y0, y1 = np.minimum(fxp, fxn), np.maximum(fxp, fxn), np.minimum(fyp, fyn), np.maximum(fyp, fyn)
        d = np.maximum(x1 - x0, y1 - y0)
        l = np.minimum(np.where(d <= 1, 0, np.ceil(np.log2(np.maximum(d, 1))).astype(np.int64)), len(levels) - 1)
        amb = np.zeros(n, bool)
        for L in np.unique(l):
            m = l == L
            X0, X1, Y0, Y1 = x0[m] >> L, x1[m] >> L, y0[m] >> L, y1[m] >> L
            if four:
                r = np.zeros(m.sum(), bool)
                for dx in (0, 1):
                    for dy in (0, 1):
                        s = ((X1 - X0) == dx) & ((Y1 - Y0) == dy)
                        r[s] = Q[L][(dx, dy)][Y0[s], X0[s]]
                amb[m] = r
            else:
                amb[m] = Q[L][(1, 1)][Y0, X0]
        inside = I[fyn, fxn]
        a_now = live_spec & ~oob & amb & ~inside
        spec_trips += live_spec.reshape(nw, 32).any(1)
        ent.append((a_now.copy(), (a_now & live_true).copy(), c0y, c1y, c0x, c1x))
        live_true &= ~(oob | hit); live_spec &= ~(oob | inside)
        fxp, fyp = fxn, fyn
    print(f"current: step trips/chunk {cur_trips.mean():.2f}, collision warp-inst/chunk {cur_cost.mean():.0f}")
    # own-entry loop: lane processes its needed entries in order
    need = np.stack([e[1] for e in ent])            # [steps, n]
    order = np.cumsum(need, 0) - 1                  # index of the entry within its lane
    maxn = need.sum(0).reshape(nw, 32).max(1)
    own = np.zeros(nw)
    for t in range(int(maxn.max())):
        act = np.zeros(n, bool); ry0 = np.zeros(n, np.int64); ry1 = np.zeros(n, np.int64); rx0 = np.zeros(n, np.int64); rx1 = np.zeros(n, np.int64)
        for s_, e in enumerate(ent):
            m = need[s_] & (order[s_] == t)
            act |= m; ry0[m] = e[2][m]; ry1[m] = e[3][m]; rx0[m] = e[4][m]; rx1[m] = e[5][m]
        A = act.reshape(nw, 32)
        own += np.where(A.any(1), 15 + FIX_STEP, 0) + walk_cost(A, ry0.reshape(nw, 32), ry1.reshape(nw, 32), rx0.reshape(nw, 32), rx1.reshape(nw, 32), start, Cc)
    print(f"Cf={Cf} four={four}: own-entry loop: trips/chunk {maxn.mean():.2f}, exact warp-inst/chunk {own.mean():.0f}  (+ query ~16/step-trip, spec step trips {spec_trips.mean():.2f})")
    # flattened queue: all amb entries (incl. speculation), step-major order, passes of 32
    allm = np.stack([e[0] for e in ent]).reshape(numDisc, nw, 32)
    flat = np.zeros(nw); passes = np.zeros(nw)
    geo = [np.stack([e[k] for e in ent]).reshape(numDisc, nw, 32) for k in (2, 3, 4, 5)]
    for wi in range(0, nw, 1):
        idx = np.argwhere(allm[:, wi, :])
        E = len(idx)
        for p in range(0, E, 32):
            sel = idx[p:p + 32]
            A = np.zeros((1, 32), bool); A[0, :len(sel)] = True
            g = [np.zeros((1, 32), np.int64) for _ in range(4)]
            for k in range(4): g[k][0, :len(sel)] = geo[k][sel[:, 0], wi, sel[:, 1]]
            flat[wi] += 45 + FIX_STEP + walk_cost(A, g[0], g[1], g[2], g[3], start, Cc)[0]
            passes[wi] += 1
    print(f"Cf={Cf} four={four}: flattened: passes/chunk {passes.mean():.2f}, exact warp-inst/chunk {flat.mean():.0f}  (+ query ~16 and push ~10 per step-trip)")

if __name__ == '__main__':
    P = np.load('bench_data/c2_frontier_sample.npz')['parents'][:1500]
    ob = w.c2_obstacles()
    run(ob, P, 256, True)
    run(ob, P, 512, False)
