"""Cost model of the CURRENT per-step walk for several cull-grid resolutions and loop shapes."""
import sys, numpy as np
sys.path.insert(0, '.')
from cudasbmp_b200 import workloads as w
FIX_ROW, TRIP, FIX_STEP = 10, 28, 28

def csr(ob, C, W=20.0):
    inv = np.float32(C / W)
    cell = lambda v: np.clip(np.floor(v * inv).astype(np.int64), 0, C - 1)
    cnt = np.zeros((C, C), np.int64)
    for o in ob:
        cnt[cell(o[1]):cell(o[3]) + 1, cell(o[0]):cell(o[2]) + 1] += 1
    return cell, np.concatenate([[0], cnt.ravel().cumsum()]), cnt

def run(ob, P, grids, numDisc=10, seed=1, per=4):
    rng = np.random.default_rng(seed); W = H = 20.0
    G = {C: csr(ob, C) for C in grids}
    par = np.repeat(P, 32, axis=0); n = len(par); nw = n // 32
    x, y, th, v = [par[:, i].astype(np.float32).copy() for i in range(4)]
    a = rng.uniform(-5, 5, n).astype(np.float32); st = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    dur = rng.uniform(0.05, 1.05, n).astype(np.float32); dt = dur / np.float32(numDisc); tanS = np.tan(st)
    live = np.ones(n, bool)
    nested = {C: np.zeros(nw) for C in grids}; flat = {C: np.zeros(nw) for C in grids}; lanes_any = {C: [] for C in grids}; pairs = {C: 0 for C in grids}
    steps = 0
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
        act = (live & ~oob); A = act.reshape(nw, 32); anylive = live.reshape(nw, 32).any(1)
        steps += act.sum()
        for C in grids:
            cell, start, _ = G[C]
            c0x, c1x, c0y, c1y = cell(bnx), cell(bxx), cell(bny), cell(bxy)
            nrows = np.where(act, c1y - c0y + 1, 0)
            tot = np.zeros(n, np.int64); cost = np.zeros(nw)
            for r in range(int(nrows.max())):
                m = act & (nrows > r)
                row = (c0y + r) * C
                k = start[np.where(m, row + c0x, 0)]; e = start[np.where(m, row + c1x + 1, 0)]
                trips = np.where(m, (e - k + per - 1) // per, 0)
                tot += trips
                cost += np.where(m.reshape(nw, 32).any(1), FIX_ROW + (TRIP * per // 4) * trips.reshape(nw, 32).max(1), 0)
            nested[C] += np.where(anylive, FIX_STEP, 0) + cost
            flat[C] += np.where(anylive, FIX_STEP + 8, 0) + (TRIP * per // 4 + 4) * tot.reshape(nw, 32).max(1)
            lanes_any[C].append((tot > 0).sum() / max(1, act.sum())); pairs[C] += tot.sum() * per
        live &= ~(oob | hit)
    for C in grids:
        _, start, cnt = G[C]
        print(f"C={C:4d} per={per}: items {start[-1]:6d} ({start[-1]*16/1024:.0f} KB + {C*C*4/1024:.0f} KB)  lanes with items {np.mean(lanes_any[C]):.2f}  pairs/step {pairs[C]/steps:.2f}  nested {nested[C].mean():.0f}  flat-loop {flat[C].mean():.0f} warp-inst/chunk")

if __name__ == '__main__':
    P = np.load('bench_data/c2_frontier_sample.npz')['parents'][:1500]
    ob = w.c2_obstacles()
    run(ob, P, (32, 48, 64, 96, 128, 192, 256))
    run(ob, P, (48, 64, 96, 128, 192), per=2)
