"""Cost model: 2x2-block obstacle lists (deduplicated, one list per step bbox) vs the per-cell CSR rows."""
import sys, numpy as np
sys.path.insert(0, '.')
from cudasbmp_b200 import workloads as w

def cells_of(C, W=20.0):
    inv = np.float32(C / W)
    return lambda v: np.clip(np.floor(v * inv).astype(np.int64), 0, C - 1)

def build(ob, C):
    cell = cells_of(C)
    cnt = np.zeros((C, C), np.int64); blk = np.zeros((C, C), np.int64)
    for o in ob:
        x0, y0, x1, y1 = cell(o[0]), cell(o[1]), cell(o[2]), cell(o[3])
        cnt[y0:y1 + 1, x0:x1 + 1] += 1
        blk[max(y0 - 1, 0):y1 + 1, max(x0 - 1, 0):x1 + 1] += 1      # block (bx,by) covers cells bx..bx+1, by..by+1
    return cell, np.concatenate([[0], cnt.ravel().cumsum()]), blk

def run(ob, P, grids, numDisc=10, seed=1):
    rng = np.random.default_rng(seed); W = H = 20.0
    G = {C: build(ob, C) for C in grids}
    par = np.repeat(P, 32, axis=0); n = len(par); nw = n // 32
    x, y, th, v = [par[:, i].astype(np.float32).copy() for i in range(4)]
    a = rng.uniform(-5, 5, n).astype(np.float32); st = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    dur = rng.uniform(0.05, 1.05, n).astype(np.float32); dt = dur / np.float32(numDisc); tanS = np.tan(st)
    live = np.ones(n, bool)
    res = {C: dict(rowtrips=0, itemtrips=0, btrips=0, big=0, blane=0, clane=0) for C in grids}
    steps = 0; trips = 0
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
        act = (live & ~oob); steps += act.sum(); trips += live.reshape(nw, 32).any(1).sum()
        for C in grids:
            cell, start, blk = G[C]; R = res[C]
            c0x, c1x, c0y, c1y = cell(bnx), cell(bxx), cell(bny), cell(bxy)
            nrows = np.where(act, c1y - c0y + 1, 0)
            for r in range(int(nrows.max())):
                m = act & (nrows > r)
                row = (c0y + r) * C
                k = start[np.where(m, row + c0x, 0)]; e = start[np.where(m, row + c1x + 1, 0)]
                t = np.where(m, (e - k + 3) // 4, 0).reshape(nw, 32)
                R['rowtrips'] += m.reshape(nw, 32).any(1).sum(); R['itemtrips'] += t.max(1).sum(); R['clane'] += t.sum()
            big = act & (((c1x - c0x) > 1) | ((c1y - c0y) > 1)); R['big'] += big.sum()
            # block walk: blocks at stride 2 over the cell rectangle
            tt = np.zeros(n, np.int64)
            for by in range(0, 8, 2):
                for bx in range(0, 8, 2):
                    m = act & (c0x + bx <= c1x) & (c0y + by <= c1y)
                    if not m.any(): continue
                    L = blk[np.where(m, np.minimum(c0y + by, C - 1), 0), np.where(m, np.minimum(c0x + bx, C - 1), 0)]
                    tt += np.where(m, (L + 3) // 4, 0)
            R['btrips'] += tt.reshape(nw, 32).max(1).sum(); R['blane'] += tt.sum()
        live &= ~(oob | hit)
    for C in grids:
        R = res[C]; _, start, blk = G[C]
        print(f"C={C:3d}: cell items {start[-1]} block items {blk.sum()} ({blk.sum()*2/1024:.0f}+{C*C*2/1024:.0f}+16 KB) | per step-trip: CSR rows {R['rowtrips']/trips:.2f} item trips {R['itemtrips']/trips:.2f} (lane-trips/step {R['clane']/steps:.2f}) -> {16*R['rowtrips']/trips+27*R['itemtrips']/trips:.0f} inst | block trips {R['btrips']/trips:.2f} (lane-trips/step {R['blane']/steps:.2f}) big {R['big']/steps:.4f} -> {12+32*R['btrips']/trips:.0f} inst")

if __name__ == '__main__':
    P = np.load('bench_data/c2_frontier_sample.npz')['parents'][:1500]
    run(w.c2_obstacles(), P, (32, 40, 48, 56, 64, 80))
