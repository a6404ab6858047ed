"""Same as stats.py but with an exact cell-rectangle query on the touch bitmap (summed-area table) instead of the clearance ball."""
import sys, numpy as np
sys.path.insert(0, '.')
from cudasbmp_b200 import workloads as w

def run(C, ob, P, numDisc=10, seed=1, defer=True):
    rng = np.random.default_rng(seed)
    W = H = 20.0
    inv = np.float32(C / W)
    cell = lambda v: np.clip(np.floor(v * inv).astype(np.int64), 0, C - 1)
    T = np.zeros((C, C), bool); I = np.zeros((C, C), bool)
    for o in ob:
        x0, y0, x1, y1 = cell(o[0]), cell(o[1]), cell(o[2]), cell(o[3])
        T[y0:y1 + 1, x0:x1 + 1] = True
        if x1 - x0 >= 2 and y1 - y0 >= 2: I[y0 + 1:y1, x0 + 1:x1] = True
    S = np.zeros((C + 1, C + 1), np.int64); S[1:, 1:] = T.cumsum(0).cumsum(1)
    par = np.repeat(P, 32, axis=0); n = len(par)
    x, y, th, v = [par[:, i].astype(np.float32).copy() for i in range(4)]
    a = rng.uniform(-5, 5, n).astype(np.float32); st = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    dur = rng.uniform(0.05, 1.05, n).astype(np.float32); dt = dur / np.float32(numDisc); tanS = np.tan(st)
    live_true = np.ones(n, bool); live_spec = np.ones(n, bool)
    amb_steps = np.zeros(n, np.int32); needed = np.zeros(n, np.int32); spans = []
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
        cnt = S[y1 + 1, x1 + 1] - S[y0, x1 + 1] - S[y1 + 1, x0] + S[y0, x0]
        inside = I[cyn, cxn]
        amb = (cnt > 0) & ~inside
        a_now = live_spec & ~oob & amb
        amb_steps += a_now; needed += (a_now & live_true)
        spans.append(np.maximum(x1 - x0, y1 - y0)[live_spec].mean())
        live_true &= ~(oob | hit); live_spec &= ~(oob | inside)
        cxp, cyp = cxn, cyn
    ch = amb_steps.reshape(-1, 32).sum(1)
    print(f"C={C}: rect amb entries/exp {amb_steps.mean():.3f} (needed {needed.mean():.3f}) per chunk mean {ch.mean():.1f} p95 {np.percentile(ch,95):.0f}  mean span {np.mean(spans):.2f} cells")

if __name__ == '__main__':
    P = np.load('bench_data/c2_frontier_sample.npz')['parents'][:6000]
    ob = w.c2_obstacles()
    for C in (48, 128, 256, 512, 1024, 4096):
        run(C, ob, P)
