"""Statistics for the deferred exact test (clearance map + queued exact AABB tests) on the config-2 frontier sample.
numpy only; random controls (not the Philox stream): this sizes the design, it is not a parity check."""
import sys, numpy as np
sys.path.insert(0, '.')
from cudasbmp_b200 import workloads as w

def run(C=256, K=1000, numDisc=10, obstacles=None, P=None, seed=1):
    rng = np.random.default_rng(seed)
    ob = obstacles
    W = H = 20.0
    inv = np.float32(C / W)
    cell = lambda v: np.clip(np.floor(v * inv).astype(np.int64), 0, C - 1)
    # touch map + inside map
    T = np.zeros((C, C), bool); I = np.zeros((C, C), bool)
    for o in ob:
        x0, y0, x1, y1 = cell(o[0]), cell(o[1]), cell(o[2]), cell(o[3])
        T[y0:y1 + 1, x0:x1 + 1] = True
        if x1 - x0 >= 2 and y1 - y0 >= 2: I[y0 + 1:y1, x0 + 1:x1] = True
    # chebyshev clearance (cap 14)
    clr = np.zeros((C, C), np.int32)
    cur = T.copy()
    for d in range(1, 15):
        nxt = cur.copy()
        nxt[1:, :] |= cur[:-1, :]; nxt[:-1, :] |= cur[1:, :]
        nxt2 = nxt.copy()
        nxt2[:, 1:] |= nxt[:, :-1]; nxt2[:, :-1] |= nxt[:, 1:]
        clr[~nxt2 & (clr == 0) & ~T] = 0
        newly = nxt2 & ~cur
        clr[newly] = d
        cur = nxt2
    clr[~cur] = 14
    par = np.repeat(P, 32, axis=0)
    n = len(par)
    x, y, th, v = [par[:, i].astype(np.float32).copy() for i in range(4)]
    a = rng.uniform(-5, 5, n).astype(np.float32); st = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    dur = rng.uniform(0.05, 1.05, n).astype(np.float32); dt = dur / np.float32(numDisc); tanS = np.tan(st)
    live_true = np.ones(n, bool)      # true semantics (exact early exit)
    live_spec = np.ones(n, bool)      # speculative lanes (stop on bounds / inside-definite only)
    amb_steps = np.zeros(n, np.int32); true_steps = np.zeros(n, np.int32); spec_steps = np.zeros(n, np.int32)
    amb_before_true_exit = np.zeros(n, np.int32)
    per_step_amb = []
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
        d = np.maximum(np.abs(cxn - cxp), np.abs(cyn - cyp))
        cp, cn = clr[cyp, cxp], clr[cyn, cxn]
        inside = I[cyn, cxn]
        amb = (d >= np.maximum(cp, cn)) & ~inside
        true_steps += live_true; spec_steps += live_spec
        a_now = live_spec & ~oob & amb
        amb_steps += a_now
        amb_before_true_exit += (a_now & live_true)
        per_step_amb.append(a_now.sum() / max(1, live_spec.sum()))
        live_true &= ~(oob | hit)
        live_spec &= ~(oob | inside)
        cxp, cyp = cxn, cyn
    ch = amb_steps.reshape(-1, 32).sum(1)
    print(f"C={C}: true steps/exp {true_steps.mean():.2f} spec steps/exp {spec_steps.mean():.2f} valid {live_true.mean():.3f} | amb entries/exp {amb_steps.mean():.3f} "
          f"(needed {amb_before_true_exit.mean():.3f}) per chunk mean {ch.mean():.1f} p50 {np.median(ch):.0f} p95 {np.percentile(ch,95):.0f} max {ch.max()} | amb/live-step {np.mean(per_step_amb):.3f}")

if __name__ == '__main__':
    P = np.load('bench_data/c2_frontier_sample.npz')['parents'][:6000]
    ob = w.c2_obstacles()
    for C in (128, 192, 256, 384, 512):
        run(C, obstacles=ob, P=P)
