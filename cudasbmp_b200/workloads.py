"""Synthetic inputs of the BASELINE.json configurations.  Pure numpy, deterministic."""
import numpy as np

# configurations/obstacles/obstacles.csv:1-5 of the reference (minx, miny, maxx, maxy) — the shipped map
C1_OBSTACLES = np.array([[2, 2, 4, 4], [7, 2, 9, 5], [3, 18, 6, 20], [2, 10, 4, 12], [0, 6, 18, 8]], dtype=np.float32)
C1_INIT = np.array([5, 5, 0, 0, 0, 0, 0], dtype=np.float32)        # demos/main.cu:33-39
C1_GOAL = np.array([2, 18, 0, 0, 0, 0, 0], dtype=np.float32)       # demos/main.cu:40-46
C1 = dict(width=20.0, height=20.0, N=16, n=8, numIterations=100, maxTreeSize=30000, numDisc=10, agentLength=1.0,
          goalThreshold=0.5)                                       # demos/main.cu:19-28


def random_boxes(K, side_lo, side_hi, width=20.0, height=20.0, keep_clear=((1.0, 1.0), (19.0, 19.0)), clear=1.0,
                 seed=0xC0FFEE):
    """K axis-aligned boxes: centres uniform in the workspace, sides uniform in [side_lo, side_hi],
    rejected when within `clear` of any keep_clear point."""
    rng = np.random.default_rng(seed)
    out = np.zeros((K, 4), dtype=np.float32)
    k = 0
    while k < K:
        m = 2 * (K - k) + 16
        cx, cy = rng.uniform(0, width, m), rng.uniform(0, height, m)
        sx, sy = rng.uniform(side_lo, side_hi, m), rng.uniform(side_lo, side_hi, m)
        box = np.stack([cx - sx / 2, cy - sy / 2, cx + sx / 2, cy + sy / 2], axis=1)
        ok = np.ones(m, dtype=bool)
        for (px, py) in keep_clear:
            dx = np.maximum(np.maximum(box[:, 0] - px, px - box[:, 2]), 0)
            dy = np.maximum(np.maximum(box[:, 1] - py, py - box[:, 3]), 0)
            ok &= np.hypot(dx, dy) >= clear
        box = box[ok][: K - k]
        out[k:k + len(box)] = box.astype(np.float32)
        k += len(box)
    return out


C2 = dict(width=20.0, height=20.0, N=16, n=32, numIterations=100, maxTreeSize=1 << 20, numDisc=10, agentLength=1.0,
          goalThreshold=0.5)
C2_INIT = np.array([1, 1, 0, 0, 0, 0, 0], dtype=np.float32)
C2_GOAL = np.array([19, 19, 0, 0, 0, 0, 0], dtype=np.float32)


def c2_obstacles(K=1000):
    """config 2: 1k boxes, sides in [0.1, 0.5]."""
    return random_boxes(K, 0.1, 0.5)


C3 = dict(C2, numDisc=40)


def c3_obstacles(K=10000):
    """config 3: 10k boxes of the same total area, sides in [0.03, 0.16]."""
    return random_boxes(K, 0.03, 0.16)


def random_parents(P, obstacles, width=20.0, height=20.0, seed=7):
    """P parent states (rows of 7) uniform in free space, theta in (-pi, pi], v in [-2, 2] (config 5)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((P, 7), dtype=np.float32)
    k = 0
    ob = np.asarray(obstacles, dtype=np.float32).reshape(-1, 4)
    while k < P:
        m = 2 * (P - k) + 16
        x, y = rng.uniform(0.05, width - 0.05, m), rng.uniform(0.05, height - 0.05, m)
        free = np.ones(m, dtype=bool)
        for lo in range(0, len(ob), 2048):
            o = ob[lo:lo + 2048]
            inside = (x[:, None] > o[None, :, 0]) & (x[:, None] < o[None, :, 2]) & \
                     (y[:, None] > o[None, :, 1]) & (y[:, None] < o[None, :, 3])
            free &= ~inside.any(axis=1)
        x, y = x[free][: P - k], y[free][: P - k]
        c = len(x)
        out[k:k + c, 0], out[k:k + c, 1] = x, y
        out[k:k + c, 2] = rng.uniform(-np.pi, np.pi, c)
        out[k:k + c, 3] = rng.uniform(-2, 2, c)
        k += c
    return out


def random_queries(Q, obstacles, width=20.0, height=20.0, min_dist=10.0, seed=0xBA7C4):
    """Q (init, goal) pairs in free space at least min_dist apart (config 4)."""
    pts = random_parents(4 * Q + 64, obstacles, width, height, seed=seed)
    init, goal = [], []
    i = 0
    while len(init) < Q and i + 1 < len(pts):
        a, b = pts[i], pts[i + 1]
        i += 2
        if np.hypot(a[0] - b[0], a[1] - b[1]) >= min_dist:
            ia = np.zeros(7, dtype=np.float32); ia[:2] = a[:2]
            gb = np.zeros(7, dtype=np.float32); gb[:2] = b[:2]
            init.append(ia); goal.append(gb)
    if len(init) < Q:
        raise RuntimeError("not enough separated query pairs")
    return np.stack(init), np.stack(goal)
