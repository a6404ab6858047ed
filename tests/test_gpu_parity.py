"""GPU parity tests proper: the CUDA path, called through the C ABI, against
  - the CPU oracle, stage by stage on identical inputs (tests/parity.py);
  - the committed golden vectors made from the reference's own host build;
  - the reference's own CUDA kernels (oracle/_ref/libref_gpu.so, unmodified, cuRAND Philox), bit for bit;
and size-independent properties at the BASELINE.json sizes."""
import os
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from cudasbmp_b200 import kgmt as K          # noqa: E402
from cudasbmp_b200 import workloads as w     # noqa: E402
from tests.parity import MARGIN, TOL_REL, bits, check_iteration, state_tolerance   # noqa: E402


def _plan(cfg, obstacles, **kw):
    p = K.KGMT(**cfg, **kw)
    p.set_obstacles(obstacles)
    return p


def _propagate(plan, parents, children, key, M):
    plan.stage_propagate(parents, children, key, 0)
    return (plan.export(K.ARR_UNEXPLORED)[:M].copy(), plan.export(K.ARR_U_VALID)[:M].copy(),
            plan.export(K.ARR_U_U3)[:M].copy(), plan.export(K.ARR_U_R1)[:M].copy(), plan.export(K.ARR_U_R2)[:M].copy())


# ------------------------------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("name,mode", [("propagate_c1.npz", K.COLLIDE_GRID), ("propagate_c1.npz", K.COLLIDE_BRUTE),
                                       ("propagate_c2.npz", K.COLLIDE_GRID), ("propagate_c2.npz", K.COLLIDE_BRUTE),
                                       ("propagate_c3.npz", K.COLLIDE_GRID), ("propagate_c3.npz", K.COLLIDE_BRUTE),
                                       ("propagate_root.npz", K.COLLIDE_GRID)])
def test_propagate_against_reference_golden(golden_dir, oracle, name, mode):
    """Golden = reference host build (glibc trig, no FMA): tolerance on states, flags equal away from boundaries."""
    g = np.load(os.path.join(golden_dir, name))
    parents, pof = g["parents"], g["parent_of"]
    P, M = len(parents), len(pof)
    children = M // P
    nd = int(g["num_disc"])
    cfg = dict(w.C1, numDisc=nd, maxTreeSize=max(M, 64))
    plan = _plan(cfg, g["obstacles"], collision_mode=mode, record_candidates=True)
    x1, valid, u3, _, _ = _propagate(plan, parents, children, int(g["key"]), M)
    assert (bits(u3) == bits(g["u3"])).all()
    # controls: duration has no contraction; a (u*10-5) and steering (u*2*pi-pi) are one FMA on the GPU, two
    # roundings in the host build: at most 1 ulp of 5 / of pi apart
    assert (bits(x1[:, 6]) == bits(g["x1"][:, 6])).all()
    assert np.abs(x1[:, 4] - g["x1"][:, 4]).max() <= 5e-7
    assert np.abs(x1[:, 5] - g["x1"][:, 5]).max() <= 5e-7
    _, _, _, margin = oracle.propagate_batch(parents, pof, int(g["key"]), 0, nd, 1.0, g["obstacles"], 20.0, 20.0,
                                             oracle.MATH_HOST)
    err = np.abs(x1[:, :4].astype(np.float64) - g["x1"][:, :4]) / np.maximum(1.0, np.abs(g["x1"][:, :4]))
    off = (valid != g["valid"]) | (err.max(axis=1) > state_tolerance(g["x1"], parents[pof, 2], nd))
    assert (margin[off] <= MARGIN).all(), (int(off.sum()), float(margin[off].max()))
    assert off.mean() <= 0.01


# --------------------------------------------------------------------------- the reference's own CUDA kernels
def _ref_gpu_or_skip(oracle):
    R = oracle.ref_gpu()
    if R is None:
        pytest.skip("oracle/_ref/libref_gpu.so not built (needs /root/reference at build time)")
    return R


@pytest.mark.parametrize("mode", [K.COLLIDE_GRID, K.COLLIDE_BRUTE])
@pytest.mark.parametrize("case", ["c1", "c2"])
def test_bit_exact_against_reference_cuda_kernels(oracle, mode, case):
    """propagateG (KGMT.cu:341-414) of the reference, compiled unmodified for sm_100a with Philox states, on the
    same parents and random streams: states, controls, flags, counters identical bit for bit."""
    _ref_gpu_or_skip(oracle)
    obstacles = w.C1_OBSTACLES if case == "c1" else w.c2_obstacles(1000)
    N, n = 16, 8
    P, children, key = 512, 32, 4711
    M = P * children
    parents = w.random_parents(P, obstacles, seed=17)
    cfg = dict(w.C1, maxTreeSize=M)
    plan = _plan(cfg, obstacles, collision_mode=mode, record_candidates=True)
    x1, valid, u3, r1, r2 = _propagate(plan, parents, children, key, M)
    c1, c2 = N * N, N * N * n * n
    maps = {k: np.zeros(c1 if k.startswith("R1") else c2, dtype=np.int32)
            for k in ("R1", "R2", "R1Valid", "R2Valid", "R1Invalid", "R2Invalid", "R1Avail", "R2Avail")}
    unx, upar, gnew, _ = oracle.ref_gpu_expand(1, children, parents, np.arange(P, dtype=np.int32), maps,
                                               np.ones(c1, dtype=np.float32), N, n, plan.R1Size_, plan.R2Size_, 10, 1.0,
                                               obstacles, 20.0, 20.0, key)
    assert (bits(unx) == bits(x1)).all(), "states/controls differ from the reference kernel"
    assert (upar == np.arange(M) // children).all()
    # with scores 1.0 and no R2 cell available the reference accepts exactly its valid candidates (KGMT.cu:396)
    inside = r1 >= 0
    assert (gnew[inside] == valid[inside]).all(), "collision flags differ from the reference kernel"
    # counters: integer adds commute, so the reference's arrays are deterministic on in-range cells
    ok = inside & (r2 >= 0)
    assert (np.bincount(r1[inside], minlength=c1) == maps["R1"]).all()
    assert (np.bincount(r1[inside & (valid == 1)], minlength=c1) == maps["R1Valid"]).all()
    assert (np.bincount(r1[inside & (valid == 0)], minlength=c1) == maps["R1Invalid"]).all()
    assert (np.bincount(r2[ok], minlength=c2) == maps["R2"]).all()
    assert (np.bincount(r2[ok & (valid == 1)], minlength=c2) == maps["R2Valid"]).all()
    assert (np.bincount(r2[ok & (valid == 0)], minlength=c2) == maps["R2Invalid"]).all()


def test_bit_exact_against_reference_kernels_large_and_extreme_headings(oracle):
    """The same comparison at scale (262 144 edges on the config-2 map) and with parents whose heading is far outside
    the fast range of the trigonometric range reduction (|theta| up to 1e7: libdevice's Payne-Hanek path, which
    sincosf shares with the reference's separate sinf / cosf) and with large speeds (long steps, several cull cells
    per step)."""
    _ref_gpu_or_skip(oracle)
    obstacles = w.c2_obstacles(1000)
    N, n = 16, 8
    P, children, key = 8192, 32, 90210
    M = P * children
    parents = w.random_parents(P, obstacles, seed=23)
    rng = np.random.default_rng(5)
    parents[: P // 2, 2] = (rng.uniform(-1.0, 1.0, P // 2) * 10.0 ** rng.integers(2, 8, P // 2)).astype(np.float32)
    parents[P // 4: 3 * P // 4, 3] = rng.uniform(-40.0, 40.0, P // 2).astype(np.float32)
    cfg = dict(w.C1, maxTreeSize=M)
    c1, c2 = N * N, N * N * n * n
    for mode in (K.COLLIDE_GRID, K.COLLIDE_BRUTE):
        plan = _plan(cfg, obstacles, collision_mode=mode, record_candidates=True)
        x1, valid, u3, r1, r2 = _propagate(plan, parents, children, key, M)
        maps = {k: np.zeros(c1 if k.startswith("R1") else c2, dtype=np.int32)
                for k in ("R1", "R2", "R1Valid", "R2Valid", "R1Invalid", "R2Invalid", "R1Avail", "R2Avail")}
        unx, upar, gnew, _ = oracle.ref_gpu_expand(1, children, parents, np.arange(P, dtype=np.int32), maps,
                                                   np.ones(c1, dtype=np.float32), N, n, plan.R1Size_, plan.R2Size_, 10, 1.0,
                                                   obstacles, 20.0, 20.0, key)
        assert (bits(unx) == bits(x1)).all(), "states/controls differ from the reference kernel"
        inside = r1 >= 0
        assert (gnew[inside] == valid[inside]).all(), "collision flags differ from the reference kernel"
        assert (np.bincount(r1[inside & (valid == 1)], minlength=c1) == maps["R1Valid"]).all()
        assert 0.05 < valid.mean() < 0.95


def test_insertion_bit_exact_against_reference_updateG(oracle):
    """scan + findInd + updateG of the reference (KGMT.cu:222-245,540-593) vs our ordered insertion."""
    _ref_gpu_or_skip(oracle)
    import ctypes as C
    cfg = dict(w.C1, maxTreeSize=4096)
    plan = _plan(cfg, w.C1_OBSTACLES, record_candidates=True)
    plan.begin(w.C1_INIT, w.C1_GOAL)
    plan.iterate()
    T0 = plan.result()["tree_size"]
    tree0, par0, cost0 = plan.export(K.ARR_SAMPLES).copy(), plan.export(K.ARR_PARENT).copy(), plan.export(K.ARR_COSTS).copy()
    st = plan.iterate()
    M = st["candidates"]
    cap = 4096
    gnew = np.zeros(cap, dtype=np.uint8); gnew[:M] = plan.export(K.ARR_U_ACCEPT)[:M]
    unx = plan.export(K.ARR_UNEXPLORED).copy()
    upar = plan.export(K.ARR_U_PARENT).copy(); upar[upar < 0] = 0
    G = np.zeros(cap, dtype=np.uint8)
    ctg = np.zeros(1, dtype=np.float32)
    R = oracle.ref_gpu()
    k = R.ref_gpu_insert(cap, gnew.ctypes.data_as(oracle.u8p), unx.ctypes.data_as(oracle.f32p),
                         upar.ctypes.data_as(oracle.i32p), T0, tree0.ctypes.data_as(oracle.f32p),
                         par0.ctypes.data_as(oracle.i32p), cost0.ctypes.data_as(oracle.f32p), G.ctypes.data_as(oracle.u8p),
                         np.ascontiguousarray(w.C1_GOAL).ctypes.data_as(oracle.f32p), 0.5, ctg.ctypes.data_as(oracle.f32p))
    assert k == st["accepted"]
    T1 = st["tree_size"]
    assert (bits(tree0[:T1]) == bits(plan.export(K.ARR_SAMPLES)[:T1])).all()
    assert (par0[:T1] == plan.export(K.ARR_PARENT)[:T1]).all()
    assert (bits(cost0[:T1]) == bits(plan.export(K.ARR_COSTS)[:T1])).all()


def test_scores_against_reference_updateR1(oracle):
    """updateR1 of the reference (N = 16 only; pow() and cub order differ in the last bits -> tolerance)."""
    _ref_gpu_or_skip(oracle)
    plan = _plan(w.C1, w.C1_OBSTACLES, record_candidates=True, seed=9)
    plan.begin(w.C1_INIT, w.C1_GOAL)
    for _ in range(4):
        plan.iterate()
    plan.stage_scores()
    m = {k: plan.export(i) for k, i in (("A1", K.ARR_R1AVAIL), ("A2", K.ARR_R2AVAIL), ("V", K.ARR_R1VALID),
                                          ("I", K.ARR_R1INVALID), ("R", K.ARR_R1))}
    ours = plan.export(K.ARR_R1SCORE)
    ref = np.zeros(256, dtype=np.float32)
    thr = np.zeros(1, dtype=np.float32)
    R = oracle.ref_gpu()
    rc = R.ref_gpu_scores(*[m[k].ctypes.data_as(oracle.i32p) for k in ("A1", "A2", "V", "I", "R")], 8,
                          float(plan.R2Size_), ref.ctypes.data_as(oracle.f32p), thr.ctypes.data_as(oracle.f32p))
    assert rc == 0
    np.testing.assert_allclose(ours, ref, rtol=2e-6, atol=0)


# ------------------------------------------------------------------------------ oracle, stage by stage
@pytest.mark.parametrize("mode", [K.COLLIDE_GRID, K.COLLIDE_BRUTE])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_c1_iterations_stage_by_stage(oracle, mode, seed):
    """The reference demo (config C1) through mode 1 and mode 2 iterations until it stops."""
    plan = _plan(w.C1, w.C1_OBSTACLES, collision_mode=mode, record_candidates=True, seed=seed)
    plan.begin(w.C1_INIT, w.C1_GOAL)
    rep, modes = [], set()
    for _ in range(40):
        st = check_iteration(plan, oracle, w.C1_OBSTACLES, w.C1, w.C1_GOAL, seed, rep)
        if st["candidates"]:
            modes.add(st["mode"])
        if st["stop"] != 0:
            break
    assert st["stop"] in (1, 2, 3, 4)
    assert 1 in modes
    print(rep[-1], modes)


def test_c2_small_tree_stage_by_stage(oracle):
    """config C2 map (1k obstacles, N=16 n=32) with a small tree so the full/prefix expansion modes are reached."""
    cfg = dict(w.C2, maxTreeSize=6000, numIterations=12)
    obs = w.c2_obstacles(1000)
    plan = _plan(cfg, obs, record_candidates=True, seed=11)
    plan.begin(w.C2_INIT, w.C2_GOAL)
    modes = set()
    for _ in range(12):
        st = check_iteration(plan, oracle, obs, cfg, w.C2_GOAL, 11)
        if st["candidates"]:
            modes.add(st["mode"])
        if st["stop"] != 0:
            break
    assert 2 in modes or 3 in modes or st["stop"] == 1, (modes, st)


def test_big_N_without_histograms(oracle):
    """N = 96 (> the shared-memory histogram budget): R1 counters go through global atomics."""
    cfg = dict(w.C1, N=96, n=2, maxTreeSize=8000, numIterations=6)
    plan = _plan(cfg, w.C1_OBSTACLES, record_candidates=True, seed=4)
    assert plan.config()["r1_hist"] == 0
    plan.begin(w.C1_INIT, w.C1_GOAL)
    for _ in range(6):
        st = check_iteration(plan, oracle, w.C1_OBSTACLES, cfg, w.C1_GOAL, 4)
        if st["stop"] != 0:
            break


def test_edge_cases(oracle):
    # no obstacles at all
    cfg = dict(w.C1, maxTreeSize=3000, numIterations=5)
    empty = np.zeros((0, 4), dtype=np.float32)
    plan = _plan(cfg, empty, record_candidates=True, seed=2)
    plan.begin(w.C1_INIT, w.C1_GOAL)
    for _ in range(5):
        if check_iteration(plan, oracle, empty, cfg, w.C1_GOAL, 2)["stop"] != 0:
            break
    # the whole workspace is one obstacle: every moving edge collides, the frontier dies
    wall = np.array([[-1, -1, 21, 21]], dtype=np.float32)
    plan = _plan(cfg, wall, record_candidates=True, seed=2)
    plan.begin(w.C1_INIT, w.C1_GOAL)
    st = check_iteration(plan, oracle, wall, cfg, w.C1_GOAL, 2)
    st = check_iteration(plan, oracle, wall, cfg, w.C1_GOAL, 2) if st["stop"] == 0 else st
    assert plan.result()["stop"] in (3, 4) or plan.result()["tree_size"] < 64
    # tree of capacity 1: nothing to do
    plan = _plan(dict(cfg, maxTreeSize=1), w.C1_OBSTACLES, seed=2)
    r = plan.plan(w.C1_INIT, w.C1_GOAL)
    assert r["stop"] == 2 and r["tree_size"] == 1 and r["expansions"] == 0
    # zero iterations
    plan = _plan(dict(cfg, numIterations=0), w.C1_OBSTACLES, seed=2)
    r = plan.plan(w.C1_INIT, w.C1_GOAL)
    assert r["stop"] == 3 and r["tree_size"] == 1
    # root already inside the goal disc: the reference only tests inserted nodes (KGMT.cu:589)
    plan = _plan(cfg, w.C1_OBSTACLES, seed=2)
    plan.begin(w.C1_INIT, w.C1_INIT)
    assert plan.result()["stop"] == 0
    # iterate before begin -> error, not a crash
    plan = _plan(cfg, w.C1_OBSTACLES, seed=2)
    with pytest.raises(K.KgmtError):
        plan.iterate()


def test_iteration_shape_never_exceeds_tree_or_candidate_capacity():
    """forced children persist across iterations and max_candidates may be smaller than the tree; neither may
    produce an iteration that overruns the candidate staging or the tree (expansion_shape clamps to both)."""
    obs = w.C1_OBSTACLES
    # forced children, several iterations in one launch: falls back to the reference policy once 8 per node no longer fit
    plan = _plan(dict(w.C1, maxTreeSize=6000, numIterations=30), obs, record_candidates=True, seed=3)
    nodes = w.random_parents(64, obs, seed=3)
    plan.seed_frontier(nodes, w.C1_GOAL)
    plan.set_children(8)
    seen = []
    for _ in range(30):
        st = plan.iterate()
        seen.append((st["mode"], st["children"], st["candidates"], st["tree_size"]))
        assert st["candidates"] <= 6000 and st["tree_size"] <= 6000, seen
        if st["stop"] != 0:
            break
    assert seen[0][0] == 4 and seen[0][1] == 8
    plan.set_children(0)
    # the same through the multi-iteration launch
    plan2 = _plan(dict(w.C1, maxTreeSize=6000, numIterations=30), obs, seed=3)
    plan2.seed_frontier(nodes, w.C1_GOAL)
    plan2.set_children(8)
    st2 = plan2.iterate_many(30)
    assert st2["tree_size"] == seen[-1][3] and plan2.result()["iterations"] == len(seen)
    # a candidate staging smaller than the tree: every iteration fits it, the plan still runs to a regular stop
    small = _plan(dict(w.C1), obs, record_candidates=True, seed=4, max_candidates=2048)
    small.begin(w.C1_INIT, w.C1_GOAL)
    for _ in range(100):
        st = small.iterate()
        assert st["candidates"] <= 2048 and st["tree_size"] <= 30000, st
        if st["stop"] != 0:
            break
    assert st["stop"] in (1, 2, 3, 4)
    r = _plan(dict(w.C1), obs, seed=4, max_candidates=2048).plan(w.C1_INIT, w.C1_GOAL)
    assert (r["tree_size"], r["stop"]) == (st["tree_size"], st["stop"])


# ------------------------------------------------------------------------------------ whole-plan properties
def _tree_checksum(plan, T):
    h = zlib.crc32(plan.export(K.ARR_SAMPLES)[:T].tobytes())
    h = zlib.crc32(plan.export(K.ARR_PARENT)[:T].tobytes(), h)
    return zlib.crc32(plan.export(K.ARR_COSTS)[:T].tobytes(), h)


def _check_tree(plan, cfg, obstacles, r):
    T = r["tree_size"]
    tree, par, cost = plan.export(K.ARR_SAMPLES)[:T], plan.export(K.ARR_PARENT)[:T], plan.export(K.ARR_COSTS)[:T]
    assert par[0] == -1 and (par[1:] >= 0).all() and (par[1:] < np.arange(1, T)).all()
    assert (np.diff(par[1:]) >= 0).all(), "insertion keeps candidate order, so parents are non-decreasing"
    np.testing.assert_array_equal(cost[1:], (cost[par[1:]] + tree[1:, 6]).astype(np.float32))
    assert ((tree[:, 0] > 0) & (tree[:, 0] < cfg["width"]) & (tree[:, 1] > 0) & (tree[:, 1] < cfg["height"])).all()
    o = np.asarray(obstacles, dtype=np.float32).reshape(-1, 4)
    pick = np.arange(T) if T <= 100000 else np.random.default_rng(0).choice(T, 100000, replace=False)
    for lo in range(0, len(pick), 32768):
        x, y = tree[pick[lo:lo + 32768], 0:1], tree[pick[lo:lo + 32768], 1:2]
        assert not ((x > o[:, 0]) & (x < o[:, 2]) & (y > o[:, 1]) & (y < o[:, 3])).any(), "a tree node lies inside an obstacle"
    if r["stop"] == 1:
        gi = r["goal_index"]
        assert np.float32(cost[gi]) == np.float32(r["cost_to_goal"])


@pytest.mark.parametrize("cfgname", ["c1", "c2"])
def test_plan_equals_stepwise_and_is_deterministic(cfgname):
    """The single cooperative launch (kgmt_plan) builds the same tree, bit for bit, as host-stepped
    iterations, as a second run, and as the exhaustive collision back end."""
    if cfgname == "c1":
        cfg, obs, init, goal = w.C1, w.C1_OBSTACLES, w.C1_INIT, w.C1_GOAL
    else:
        cfg, obs, init, goal = w.C2, w.c2_obstacles(1000), w.C2_INIT, w.C2_GOAL
    a = _plan(cfg, obs, seed=21)
    ra = a.plan(init, goal)
    _check_tree(a, cfg, obs, ra)
    ca = _tree_checksum(a, ra["tree_size"])
    rb = a.plan(init, goal)                                     # re-plan on the same context (kgmt_reset path)
    assert (rb["tree_size"], rb["iterations"], rb["stop"]) == (ra["tree_size"], ra["iterations"], ra["stop"])
    assert _tree_checksum(a, rb["tree_size"]) == ca
    b = _plan(cfg, obs, seed=21)
    b.begin(init, goal)
    while b.iterate()["stop"] == 0:
        pass
    rs = b.result()
    assert (rs["tree_size"], rs["iterations"], rs["stop"], rs["expansions"]) == \
           (ra["tree_size"], ra["iterations"], ra["stop"], ra["expansions"])
    assert _tree_checksum(b, rs["tree_size"]) == ca
    c = _plan(cfg, obs, seed=21, collision_mode=K.COLLIDE_BRUTE)
    rc = c.plan(init, goal)
    assert rc["tree_size"] == ra["tree_size"] and _tree_checksum(c, rc["tree_size"]) == ca
    for k in (K.ARR_R1, K.ARR_R1VALID, K.ARR_R1INVALID, K.ARR_R1AVAIL, K.ARR_R2, K.ARR_R2VALID, K.ARR_R2INVALID,
              K.ARR_R2AVAIL):
        assert (a.export(k) == b.export(k)).all() and (a.export(k) == c.export(k)).all()
    # counters are consistent: R = valid + invalid (+1 for the root's cell, KGMT.cu:94,97)
    R1, V, I = a.export(K.ARR_R1), a.export(K.ARR_R1VALID), a.export(K.ARR_R1INVALID)
    assert (R1 == V + I).all()
    assert int(a.export(K.ARR_R2).sum()) <= ra["expansions"]
    if ra["stop"] == 1:
        path = a.extract_path()
        assert np.allclose(path[0, :2], init[:2]) and np.hypot(*(path[-1, :2] - goal[:2])) < cfg["goalThreshold"]


def test_c3_stress_map_grid_equals_brute():
    """config C3: 10k obstacles, 40 integration steps; grid-culled == exhaustive, bit for bit."""
    obs = w.c3_obstacles(10000)
    parents = w.random_parents(256, obs, seed=5)
    cfg = dict(w.C3, maxTreeSize=8192)
    out = []
    for mode in (K.COLLIDE_GRID, K.COLLIDE_BRUTE):
        plan = _plan(cfg, obs, collision_mode=mode, record_candidates=True)
        out.append(_propagate(plan, parents, 32, 31337, 8192))
    for a, b in zip(out[0], out[1]):
        assert (np.ascontiguousarray(a).view(np.uint8) == np.ascontiguousarray(b).view(np.uint8)).all()
    assert 0.01 < out[0][1].mean() < 0.99


@pytest.mark.parametrize("K_obs,limit,backend", [(10000, 0, 4), (2500, 16 * 1024, 4), (2500, 0, 2)])
def test_c3_streamed_tiles_equal_grid(K_obs, limit, backend):
    """Exhaustive back end with the obstacles streamed through two TMA-fed shared-memory tiles (config 3: the set
    exceeds the staging budget; 2 500 obstacles = 3 tiles, the last one padded) builds the same tree, bit for bit, as
    the grid-culled back end and as the single-piece exhaustive one."""
    obs = w.c3_obstacles(K_obs)
    cfg = dict(w.C3, maxTreeSize=6000)
    a = _plan(cfg, obs, seed=9)
    ra = a.plan(w.C2_INIT, w.C2_GOAL)
    b = _plan(cfg, obs, seed=9, collision_mode=K.COLLIDE_BRUTE, stage_limit_bytes=limit)
    assert b.config()["collide_backend"] == backend
    rb = b.plan(w.C2_INIT, w.C2_GOAL)
    for k in ("stop", "iterations", "tree_size", "expansions", "cost_to_goal", "goal_index"):
        assert ra[k] == rb[k], (k, ra[k], rb[k])
    assert _tree_checksum(a, ra["tree_size"]) == _tree_checksum(b, rb["tree_size"])
    for k in (K.ARR_R1, K.ARR_R1VALID, K.ARR_R1INVALID, K.ARR_R2, K.ARR_R2VALID, K.ARR_R2INVALID, K.ARR_R2AVAIL):
        assert (a.export(k) == b.export(k)).all()
    rb2 = b.plan(w.C2_INIT, w.C2_GOAL)                          # the tile stream restarts cleanly on a re-plan
    assert rb2["tree_size"] == rb["tree_size"] and _tree_checksum(b, rb2["tree_size"]) == _tree_checksum(a, ra["tree_size"])


def test_work_counters_of_the_recording_kernels():
    """kgmt_work_counters: Euler steps and overlap tests executed by a recorded plan — between one step per candidate and
    numDisc, and far fewer overlap tests than the exhaustive K per step; the exhaustive back end tests more pairs for the
    same steps and the same tree."""
    obs = w.c2_obstacles(1000)
    cfg = dict(w.C2, maxTreeSize=20000, n=8)
    a = _plan(cfg, obs, record_candidates=True, seed=8)
    b = _plan(cfg, obs, record_candidates=True, seed=8, collision_mode=K.COLLIDE_BRUTE)
    ra, rb = a.plan(w.C2_INIT, w.C2_GOAL), b.plan(w.C2_INIT, w.C2_GOAL)
    assert ra["tree_size"] == rb["tree_size"] and ra["expansions"] == rb["expansions"]
    wa, wb = a.work_counters(), b.work_counters()
    assert wa["expansions"] == ra["expansions"] and wa["steps"] == wb["steps"]
    assert wa["expansions"] <= wa["steps"] <= 10 * wa["expansions"]
    assert 0 < wa["pairs"] < wb["pairs"] <= 1000 * wb["steps"] + 4 * wb["steps"]


def test_csv_dump_matches_reference_format(tmp_path):
    """The 13 files of KGMT.cu:299-311 in the format of helper.cuh:53-72 ("%.10f", one row per node)."""
    plan = _plan(dict(w.C1, maxTreeSize=2000), w.C1_OBSTACLES, record_candidates=True, seed=3)
    r = plan.plan(w.C1_INIT, w.C1_GOAL)
    plan.dump_csv(str(tmp_path))
    names = ["samples.csv", "unexploredSamples.csv", "parentRelations.csv", "uParentIdx.csv", "G.csv", "R2Avail.csv",
             "R1Avail.csv", "R1Valid.csv", "R2Valid.csv", "R1Invalid.csv", "R2Invalid.csv", "R1Score.csv", "R1.csv"]
    for nm in names:
        assert (tmp_path / nm).exists(), nm
    rows = (tmp_path / "samples.csv").read_text().splitlines()
    assert len(rows) == 2000 and rows[0] == "5.0000000000,5.0000000000,0.0000000000,0.0000000000,0.0000000000,0.0000000000,0.0000000000"
    par = np.loadtxt(tmp_path / "parentRelations.csv", dtype=np.int32)
    assert par[0] == -1 and (par[r["tree_size"]:] == -1).all()
    s = np.loadtxt(tmp_path / "samples.csv", delimiter=",", dtype=np.float64)
    np.testing.assert_allclose(s, plan.export(K.ARR_SAMPLES), atol=6e-11 + 0, rtol=1e-7)


# --------------------------------------------------------------------------------- batched planning (config 4)
@pytest.mark.parametrize("cluster", [1, 4, 8])
def test_batch_equals_one_plan_per_query(cluster):
    """kgmt_plan_batch (thread-block clusters, one query each) gives, for every query, exactly what kgmt_plan gives
    for the same (init, goal, seed): stop reason, iterations, tree size, cost, expansions, goal index, solution path."""
    Q = 48
    inits, goals = w.random_queries(Q, w.C1_OBSTACLES)
    seeds = np.arange(100, 100 + Q)
    cfg = dict(w.C1, maxTreeSize=12000)
    one = _plan(cfg, w.C1_OBSTACLES)
    ref = []
    for q in range(Q):
        one.set_seed(int(seeds[q]))
        r = one.plan(inits[q], goals[q])
        ref.append((r, one.extract_path() if r["stop"] == 1 else np.zeros((0, 7), np.float32)))
    batch = _plan(cfg, w.C1_OBSTACLES)
    res, ms, paths, ws = batch.plan_batch(inits, goals, seeds, cluster_size=cluster, max_path=128)
    assert ws >= 1 and ms > 0
    for q in range(Q):
        r, path = ref[q]
        for k in ("stop", "iterations", "tree_size", "expansions", "goal_index"):
            assert res[q][k] == r[k], (q, k, res[q], r)
        assert np.float32(res[q]["cost_to_goal"]) == np.float32(r["cost_to_goal"])
        assert paths[q].shape == path.shape and (bits(paths[q]) == bits(path)).all()
    assert sum(r["stop"] == 1 for r, _ in ref) > Q // 2
