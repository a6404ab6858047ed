"""GPU parity against the reference's OWN CUDA kernels beyond propagateG-with-trivial-maps (VERDICT r01 "next" item 1).

oracle/_ref/libref_gpu.so = /root/reference/src/planners/KGMT.cu + statePropagator.cu + collisionCheck.cu compiled
unmodified for sm_100a with cuRAND Philox states (oracle/Makefile).  Here:

  (a) propagateGV2 (KGMT.cu:415-482, the "tree nearly full" policy) vs the product's mode-2 iteration;
  (b) propagateG at numDisc = 40 on the config-3 map, and with the fine N=16 / n=32 region grid of config 2;
  (c) the accept rule (KGMT.cu:394-407) with NON-trivial R1 scores and a pre-filled R2Avail, through the stage entry
      point kgmt_stage_update_maps, then kgmt_stage_insert against the reference's scan + findInd + updateG;
  (d) the reference-compatible DEVICE functions of cudasbmp_b200/include (propagateAndCheck, isMotionValid,
      isBroadPhaseValid) called from a __global__ test kernel (tests/native/facade_probe.cu).
All comparisons are bit-exact.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from cudasbmp_b200 import kgmt as K          # noqa: E402
from cudasbmp_b200 import workloads as w     # noqa: E402
from tests.parity import bits                # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAPS = ("R1", "R2", "R1Valid", "R2Valid", "R1Invalid", "R2Invalid", "R1Avail", "R2Avail")
MAP_IDS = {"R1": K.ARR_R1, "R2": K.ARR_R2, "R1Valid": K.ARR_R1VALID, "R2Valid": K.ARR_R2VALID,
           "R1Invalid": K.ARR_R1INVALID, "R2Invalid": K.ARR_R2INVALID, "R1Avail": K.ARR_R1AVAIL, "R2Avail": K.ARR_R2AVAIL}


def _ref(oracle):
    R = oracle.ref_gpu()
    if R is None:
        pytest.skip("oracle/_ref/libref_gpu.so not built (needs /root/reference at build time)")
    return R


def _zero_maps(c1, c2):
    return {k: np.zeros(c1 if k.startswith("R1") else c2, dtype=np.int32) for k in MAPS}


def _counters_match(maps, r1, r2, valid, c1, c2):
    inside = r1 >= 0
    ok = inside & (r2 >= 0)
    assert (np.bincount(r1[inside], minlength=c1) == maps["R1"]).all()
    assert (np.bincount(r1[inside & (valid == 1)], minlength=c1) == maps["R1Valid"]).all()
    assert (np.bincount(r1[inside & (valid == 0)], minlength=c1) == maps["R1Invalid"]).all()
    assert (np.bincount(r2[ok], minlength=c2) == maps["R2"]).all()
    assert (np.bincount(r2[ok & (valid == 1)], minlength=c2) == maps["R2Valid"]).all()
    assert (np.bincount(r2[ok & (valid == 0)], minlength=c2) == maps["R2Invalid"]).all()


# --------------------------------------------------------------------------------------------- (a) propagateGV2
@pytest.mark.parametrize("mode", [K.COLLIDE_GRID, K.COLLIDE_BRUTE])
@pytest.mark.parametrize("children", [1, 5, 17])
def test_mode2_iteration_bit_exact_against_reference_propagateGV2(oracle, mode, children):
    """A frontier of P nodes with room for floor(remaining / P) = `children` children each: the product takes the
    reference's propagateGV2 branch (KGMT.cu:153-158).  One whole iteration of the product (cooperative kernel) against
    ONE launch of the reference's propagateGV2 from the same tree, maps, scores and Philox key: candidate rows, parent
    indices, flags, every counter array."""
    _ref(oracle)
    obstacles = w.c2_obstacles(1000)
    N, n, P, seed = 16, 8, 700, 77
    c1, c2 = N * N, N * N * n * n
    maxTree = P + P * children + P // 2                      # floor((maxTree - P) / P) == children, 32 P does not fit
    nodes = w.random_parents(P, obstacles, seed=31)
    plan = K.KGMT(**dict(w.C1, maxTreeSize=maxTree), seed=seed, collision_mode=mode, record_candidates=True)
    plan.set_obstacles(obstacles)
    plan.seed_frontier(nodes, w.C2_GOAL)
    before = {k: plan.export(i).copy() for k, i in MAP_IDS.items()}
    score = plan.export(K.ARR_R1SCORE).copy()
    st = plan.iterate()
    assert (st["mode"], st["children"], st["candidates"]) == (2, children, P * children), st
    M = st["candidates"]
    cand = plan.export(K.ARR_UNEXPLORED)[:M]
    valid = plan.export(K.ARR_U_VALID)[:M]
    accept = plan.export(K.ARR_U_ACCEPT)[:M]
    r1, r2 = plan.export(K.ARR_U_R1)[:M], plan.export(K.ARR_U_R2)[:M]
    u3 = plan.export(K.ARR_U_U3)[:M]
    after = {k: plan.export(i).copy() for k, i in MAP_IDS.items()}

    maps = {k: before[k].copy() for k in MAPS}
    unx, upar, gnew, _ = oracle.ref_gpu_expand(2, children, nodes, np.arange(P, dtype=np.int32), maps, score, N, n,
                                               plan.R1Size_, plan.R2Size_, 10, 1.0, obstacles, 20.0, 20.0,
                                               (seed + st["iteration"]) & 0xFFFFFFFF)
    assert (bits(unx) == bits(cand)).all(), "candidate rows differ from propagateGV2"
    assert (upar == np.arange(M) // children).all() and (plan.export(K.ARR_U_PARENT)[:M] == upar).all()
    # counters: integer adds commute -> the reference's arrays are deterministic on in-range cells
    for k in ("R1", "R2", "R1Valid", "R2Valid", "R1Invalid", "R2Invalid", "R1Avail", "R2Avail"):
        assert (maps[k] == after[k]).all(), k
    # accept (KGMT.cu:461-464): exact where the R2 cell was available at launch; where it was not, the reference reads
    # R2Avail while sibling threads set it — canonical = the iteration-start snapshot
    inside = (r1 >= 0) & (r2 >= 0) & (valid == 1)
    was = np.zeros(M, dtype=bool)
    was[inside] = before["R2Avail"][r2[inside]] != 0
    sc = np.ones(M, dtype=np.float32)
    sc[inside] = score[r1[inside]]
    canonical = inside & ((u3 <= sc) | ~was)
    assert (accept[inside] == canonical[inside]).all()
    assert (gnew[inside & was] == canonical[inside & was]).all(), "accept differs where the reference is deterministic"
    lower = inside & (u3 <= sc)
    assert (gnew[lower] == 1).all() and (gnew[inside & ~canonical] == 0).all()
    assert (gnew[~inside & (r1 >= 0)] == 0).all()
    assert 0.02 < valid.mean() < 0.98


# ------------------------------------------------------------------- (b) numDisc = 40 (config 3), fine region grid
@pytest.mark.parametrize("mode", [K.COLLIDE_GRID, K.COLLIDE_BRUTE])
def test_propagateG_bit_exact_config3_numdisc40(oracle, mode):
    """propagateG on the config-3 map: 10 000 obstacles, 40 Euler steps per edge (brute = the TMA tile stream in the
    planner; stage_propagate uses the exhaustive walk of the same arithmetic)."""
    _ref(oracle)
    obstacles = w.c3_obstacles(10000)
    N, n, P, children, key = 16, 8, 1024, 32, 1234
    M = P * children
    parents = w.random_parents(P, obstacles, seed=41)
    plan = K.KGMT(**dict(w.C1, numDisc=40, maxTreeSize=M), collision_mode=mode, record_candidates=True)
    plan.set_obstacles(obstacles)
    plan.stage_propagate(parents, children, key, 0)
    x1 = plan.export(K.ARR_UNEXPLORED)[:M]
    valid = plan.export(K.ARR_U_VALID)[:M]
    r1, r2 = plan.export(K.ARR_U_R1)[:M], plan.export(K.ARR_U_R2)[:M]
    c1, c2 = N * N, N * N * n * n
    maps = _zero_maps(c1, c2)
    unx, upar, gnew, _ = oracle.ref_gpu_expand(1, children, parents, np.arange(P, dtype=np.int32), maps,
                                               np.ones(c1, dtype=np.float32), N, n, plan.R1Size_, plan.R2Size_, 40, 1.0,
                                               obstacles, 20.0, 20.0, key)
    assert (bits(unx) == bits(x1)).all()
    inside = r1 >= 0
    assert (gnew[inside] == valid[inside]).all()
    _counters_match(maps, r1, r2, valid, c1, c2)
    assert 0.05 < valid.mean() < 0.95


def test_config3_planner_iteration_streamed_tiles_against_reference(oracle):
    """The same numDisc = 40 / 10 000-obstacle comparison through the PLANNER's kernels: one iteration from a seeded
    frontier with the exhaustive back end (TMA tile stream, COL_BRUTE_STREAM) and with the culled back end."""
    _ref(oracle)
    obstacles = w.c3_obstacles(10000)
    N, n, P, seed = 16, 8, 256, 5
    c1, c2 = N * N, N * N * n * n
    nodes = w.random_parents(P, obstacles, seed=43)
    for mode in (K.COLLIDE_BRUTE, K.COLLIDE_GRID):
        plan = K.KGMT(**dict(w.C1, numDisc=40, maxTreeSize=P + 32 * P + 64), seed=seed, collision_mode=mode,
                      record_candidates=True)
        plan.set_obstacles(obstacles)
        if mode == K.COLLIDE_BRUTE:
            assert plan.config()["collide_backend"] == 4, plan.config()       # exhaustive, TMA-streamed tiles
        plan.seed_frontier(nodes, w.C2_GOAL)
        before = {k: plan.export(i).copy() for k, i in MAP_IDS.items()}
        score = plan.export(K.ARR_R1SCORE).copy()
        st = plan.iterate()
        M = st["candidates"]
        assert st["mode"] == 1 and M == 32 * P
        maps = {k: before[k].copy() for k in MAPS}
        unx, upar, gnew, _ = oracle.ref_gpu_expand(1, 32, nodes, np.arange(P, dtype=np.int32), maps, score, N, n,
                                                   plan.R1Size_, plan.R2Size_, 40, 1.0, obstacles, 20.0, 20.0,
                                                   (seed + st["iteration"]) & 0xFFFFFFFF)
        assert (bits(unx) == bits(plan.export(K.ARR_UNEXPLORED)[:M])).all(), mode
        for k in ("R1", "R2", "R1Valid", "R2Valid", "R1Invalid", "R2Invalid"):
            assert (maps[k] == plan.export(MAP_IDS[k])).all(), (mode, k)


def test_propagateG_bit_exact_fine_region_grid(oracle):
    """Config 2's region grid (N = 16, n = 32: 262 144 R2 cells) against the reference kernel: region indices enter
    only through the counter arrays, so they are compared through them."""
    _ref(oracle)
    obstacles = w.c2_obstacles(1000)
    N, n, P, children, key = 16, 32, 2048, 32, 60606
    M = P * children
    parents = w.random_parents(P, obstacles, seed=47)
    plan = K.KGMT(**dict(w.C2, maxTreeSize=M), record_candidates=True)
    plan.set_obstacles(obstacles)
    plan.stage_propagate(parents, children, key, 0)
    x1 = plan.export(K.ARR_UNEXPLORED)[:M]
    valid = plan.export(K.ARR_U_VALID)[:M]
    r1, r2 = plan.export(K.ARR_U_R1)[:M], plan.export(K.ARR_U_R2)[:M]
    c1, c2 = N * N, N * N * n * n
    maps = _zero_maps(c1, c2)
    unx, _, gnew, _ = oracle.ref_gpu_expand(1, children, parents, np.arange(P, dtype=np.int32), maps,
                                            np.ones(c1, dtype=np.float32), N, n, plan.R1Size_, plan.R2Size_, 10, 1.0,
                                            obstacles, 20.0, 20.0, key)
    assert (bits(unx) == bits(x1)).all()
    inside = r1 >= 0
    assert (gnew[inside] == valid[inside]).all()
    _counters_match(maps, r1, r2, valid, c1, c2)
    assert len(np.unique(r2[r2 >= 0])) > 20000


# ------------------------------------------------------------- (c) accept rule with non-trivial scores and R2Avail
def test_accept_rule_and_insertion_against_reference_with_nontrivial_maps(oracle):
    """kgmt_stage_update_maps + kgmt_stage_insert on caller-supplied candidates against the reference's propagateG
    (maps, GNew) and scan + findInd + updateG, with R1 scores spread over (0, 1] and half of the R2 cells marked
    available before the launch (KGMT.cu:394-407)."""
    R = _ref(oracle)
    obstacles = w.c2_obstacles(1000)
    N, n, P, children, key, seed = 16, 8, 900, 32, 4242, 11
    c1, c2 = N * N, N * N * n * n
    M = P * children
    cap = P + M + 64
    rng = np.random.default_rng(3)
    nodes = w.random_parents(P, obstacles, seed=53)
    plan = K.KGMT(**dict(w.C1, maxTreeSize=cap), seed=seed, record_candidates=True)
    plan.set_obstacles(obstacles)
    goal = np.array([-100, -100, 0, 0, 0, 0, 0], dtype=np.float32)      # unreachable: this test is not about the goal
    plan.seed_frontier(nodes, goal)
    # candidates of the pending iteration, computed by the product's stages 2-4 (bit-exact with the reference: tests above)
    key0 = (seed + 1) & 0xFFFFFFFF
    plan.stage_propagate(nodes, children, key0, 0)
    cand = plan.export(K.ARR_UNEXPLORED)[:M].copy()
    valid = plan.export(K.ARR_U_VALID)[:M].copy()
    u3 = plan.export(K.ARR_U_U3)[:M].copy()
    parent = (np.arange(M) // children).astype(np.int32)
    # non-trivial maps: random counters, scores in (0, 1], half of the R2 cells already available
    maps0 = {"R1": rng.integers(0, 50, c1), "R1Valid": rng.integers(0, 30, c1), "R1Invalid": rng.integers(0, 20, c1),
             "R1Avail": rng.integers(0, 2, c1), "R2": rng.integers(0, 9, c2), "R2Valid": rng.integers(0, 5, c2),
             "R2Invalid": rng.integers(0, 4, c2), "R2Avail": rng.integers(0, 2, c2)}
    maps0 = {k: v.astype(np.int32) for k, v in maps0.items()}
    score = rng.uniform(0.02, 1.0, c1).astype(np.float32)
    for k, i in MAP_IDS.items():
        plan.import_(i, maps0[k])
    plan.import_(K.ARR_R1SCORE, score)
    tree0 = plan.export(K.ARR_SAMPLES).copy()
    par0, cost0 = plan.export(K.ARR_PARENT).copy(), plan.export(K.ARR_COSTS).copy()

    plan.stage_update_maps(cand, valid, u3, parent)
    accept = plan.export(K.ARR_U_ACCEPT)[:M].copy()
    r1, r2 = plan.export(K.ARR_U_R1)[:M].copy(), plan.export(K.ARR_U_R2)[:M].copy()
    after = {k: plan.export(i).copy() for k, i in MAP_IDS.items()}

    maps = {k: maps0[k].copy() for k in MAPS}
    unx, upar, gnew, _ = oracle.ref_gpu_expand(1, children, nodes, np.arange(P, dtype=np.int32), maps, score, N, n,
                                               plan.R1Size_, plan.R2Size_, 10, 1.0, obstacles, 20.0, 20.0, key0)
    assert (bits(unx) == bits(cand)).all()
    for k in MAPS:
        assert (maps[k] == after[k]).all(), k
    inside = (r1 >= 0) & (r2 >= 0) & (valid == 1)
    was = np.zeros(M, dtype=bool)
    was[inside] = maps0["R2Avail"][r2[inside]] != 0
    sc = np.ones(M, dtype=np.float32)
    sc[inside] = score[r1[inside]]
    canonical = inside & ((u3 <= sc) | ~was)
    assert (accept.astype(bool) == canonical).all(), "product accept != canonical snapshot rule"
    assert (gnew[inside & was].astype(bool) == canonical[inside & was]).all(), "accept differs from the reference kernel"
    assert (gnew[inside & (u3 <= sc)] == 1).all() and (gnew[~canonical & (r1 >= 0)] == 0).all()
    # both branches of the rule really were exercised
    assert (inside & was & (u3 <= sc)).sum() > 500 and (inside & was & (u3 > sc)).sum() > 500 and (inside & ~was).sum() > 500

    # stage 5b: the product's ordered insertion of ITS accept mask vs the reference's scan + findInd + updateG on the same mask
    st = plan.stage_insert()
    assert st["accepted"] == int(accept.sum()) and st["tree_size"] == P + st["accepted"]
    g = np.zeros(cap, dtype=np.uint8); g[:M] = accept
    unx_cap = np.zeros((cap, 7), dtype=np.float32); unx_cap[:M] = cand
    upar_cap = np.zeros(cap, dtype=np.int32); upar_cap[:M] = parent
    G = np.zeros(cap, dtype=np.uint8)
    ctg = np.zeros(1, dtype=np.float32)
    k = R.ref_gpu_insert(cap, g.ctypes.data_as(oracle.u8p), unx_cap.ctypes.data_as(oracle.f32p),
                         upar_cap.ctypes.data_as(oracle.i32p), P, tree0.ctypes.data_as(oracle.f32p),
                         par0.ctypes.data_as(oracle.i32p), cost0.ctypes.data_as(oracle.f32p), G.ctypes.data_as(oracle.u8p),
                         goal.ctypes.data_as(oracle.f32p), 0.5, ctg.ctypes.data_as(oracle.f32p))
    assert k == st["accepted"]
    T1 = st["tree_size"]
    assert (bits(tree0[:T1]) == bits(plan.export(K.ARR_SAMPLES)[:T1])).all()
    assert (par0[:T1] == plan.export(K.ARR_PARENT)[:T1]).all()
    assert (bits(cost0[:T1]) == bits(plan.export(K.ARR_COSTS)[:T1])).all()
    # and the planner carries on from there with its own loop
    st2 = plan.iterate()
    assert st2["frontier"] == st["accepted"] and st2["iteration"] == 2


# ----------------------------------------------------------------------- (d) the facade's device functions
def _probe():
    so = os.path.join(ROOT, "tests", "native", "libfacade_probe.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "native")])
    L = C.CDLL(so)
    f32p, u8p = C.POINTER(C.c_float), C.POINTER(C.c_ubyte)
    L.facade_probe_propagate.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float, f32p, C.c_int, C.c_float, C.c_float,
                                         C.c_uint, f32p, u8p]
    L.facade_probe_motion.argtypes = [f32p, C.c_long, f32p, C.c_int, u8p, u8p]
    return L, f32p, u8p


@pytest.mark.parametrize("case", ["c1", "c2"])
def test_facade_device_propagateAndCheck_against_reference_kernel(oracle, case):
    """cudasbmp_b200/include/statePropagator/statePropagator.cuh::propagateAndCheck called from a test kernel with a
    cuRAND Philox state per slot, against the reference's propagateG: x1[7] rows and validity bit for bit."""
    _ref(oracle)
    L, f32p, u8p = _probe()
    obstacles = np.ascontiguousarray(w.C1_OBSTACLES if case == "c1" else w.c2_obstacles(1000), dtype=np.float32)
    N, n, P, children, key = 16, 8, 1024, 32, 9119
    M = P * children
    parents = np.ascontiguousarray(w.random_parents(P, obstacles, seed=59))
    rng = np.random.default_rng(8)
    parents[: P // 4, 2] = (rng.uniform(-1, 1, P // 4) * 1e5).astype(np.float32)      # Payne-Hanek headings
    x1 = np.zeros((M, 7), dtype=np.float32)
    valid = np.zeros(M, dtype=np.uint8)
    rc = L.facade_probe_propagate(parents.ctypes.data_as(f32p), P, children, 10, 1.0, obstacles.ctypes.data_as(f32p),
                                  len(obstacles), 20.0, 20.0, key, x1.ctypes.data_as(f32p), valid.ctypes.data_as(u8p))
    assert rc == 0
    c1, c2 = N * N, N * N * n * n
    maps = _zero_maps(c1, c2)
    unx, _, gnew, _ = oracle.ref_gpu_expand(1, children, parents, np.arange(P, dtype=np.int32), maps,
                                            np.ones(c1, dtype=np.float32), N, n, 20.0 / N, 20.0 / (N * n), 10, 1.0,
                                            obstacles, 20.0, 20.0, key)
    assert (bits(unx) == bits(x1)).all(), "facade propagateAndCheck differs from the reference kernel"
    r1 = np.array([oracle.getR1(float(a), float(b), 20.0 / N, N) for a, b in x1[:, :2]])
    inside = r1 >= 0
    assert (gnew[inside] == valid[inside]).all()
    # and the planner's own stages 2-4 give the same rows
    plan = K.KGMT(**dict(w.C1, maxTreeSize=M), record_candidates=True)
    plan.set_obstacles(obstacles)
    plan.stage_propagate(parents, children, key, 0)
    assert (bits(plan.export(K.ARR_UNEXPLORED)[:M]) == bits(x1)).all()
    assert (plan.export(K.ARR_U_VALID)[:M] == valid).all()


def test_facade_device_collision_functions(oracle):
    """isMotionValid / isBroadPhaseValid of cudasbmp_b200/include/collisionCheck/collisionCheck.cuh on random and
    edge-touching boxes against the oracle's restatement of collisionCheck.cu:6-28 (itself pinned to the reference's
    host build) — strict overlap: touching boxes do not collide."""
    L, f32p, u8p = _probe()
    obstacles = np.ascontiguousarray(w.c2_obstacles(1000), dtype=np.float32)
    rng = np.random.default_rng(12)
    B = 20000
    lo = rng.uniform(0, 19.5, (B, 2)).astype(np.float32)
    ext = rng.uniform(0, 0.4, (B, 2)).astype(np.float32)
    boxes = np.concatenate([lo, lo + ext], axis=1).astype(np.float32)
    # boxes that exactly touch an obstacle face (max == obstacle min, and min == obstacle max)
    k = rng.integers(0, len(obstacles), 2000)
    boxes[:1000, 2] = obstacles[k[:1000], 0]; boxes[:1000, 0] = boxes[:1000, 2] - 0.1
    boxes[:1000, 1] = obstacles[k[:1000], 1]; boxes[:1000, 3] = obstacles[k[:1000], 3]
    boxes[1000:2000, 0] = obstacles[k[1000:], 2]; boxes[1000:2000, 2] = boxes[1000:2000, 0] + 0.1
    boxes[1000:2000, 1] = obstacles[k[1000:], 1]; boxes[1000:2000, 3] = obstacles[k[1000:], 3]
    boxes = np.ascontiguousarray(boxes)
    mv = np.zeros(B, dtype=np.uint8); fb = np.zeros(B, dtype=np.uint8)
    rc = L.facade_probe_motion(boxes.ctypes.data_as(f32p), B, obstacles.ctypes.data_as(f32p), len(obstacles),
                               mv.ctypes.data_as(u8p), fb.ctypes.data_as(u8p))
    assert rc == 0
    OL = oracle.lib()
    OL.orc_motion_valid.restype = C.c_int
    OL.orc_motion_valid.argtypes = [f32p, f32p, f32p, C.c_int]
    ob = obstacles.ctypes.data_as(f32p)
    for i in range(B):
        bmin = (C.c_float * 2)(boxes[i, 0], boxes[i, 1]); bmax = (C.c_float * 2)(boxes[i, 2], boxes[i, 3])
        assert OL.orc_motion_valid(bmin, bmax, ob, len(obstacles)) == mv[i], i
        assert OL.orc_motion_valid(bmin, bmax, ob, 1) == fb[i], i
    assert 0.1 < mv.mean() < 0.9


# ------------------------------------------------------------------------- runtime car model
def test_nondefault_control_ranges_against_oracle(oracle):
    """kgmt_params control ranges other than the literals of statePropagator.cu:17-19: controls bit-exact against the
    oracle's general form (which equals the literal form at the defaults: tests/test_oracle_pin.py), end states within
    the FP32 tolerance, and the default ranges give the reference stream bit for bit."""
    from tests.parity import MARGIN, state_tolerance
    obstacles = w.c2_obstacles(1000)
    P, children, key = 512, 32, 777
    M = P * children
    parents = w.random_parents(P, obstacles, seed=61)
    pof = (np.arange(M) // children).astype(np.int32)
    ranges = dict(accel_min=-2.0, accel_max=3.5, steer_min=-0.6, steer_max=0.45, duration_min=0.1, duration_max=0.45)
    plan = K.KGMT(**dict(w.C1, maxTreeSize=M, agentLength=2.5), record_candidates=True, car=ranges)
    plan.set_obstacles(obstacles)
    plan.stage_propagate(parents, children, key, 0)
    x1 = plan.export(K.ARR_UNEXPLORED)[:M]
    valid = plan.export(K.ARR_U_VALID)[:M]
    r6 = [ranges[k] for k in ("accel_min", "accel_max", "steer_min", "steer_max", "duration_min", "duration_max")]
    oracle.set_car_ranges(r6)
    try:
        xo, vo, _, margin = oracle.propagate_batch(parents, pof, key, 0, 10, 2.5, obstacles, 20.0, 20.0, oracle.MATH_FMA)
    finally:
        oracle.set_car_ranges(None)
    assert (bits(x1[:, 4:7]) == bits(xo[:, 4:7])).all(), "controls differ from the oracle's general form"
    assert x1[:, 4].min() > -2.0 and x1[:, 4].max() <= 3.5 and x1[:, 5].min() > -0.6 and x1[:, 5].max() <= 0.45
    assert x1[:, 6].min() > 0.1 and x1[:, 6].max() <= 0.45 + 1e-6
    err = np.abs(x1[:, :4].astype(np.float64) - xo[:, :4]) / np.maximum(1.0, np.abs(xo[:, :4]))
    off = (valid != vo) | (err.max(axis=1) > state_tolerance(xo, parents[pof, 2], 10))
    assert (margin[off] <= MARGIN).all() and off.mean() <= 0.01
