"""CPU: pins the oracle (oracle/kgmt_oracle.c) against vectors produced by the REFERENCE'S OWN code
(tests/golden/make_golden.py ran libref_host.so / libref_gpu.so in the build container) and, when those
builds are present, against the reference live."""
import ctypes as C
import os

import numpy as np
import pytest


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_philox_known_answers(oracle, golden_dir):
    g = _load(golden_dir, "philox.npz")
    for ctr, key, out in zip(g["ctr"], g["key"], g["out"]):
        assert (oracle.philox(ctr, key) == out).all()
    # curand_init(1234, 7, 0) + 4 x curand_uniform
    u = oracle.slot_uniforms(1234, 7)
    np.testing.assert_array_equal(u, np.array([0.33944732, 0.129023373, 0.912783206, 0.116881445], dtype=np.float32))


@pytest.mark.parametrize("name", ["propagate_c1.npz", "propagate_c2.npz", "propagate_c3.npz", "propagate_root.npz"])
def test_propagate_matches_reference_host_build_bit_exact(oracle, golden_dir, name):
    g = _load(golden_dir, name)
    x1, valid, u3, _ = oracle.propagate_batch(g["parents"], g["parent_of"], int(g["key"]), 0, int(g["num_disc"]),
                                              float(g["L"]), g["obstacles"], float(g["W"]), float(g["H"]),
                                              oracle.MATH_HOST)
    assert (x1.view(np.uint32) == g["x1"].view(np.uint32)).all()          # states: bit-exact
    assert (valid == g["valid"]).all()                                     # collision / bounds flags
    assert (u3.view(np.uint32) == g["u3"].view(np.uint32)).all()          # accept uniform


def test_fma_mode_stays_within_tolerance_of_host_mode(oracle, golden_dir):
    """ORC_MATH_FMA (nvcc's contractions) vs the host build: <= 1e-5 relative per step (north_star)."""
    g = _load(golden_dir, "propagate_c1.npz")
    args = (g["parents"], g["parent_of"], int(g["key"]), 0, int(g["num_disc"]), 1.0, g["obstacles"], 20.0, 20.0)
    xh, vh, _, mh = oracle.propagate_batch(*args, oracle.MATH_HOST)
    xf, vf, _, _ = oracle.propagate_batch(*args, oracle.MATH_FMA)
    same = vh == vf
    tol = 1e-5 * int(g["num_disc"])
    err = np.abs(xh[same, :4] - xf[same, :4]) / np.maximum(1.0, np.abs(xh[same, :4]))
    assert err.max() <= tol
    assert (mh[~same] <= tol * 20.0).all()          # flips only next to a boundary


@pytest.mark.parametrize("tag", ["c1", "c2", "n64"])
def test_region_indices_match_reference(oracle, golden_dir, tag):
    g = _load(golden_dir, "regions.npz")
    N, n = (int(v) for v in g[tag + "_Nn"])
    R1 = float(np.float32(20.0) / np.float32(N))
    R2 = float(np.float32(20.0) / np.float32(n * N))
    for x, y, r1, r2 in zip(g[tag + "_x"], g[tag + "_y"], g[tag + "_r1"], g[tag + "_r2"]):
        a = oracle.getR1(x, y, R1, N)
        assert a == r1
        assert oracle.getR2(x, y, a, R1, N, R2, n) == r2


def test_hand_checkable_c1_values(oracle):
    """SURVEY.md App. A.3, derived from the reference source by hand."""
    assert oracle.getR1(5, 5, 1.25, 16) == 68
    assert oracle.getR2(5, 5, 68, 1.25, 16, 0.15625, 8) == 4352
    assert oracle.getR1(2, 18, 1.25, 16) == 225
    assert oracle.expansion_shape(1, 1, 30000) == (1, 32, 32)
    assert oracle.expansion_shape(1000, 1025, 30000) == (2, 28, 28000)
    # first scores: only the root cell is available -> every score is 1.0
    c1, c2 = 256, 256 * 64
    A1 = np.zeros(c1, np.int32); A2 = np.zeros(c2, np.int32); V = np.zeros(c1, np.int32); I = np.zeros(c1, np.int32)
    R = np.zeros(c1, np.int32)
    A1[68] = 1; V[68] = 1; R[68] = 1; A2[4352] = 1
    s, thr = oracle.scores(A1, A2, V, I, R, 16, 8)
    assert (s == 1.0).all()
    assert abs(thr - 1.0 / ((1 + 1 / 64) * 2)) < 1e-7


def test_first_iteration_accepts_every_valid_candidate(oracle):
    from cudasbmp_b200 import workloads as w
    p = oracle.Planner(**{k: v for k, v in dict(width=20.0, height=20.0, N=16, n=8, num_iterations=100, max_tree=30000,
                                               num_disc=10, agent_length=1.0, goal_threshold=0.5).items()}, seed=1)
    p.set_obstacles(w.C1_OBSTACLES)
    p.begin(w.C1_INIT, w.C1_GOAL)
    p.iterate()
    assert p.last_M == 32 and p.last_mode == 1
    valid = p.array(oracle.ARR_U_VALID)[:32]
    assert p.last_accepted == int(valid.sum())
    par = p.array(oracle.ARR_TREE_PARENT)
    assert (par[1:p.tree_size] == 0).all() and par[0] == -1
    costs = p.array(oracle.ARR_COSTS)
    tree = p.array(oracle.ARR_TREE_SAMPLES)
    np.testing.assert_array_equal(costs[1:p.tree_size], tree[1:p.tree_size, 6])       # cost = duration of the edge
    p.close()


def test_oracle_plan_runs_to_a_stop(oracle):
    from cudasbmp_b200 import workloads as w
    p = oracle.Planner(20.0, 20.0, 16, 8, 100, 30000, 10, 1.0, 0.5, seed=3)
    p.set_obstacles(w.C1_OBSTACLES)
    st = p.plan(w.C1_INIT, w.C1_GOAL)
    assert st in (1, 2, 3, 4)
    T = p.tree_size
    par = p.array(oracle.ARR_TREE_PARENT)[:T]
    assert (par[1:] < np.arange(1, T)).all() and (par[1:] >= 0).all()
    if st == 1:
        gi = p.goal_index
        node = p.array(oracle.ARR_TREE_SAMPLES)[gi]
        assert np.hypot(node[0] - 2, node[1] - 18) < 0.5
    p.close()


@pytest.mark.skipif(not os.path.exists("/root/reference"), reason="reference tree only exists in the build container")
def test_live_reference_host_build(oracle):
    """Same check as the golden one, on fresh random inputs, against the reference running live."""
    from cudasbmp_b200 import workloads as w
    assert oracle.ref_host() is not None
    obs = w.c2_obstacles(300)
    par = w.random_parents(50, obs, seed=21)
    pof = np.repeat(np.arange(50, dtype=np.int32), 20)
    x1, v, u3, _ = oracle.propagate_batch(par, pof, 4242, 100, 10, 1.0, obs, 20.0, 20.0, oracle.MATH_HOST)
    _, rx, rv, ru = oracle.ref_host_batch(par, pof, 4242, 100, 10, 1.0, obs, 20.0, 20.0)
    assert (x1.view(np.uint32) == rx.view(np.uint32)).all() and (v == rv).all() and (u3 == ru).all()


def test_general_control_ranges_equal_the_literals_at_the_defaults(oracle):
    """The product draws controls as lo + u * (hi - lo) from kgmt_params (runtime car model); the oracle restates that
    general form next to the reference's literal expressions (statePropagator.cu:17-19, pinned above against the
    reference's own host build).  At the default ranges the two must agree bit for bit, in both math modes, for random
    uniforms and for the extremes of curand_uniform's range (2^-33 and 1)."""
    import math
    d = [-5.0, 5.0, -math.pi, math.pi, float(np.float32(0.05)), float(np.float32(0.05)) + 1.0]
    rng = np.random.default_rng(1)
    words = list(rng.integers(0, 2 ** 32, 30000, dtype=np.uint64)) + [0, 1, 2 ** 32 - 1, 2 ** 31, 2 ** 31 - 1]
    L = oracle.lib()
    for i in range(0, len(words) - 2, 3):
        u = np.array([L.orc_uniform(int(words[i + j])) for j in range(3)], dtype=np.float32)
        for mode in (oracle.MATH_FMA, oracle.MATH_HOST):
            a = np.array(oracle.controls(u, mode), dtype=np.float32)
            b = np.array(oracle.controls(u, mode, d), dtype=np.float32)
            assert a.tobytes() == b.tobytes(), (u, mode, a, b)
