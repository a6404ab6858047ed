"""Stage-by-stage parity of one expansion iteration of the CUDA product against the CPU oracle.

Both sides get IDENTICAL inputs for every stage (the product's own state before the step), so each
comparison is well defined even where the reference itself is racy:

  (a) R1 scores            bit-exact   (orc_scores on the maps before the step)
  (b) expansion shape      exact       (mode, children, M)
  (c) sample+propagate     controls / accept uniform bit-exact; states within TOL_REL per integration step;
      +collide             flags equal except for candidates the oracle itself reports as within MARGIN of a
                           bounds/obstacle comparison flipping (north_star rule)
  (d) region indices       bit-exact   (on the product's own end states)
  (e) map update + accept  bit-exact   (orc_update_maps on the product's r1/r2/valid/u3)
  (f) ordered insertion    bit-exact   (orc_insert on the product's accept mask + candidates): tree rows,
                           parent links, costs, goal cost / index

Used by tests/test_gpu_parity.py and __graft_entry__.smoke().
"""
import numpy as np

from cudasbmp_b200 import kgmt as K

TOL_REL = 1e-5          # per integration step, relative to max(1, |value|)  (BASELINE.json north_star)
MARGIN = 5e-4           # a flag may differ only if some deciding comparison was this close to flipping

MAP_IDS = {"R1": K.ARR_R1, "R2": K.ARR_R2, "R1Valid": K.ARR_R1VALID, "R2Valid": K.ARR_R2VALID,
           "R1Invalid": K.ARR_R1INVALID, "R2Invalid": K.ARR_R2INVALID, "R1Avail": K.ARR_R1AVAIL,
           "R2Avail": K.ARR_R2AVAIL}


def export_maps(plan):
    m = {k: plan.export(i).copy() for k, i in MAP_IDS.items()}
    m["R1Score"] = plan.export(K.ARR_R1SCORE).copy()
    return m


def state_tolerance(x1, theta0, nd):
    """Per-candidate relative tolerance on the propagated state: TOL_REL per integration step (north_star),
    scaled by the heading swept over the edge.  sinf/cosf/tanf of libdevice and glibc differ by <= 1-2 ulp; an
    edge that turns through D radians (steering near +-pi/2: tan up to 1e7) amplifies one ulp of tanf by D, so
    no fixed tolerance can hold for it.  Those edges are ALSO checked bit for bit against the reference's own
    CUDA kernels (test_bit_exact_against_reference_cuda_kernels)."""
    swept = np.abs(x1[:, 2].astype(np.float64) - np.asarray(theta0, dtype=np.float64))
    return TOL_REL * nd * np.maximum(1.0, swept)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def check_iteration(plan, po, obstacles, cfg, goal, seed, report=None):
    """Runs ONE kgmt_expand_iteration on `plan` (record_candidates=True) and checks every stage.
    Returns the iteration stats dict (stop != 0 when the planner has stopped)."""
    N, n = cfg["N"], cfg["n"]
    W, H, L, nd = cfg["width"], cfg["height"], cfg["agentLength"], cfg["numDisc"]
    R1Size, R2Size = plan.R1Size_, plan.R2Size_
    res0 = plan.result()
    if res0["stop"] != 0:
        return dict(stop=res0["stop"], candidates=0)
    T0 = res0["tree_size"]
    before = export_maps(plan)
    tree0 = plan.export(K.ARR_SAMPLES).copy()
    par0 = plan.export(K.ARR_PARENT).copy()
    cost0 = plan.export(K.ARR_COSTS).copy()

    st = plan.iterate()
    M, children, frontier, itr = st["candidates"], st["children"], st["frontier"], st["iteration"]
    after = export_maps(plan)

    # (a) scores used by this iteration
    s, _ = po.scores(before["R1Avail"], before["R2Avail"], before["R1Valid"], before["R1Invalid"], before["R1"], N, n)
    assert (bits(s) == bits(before["R1Score"])).all(), "R1 scores differ from the oracle"

    # (b) shape
    mode, ch, Mo = po.expansion_shape(frontier, T0, cfg["maxTreeSize"])
    assert (mode, ch, Mo) == (st["mode"], children, M), ((mode, ch, Mo), st)

    # (c) candidates
    cand = plan.export(K.ARR_UNEXPLORED)[:M]
    valid = plan.export(K.ARR_U_VALID)[:M]
    r1 = plan.export(K.ARR_U_R1)[:M]
    r2 = plan.export(K.ARR_U_R2)[:M]
    u3 = plan.export(K.ARR_U_U3)[:M]
    accept = plan.export(K.ARR_U_ACCEPT)[:M]
    upar = plan.export(K.ARR_U_PARENT)[:M]
    fstart = T0 - frontier
    assert (upar == fstart + np.arange(M) // children).all(), "candidate -> parent mapping"
    key0 = (seed + itr) & 0xFFFFFFFF
    xo, vo, u3o, margin = po.propagate_batch(tree0, upar, key0, 0, nd, L, obstacles, W, H, po.MATH_FMA)
    assert (bits(cand[:, 4:7]) == bits(xo[:, 4:7])).all(), "sampled controls differ"
    assert (bits(u3) == bits(u3o)).all(), "accept uniforms differ"
    err = np.abs(cand[:, :4].astype(np.float64) - xo[:, :4]) / np.maximum(1.0, np.abs(xo[:, :4]))
    off = (valid != vo) | (err.max(axis=1) > state_tolerance(xo, tree0[upar, 2], nd))
    assert (margin[off] <= MARGIN).all(), \
        "propagated state / flag differs away from any boundary: %d candidates, worst margin %g" % (
            int(off.sum()), float(margin[off].max()))
    assert off.mean() <= 0.01, "too many boundary cases: %g" % off.mean()

    # (d) region indices of the product's own end states
    o1, o2 = po.regions_batch(cand, R1Size, N, R2Size, n)
    assert (o1 == r1).all() and (o2 == r2).all(), "region indices differ"

    # (e) maps + accept
    maps = {k: before[k].copy() for k in MAP_IDS}
    acc = po.update_maps(r1, r2, valid, u3, before["R1Score"], before["R2Avail"], maps)
    assert (acc == accept).all(), "accept mask differs"
    for k in MAP_IDS:
        assert (maps[k] == after[k]).all(), "map %s differs after the iteration" % k

    # (f) insertion
    G = np.zeros(len(par0), dtype=np.uint8)
    tree, par, cost = tree0.copy(), par0.copy(), cost0.copy()
    ctg0, gi0 = res0["cost_to_goal"], res0["goal_index"]
    k, ctg, gi = po.insert(accept, cand, upar, T0, tree, par, cost, G, goal, cfg["goalThreshold"], ctg0, gi0)
    assert k == st["accepted"] and T0 + k == st["tree_size"]
    assert (bits(tree) == bits(plan.export(K.ARR_SAMPLES))).all(), "tree rows differ"
    assert (par == plan.export(K.ARR_PARENT)).all(), "parent links differ"
    assert (bits(cost) == bits(plan.export(K.ARR_COSTS))).all(), "costs differ"
    assert (G == plan.export(K.ARR_G)).all(), "frontier flags differ"
    assert np.float32(ctg) == np.float32(st["cost_to_goal"]) and gi == st["goal_index"], "goal cost / index differ"
    if report is not None:
        report.append(dict(itr=itr, M=M, accepted=k, boundary=int(off.sum()), max_err=float(err[~off].max()) if (~off).any() else 0.0))
    return st
