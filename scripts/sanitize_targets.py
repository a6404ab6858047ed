"""Small, complete runs of every kernel family of the library, for compute-sanitizer (scripts/gpu_sanitize.sh) and for the
bounds-checked build (scripts/bounds_check.sh: KGMT_LIB=cudasbmp_b200/libkgmt_b200_check.so):
    python scripts/sanitize_targets.py c1|c2|c3s|batch|peer2|fused1|stage
Each target ends with a result check, so a sanitizer run that also prints 'target ok' exercised the real path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cudasbmp_b200 import kgmt as K, workloads as w

what = sys.argv[1]
if what == "c1":                       # reference demo, whole plan, culled back end + recording kernels + export + path
    p = K.KGMT(**w.C1, seed=3, record_candidates=True); p.set_obstacles(w.C1_OBSTACLES)
    r = p.plan(w.C1_INIT, w.C1_GOAL)
    assert r["stop"] in (1, 2, 3, 4) and r["tree_size"] > 1
    p.export(K.ARR_SAMPLES); p.export(K.ARR_U_ACCEPT)
    if r["stop"] == 1:
        assert len(p.extract_path()) > 1
    q = K.KGMT(**w.C1, seed=3); q.set_obstacles(w.C1_OBSTACLES)
    assert q.plan(w.C1_INIT, w.C1_GOAL)["tree_size"] == r["tree_size"]
elif what == "c2":                     # config-2 map (cull grid in shared memory, 3 CTAs per SM), tree capped for the tool's speed
    cfg = dict(w.C2, maxTreeSize=120000)
    p = K.KGMT(**cfg, seed=1); p.set_obstacles(w.c2_obstacles(1000))
    r = p.plan(w.C2_INIT, w.C2_GOAL)
    assert r["tree_size"] > 1000, r
    r2 = p.plan(w.C2_INIT, w.C2_GOAL)
    assert r2["tree_size"] == r["tree_size"]
elif what == "c3s":                    # config-3 obstacles, exhaustive back end: TMA tile stream (mbarrier, bulk copies)
    cfg = dict(w.C3, maxTreeSize=3000, numDisc=40)
    p = K.KGMT(**cfg, seed=1, collision_mode=K.COLLIDE_BRUTE); p.set_obstacles(w.c3_obstacles(10000))
    assert p.config()["collide_backend"] == 4
    r = p.plan(w.C2_INIT, w.C2_GOAL)
    g = K.KGMT(**cfg, seed=1); g.set_obstacles(w.c3_obstacles(10000))
    assert g.plan(w.C2_INIT, w.C2_GOAL)["tree_size"] == r["tree_size"]
elif what == "batch":                  # config 4: clusters, one query each
    Q = 48
    inits, goals = w.random_queries(Q, w.C1_OBSTACLES)
    p = K.KGMT(**dict(w.C1, maxTreeSize=6000), seed=1); p.set_obstacles(w.C1_OBSTACLES)
    for cs in (2, 4):
        res, ms, paths, ws = p.plan_batch(inits, goals, np.arange(Q), cluster_size=cs, max_path=64)
        assert len(res) == Q and all(r["tree_size"] >= 1 for r in res)
elif what == "peer2":                  # two emulated ranks, multi-launch peer exchange (mailboxes, peer stores)
    cfg = dict(w.C1, maxTreeSize=8000)
    ranks = []
    for g in range(2):
        p = K.KGMT(**cfg, seed=13); p.set_obstacles(w.C1_OBSTACLES); p.begin(w.C1_INIT, w.C1_GOAL); ranks.append(p)
    ref = K.KGMT(**cfg, seed=13); ref.set_obstacles(w.C1_OBSTACLES); ref.begin(w.C1_INIT, w.C1_GOAL)
    for g, p in enumerate(ranks):
        p.peer_attach_local(g, ranks)
    for it in range(6):
        want = ref.iterate()
        for p in ranks:
            p.peer_expand_begin()
        got = [p.peer_expand_end() for p in ranks]
        assert all(x == want for x in got), (it, got, want)
        if want["stop"]:
            break
elif what == "fused1":                 # the fused persistent kernel against itself (world 1) + the communicator calls
    cfg = dict(w.C1, maxTreeSize=8000)
    ref = K.KGMT(**cfg, seed=17); ref.set_obstacles(w.C1_OBSTACLES)
    p = K.KGMT(**cfg, seed=17); p.set_obstacles(w.C1_OBSTACLES)
    p.comm_init(0, 1, K.KGMT.comm_unique_id())
    a, b = ref.plan(w.C1_INIT, w.C1_GOAL), p.plan_sharded(w.C1_INIT, w.C1_GOAL)
    assert (a["tree_size"], a["iterations"], a["cost_to_goal"]) == (b["tree_size"], b["iterations"], b["cost_to_goal"]), (a, b)
    p.begin(w.C1_INIT, w.C1_GOAL)
    for ex in (K.EXCHANGE_FUSED, K.EXCHANGE_PEER_LAUNCHES, K.EXCHANGE_NCCL):
        p.expand_sharded(ex)
    p.comm_destroy()
elif what == "stage":                  # stage entry points: propagate, update_maps, insert, scores
    obs = w.c2_obstacles(1000)
    P, M = 128, 4096
    nodes = w.random_parents(P, obs, seed=5)
    p = K.KGMT(**dict(w.C1, maxTreeSize=P + M + 64), seed=2, record_candidates=True); p.set_obstacles(obs)
    p.seed_frontier(nodes, w.C2_GOAL)
    p.stage_propagate(nodes, 32, 3, 0)
    cand, valid, u3 = p.export(K.ARR_UNEXPLORED)[:M].copy(), p.export(K.ARR_U_VALID)[:M].copy(), p.export(K.ARR_U_U3)[:M].copy()
    p.stage_update_maps(cand, valid, u3, (np.arange(M) // 32).astype(np.int32))
    st = p.stage_insert()
    assert st["accepted"] == int(p.export(K.ARR_U_ACCEPT)[:M].sum())
    p.stage_scores()
else:
    raise SystemExit("unknown target " + what)
if os.environ.get("KGMT_LIB"):         # the bounds-checked twin of the library: the device-side index checks must all have held
    chk = p.debug_checks()
    print("bounds checks:", chk)
    assert chk["failures"] == 0, chk
print("target ok:", what)
