"""Sharded expansion (config 5) on ONE GPU: `world` ranks are emulated by `world` contexts that each run
kgmt_shard_expand / pack on their slot range; the collectives are done by hand (concatenate the send buffers, add the
delta slabs).  Every rank's tree, parent links, costs and region maps must stay bit-identical to a single-context run of
kgmt_expand_iteration, iteration by iteration (the sharding is invisible in the result)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cudasbmp_b200 import kgmt as K          # noqa: E402
from cudasbmp_b200 import workloads as w     # noqa: E402
from cudasbmp_b200.sharded import ROW_BYTES, ShardedExpander, round_up4     # noqa: E402

MAPS = (K.ARR_R1, K.ARR_R1VALID, K.ARR_R1INVALID, K.ARR_R1AVAIL, K.ARR_R1SCORE, K.ARR_R2, K.ARR_R2VALID, K.ARR_R2INVALID,
        K.ARR_R2AVAIL)


def _same_state(a, b, upto):
    np.testing.assert_array_equal(a.export(K.ARR_SAMPLES)[:upto].view(np.uint32), b.export(K.ARR_SAMPLES)[:upto].view(np.uint32))
    np.testing.assert_array_equal(a.export(K.ARR_PARENT)[:upto], b.export(K.ARR_PARENT)[:upto])
    np.testing.assert_array_equal(a.export(K.ARR_COSTS)[:upto].view(np.uint32), b.export(K.ARR_COSTS)[:upto].view(np.uint32))
    for m in MAPS:
        np.testing.assert_array_equal(a.export(m).view(np.uint32), b.export(m).view(np.uint32), err_msg="map %d" % m)


def _emulated_round(ranks, deltas, dev):
    world = len(ranks)
    torch.cuda.synchronize()
    infos = [p.shard_expand(g, world, deltas[g].data_ptr()) for g, p in enumerate(ranks)]
    counts = [i["accepted_local"] for i in infos]
    cap = round_up4(max(counts))
    sends = [torch.zeros(cap * ROW_BYTES, dtype=torch.uint8, device=dev) for _ in ranks]
    for p, s in zip(ranks, sends):
        p.shard_pack(s.data_ptr(), cap)
    torch.cuda.synchronize()
    recv = torch.cat(sends)                                   # the all-gather
    total = torch.stack(deltas).sum(dim=0).to(torch.int32)    # the all-reduce
    stats = []
    for g, p in enumerate(ranks):
        deltas[g].copy_(total)
        torch.cuda.synchronize()
        stats.append(p.shard_commit(recv.data_ptr(), cap, counts, deltas[g].data_ptr()))
        torch.cuda.synchronize()
        assert int(deltas[g].abs().sum()) == 0                # the slab is handed back zeroed
    assert all(s == stats[0] for s in stats)
    return stats[0], infos, counts


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("cfgname", ["c1", "c2small"])
def test_sharded_iterations_equal_single_gpu(world, cfgname):
    dev = torch.device("cuda", 0)
    if cfgname == "c1":
        cfg, obs, init, goal = w.C1, w.C1_OBSTACLES, w.C1_INIT, w.C1_GOAL
    else:
        cfg, obs, init, goal = dict(w.C2, maxTreeSize=200000), w.c2_obstacles(1000), w.C2_INIT, w.C2_GOAL
    ref = K.KGMT(**cfg, seed=11); ref.set_obstacles(obs); ref.begin(init, goal)
    ranks = []
    for g in range(world):
        p = K.KGMT(**cfg, seed=11); p.set_obstacles(obs); p.begin(init, goal)
        ranks.append(p)
    deltas = [torch.zeros(ranks[0].shard_delta_ints(), dtype=torch.int32, device=dev) for _ in ranks]
    for it in range(120):
        want = ref.iterate()
        got, infos, counts = _emulated_round(ranks, deltas, dev)
        assert got == want, (it, got, want)
        assert sum(counts) == want["accepted"]
        # contiguous chunk ranges that tile the iteration
        assert infos[0]["chunk_lo"] == 0 and all(a["chunk_hi"] == b["chunk_lo"] or b["chunk_hi"] == b["chunk_lo"]
                                                  for a, b in zip(infos, infos[1:]))
        if it < 6 or want["stop"] != 0:
            for p in ranks:
                _same_state(ref, p, want["tree_size"])
        if want["stop"] != 0:
            break
    assert want["stop"] in (1, 2, 3, 4)
    if want["stop"] == 1:
        for p in ranks:
            np.testing.assert_array_equal(ref.extract_path(), p.extract_path())


def test_sharded_driver_single_process_and_mixed_with_cooperative_iterations():
    """ShardedExpander (world 1, planner on torch's stream) == kgmt_plan; shard rounds and cooperative launches mix."""
    cfg, obs = w.C1, w.C1_OBSTACLES
    ref = K.KGMT(**cfg, seed=3); ref.set_obstacles(obs)
    want = ref.plan(w.C1_INIT, w.C1_GOAL)
    p = K.KGMT(**cfg, seed=3); p.set_obstacles(obs); p.begin(w.C1_INIT, w.C1_GOAL)
    ex = ShardedExpander(p, timing=True)
    st = ex.iterate(); st = ex.iterate()
    assert "compute_ms" in st and st["comm_bytes"] == 0
    st = p.iterate_many(2)                      # cooperative kernel continues from the sharded state
    hist = ex.run()
    r = p.result()
    for k in ("stop", "iterations", "tree_size", "cost_to_goal", "goal_index", "expansions"):
        assert r[k] == want[k], (k, r[k], want[k])
    _same_state(ref, p, want["tree_size"])


def test_forced_children_sweep_shape():
    """config 5 shape: P parents seeded as the frontier, M = P * children candidates in one sharded iteration."""
    obs = w.c2_obstacles(1000)
    P, children = 4096, 64
    M = P * children
    cfg = dict(w.C1, maxTreeSize=M + P, numIterations=3)
    parents = w.random_parents(P, obs, seed=7)
    outs = []
    for world in (1, 2):
        ranks = []
        for g in range(world):
            p = K.KGMT(**cfg, seed=5, max_candidates=M); p.set_obstacles(obs)
            p.seed_frontier(parents, w.C2_GOAL); p.set_children(children)
            ranks.append(p)
        deltas = [torch.zeros(ranks[0].shard_delta_ints(), dtype=torch.int32, device="cuda") for _ in ranks]
        st, infos, counts = _emulated_round(ranks, deltas, torch.device("cuda", 0))
        assert st["candidates"] == M and st["children"] == children
        outs.append((st, ranks[0].export(K.ARR_SAMPLES)[:st["tree_size"]].copy(), ranks[0].export(K.ARR_PARENT)[:st["tree_size"]].copy()))
    assert outs[0][0] == outs[1][0]
    np.testing.assert_array_equal(outs[0][1].view(np.uint32), outs[1][1].view(np.uint32))
    np.testing.assert_array_equal(outs[0][2], outs[1][2])


# ---------------------------------------------------------------- peer-memory exchange (no NCCL on the data path)
@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("cfgname", ["c1", "c2small"])
def test_peer_memory_iterations_equal_single_gpu(world, cfgname):
    """kgmt_peer_expand_*: accepted rows written straight into every replica's tree, deltas reduced through peer loads
    and stores, counts / goal through device mailboxes.  `world` contexts of this process share the GPU and run their
    exchanges concurrently on their own streams; every replica must equal the single-context run, bit for bit."""
    if cfgname == "c1":
        cfg, obs, init, goal = w.C1, w.C1_OBSTACLES, w.C1_INIT, w.C1_GOAL
    else:
        cfg, obs, init, goal = dict(w.C2, maxTreeSize=200000), w.c2_obstacles(1000), w.C2_INIT, w.C2_GOAL
    ref = K.KGMT(**cfg, seed=13); ref.set_obstacles(obs); ref.begin(init, goal)
    ranks = []
    for g in range(world):
        p = K.KGMT(**cfg, seed=13); p.set_obstacles(obs); p.begin(init, goal)
        ranks.append(p)
    for g, p in enumerate(ranks):
        p.peer_attach_local(g, ranks)
    for it in range(120):
        want = ref.iterate()
        for p in ranks:
            p.peer_expand_begin()
        got = [p.peer_expand_end() for p in ranks]
        assert all(g == want for g in got), (it, got[0], want)
        if it < 6 or want["stop"] != 0:
            for p in ranks:
                _same_state(ref, p, want["tree_size"])
        if want["stop"] != 0:
            break
    assert want["stop"] in (1, 2, 3, 4)
    if want["stop"] == 1:
        for p in ranks:
            np.testing.assert_array_equal(ref.extract_path(), p.extract_path())
    for p in ranks:
        p.peer_detach()


def test_peer_portfolio_race_flag():
    """kgmt_peer_race: the rank that reaches the goal tells the others through a word in their memory; a rank that finds
    the word set for the current race stops at its next iteration boundary with stop == 5 (peer solved).  Two contexts of
    this process take turns (two cooperative launches cannot share one GPU), which makes the outcome deterministic."""
    cfg, obs = w.C1, w.C1_OBSTACLES
    a = K.KGMT(**cfg, seed=3); a.set_obstacles(obs)          # the race lives in the grid-barrier loop
    b = K.KGMT(**cfg, seed=4); b.set_obstacles(obs)
    ref = K.KGMT(**cfg, seed=3); ref.set_obstacles(obs)
    want = ref.plan(w.C1_INIT, w.C1_GOAL)
    a.peer_attach_local(0, [a, b]); b.peer_attach_local(1, [a, b])
    ra = a.peer_race(w.C1_INIT, w.C1_GOAL, race_id=1)
    assert ra["stop"] == 1 and (ra["tree_size"], ra["iterations"], ra["cost_to_goal"]) == (want["tree_size"], want["iterations"], want["cost_to_goal"])
    rb = b.peer_race(w.C1_INIT, w.C1_GOAL, race_id=1)               # a has already won race 1
    assert rb["stop"] == 5 and rb["iterations"] == 1
    rb2 = b.peer_race(w.C1_INIT, w.C1_GOAL, race_id=2)              # a new race: nobody has solved it yet
    assert rb2["stop"] == 1 and rb2["iterations"] > 1
    ra2 = a.peer_race(w.C1_INIT, w.C1_GOAL, race_id=2)              # ... and now b has
    assert ra2["stop"] == 5
    # an ordinary plan ignores the race words
    assert a.plan(w.C1_INIT, w.C1_GOAL)["stop"] == 1
    a.peer_detach(); b.peer_detach()


def test_peer_exchange_times_out_instead_of_hanging():
    """A rank whose peer never arrives gets KGMT_ERR_COMM after ~5 s — the spin-wait kernels give up instead of hanging
    the GPU.  The replicas of an aborted exchange are undefined (rows and counters may be partly applied); after a fresh
    kgmt_begin on every rank the exchange works again."""
    import time
    cfg, obs = dict(w.C1, maxTreeSize=4000), w.C1_OBSTACLES
    ranks = []
    for g in range(2):
        p = K.KGMT(**cfg, seed=21); p.set_obstacles(obs); p.begin(w.C1_INIT, w.C1_GOAL)
        ranks.append(p)
    ref = K.KGMT(**cfg, seed=21); ref.set_obstacles(obs); ref.begin(w.C1_INIT, w.C1_GOAL)
    for g, p in enumerate(ranks):
        p.peer_attach_local(g, ranks)
    t0 = time.perf_counter()
    ranks[0].peer_expand_begin()                      # rank 1 does not take part
    with pytest.raises(K.KgmtError, match="did not arrive"):
        ranks[0].peer_expand_end()
    assert 4.0 < time.perf_counter() - t0 < 20.0
    ranks[1].peer_expand_begin()                      # rank 1 catches up with the exchange counter: its partner is gone
    with pytest.raises(K.KgmtError):
        ranks[1].peer_expand_end()
    for p in ranks:
        p.begin(w.C1_INIT, w.C1_GOAL)                 # restart the plan on every replica
    for it in range(3):
        want = ref.iterate()
        for p in ranks:
            p.peer_expand_begin()
        got = [p.peer_expand_end() for p in ranks]
        assert all(g == want for g in got), (it, got, want)
    for p in ranks:
        _same_state(ref, p, want["tree_size"])
        p.peer_detach()


def test_c5_full_size_one_iteration_property():
    """BASELINE config 5 at its largest size: ONE iteration of 2^26 candidates (32 768 seeded parents x 2 048 children,
    config-2 map).  The cooperative kernel and the sharded sequence (peer-memory path, one rank) must agree on every
    region counter, on the accepted count and on every parent link; counters must add up to the candidates."""
    obs = w.c2_obstacles(1000)
    P, M = 32768, 1 << 26
    parents = w.random_parents(P, obs, seed=7)
    cfg = dict(w.C1, maxTreeSize=M + P, numIterations=2)
    out = []
    for mode in ("coop", "peer"):
        p = K.KGMT(**cfg, seed=5, max_candidates=M); p.set_obstacles(obs)
        p.seed_frontier(parents, w.C2_GOAL); p.set_children(M // P)
        if mode == "coop":
            st = p.iterate()
        else:
            p.peer_attach_local(0, [p])
            st = p.peer_iterate()
            p.peer_detach()
        assert st["candidates"] == M and st["children"] == M // P and st["frontier"] == P
        maps = {k: p.export(k) for k in (K.ARR_R1, K.ARR_R1VALID, K.ARR_R1INVALID, K.ARR_R2, K.ARR_R2VALID, K.ARR_R2INVALID, K.ARR_R2AVAIL)}
        par = p.export(K.ARR_PARENT)[: st["tree_size"]].copy()
        out.append((st, maps, par))
        p.close()
    (sa, ma, pa), (sb, mb, pb) = out
    assert sa == sb
    for k in ma:
        np.testing.assert_array_equal(ma[k], mb[k])
    np.testing.assert_array_equal(pa, pb)
    # every candidate that ended inside the workspace is counted exactly once, valid or invalid
    assert int(ma[K.ARR_R1].sum()) == int(ma[K.ARR_R1VALID].sum()) + int(ma[K.ARR_R1INVALID].sum()) <= M + P
    assert int(ma[K.ARR_R1].sum()) >= int(0.8 * M)
    assert sa["accepted"] == sa["tree_size"] - P and (np.diff(pa[P:]) >= 0).all() and pa[P:].max() < P


# ------------------------------------------------- compute + exchange FUSED in one persistent kernel per rank
def _cfgs(name):
    if name == "c1":
        return w.C1, w.C1_OBSTACLES, w.C1_INIT, w.C1_GOAL
    return dict(w.C2, maxTreeSize=200000), w.c2_obstacles(1000), w.C2_INIT, w.C2_GOAL


@pytest.mark.parametrize("cfgname", ["c1", "c2small"])
def test_fused_sharded_kernel_one_rank_equals_single_gpu(cfgname):
    """kgmt_peer_plan / kgmt_peer_expand_iterations with a world of one (two cooperative kernels cannot share a GPU, so
    a single-GPU box can only run the fused kernel against itself): counts and goal through its own mailbox, rows packed
    into its own tree, deltas reduced from its own slab — every code path of the fused kernel except the remote addresses.
    Whole plans and stepped iterations must equal kgmt_plan / kgmt_expand_iteration bit for bit."""
    cfg, obs, init, goal = _cfgs(cfgname)
    ref = K.KGMT(**cfg, seed=17); ref.set_obstacles(obs)
    p = K.KGMT(**cfg, seed=17); p.set_obstacles(obs)
    p.peer_attach_local(0, [p])
    for seed in (17, 18):
        ref.set_seed(seed); p.set_seed(seed)
        want, got = ref.plan(init, goal), p.peer_plan(init, goal)
        for k in ("stop", "iterations", "tree_size", "cost_to_goal", "goal_index", "expansions"):
            assert want[k] == got[k], (seed, k, want[k], got[k])
        assert got["kernel_launches"] == 2
        _same_state(ref, p, want["tree_size"])
        if want["stop"] == 1:
            np.testing.assert_array_equal(ref.extract_path(), p.extract_path())
    # stepped: 1, then 2 iterations per launch; then the single-GPU loop carries on from the fused kernel's state
    ref.set_seed(19); p.set_seed(19)
    ref.begin(init, goal); p.begin(init, goal)
    assert ref.iterate() == p.peer_iterate_fused(1)
    ref.iterate(); a = ref.iterate()
    assert a == p.peer_iterate_fused(2)
    _same_state(ref, p, a["tree_size"])
    if a["stop"] == 0:
        assert ref.iterate() == p.iterate()
        assert ref.iterate() == p.peer_iterate_fused(1)
        _same_state(ref, p, ref.result()["tree_size"])
    p.peer_detach()


@pytest.mark.parametrize("cfgname", ["c1", "c2small", "c2"])
def test_fused_sharded_plans_on_several_gpus(cfgname):
    """Whole sharded plans on every GPU of the box (>= 2, else skipped): one context per GPU in this process, wired with
    kgmt_peer_attach_local (peer access over NVLink), one host thread per rank calling kgmt_peer_plan.  Every replica
    must hold the single-GPU tree, links, costs and maps bit for bit, and every rank must report the same result."""
    import threading
    G = min(torch.cuda.device_count(), 8)
    if G < 2:
        pytest.skip("needs at least 2 GPUs")
    cfg, obs, init, goal = (w.C2, w.c2_obstacles(1000), w.C2_INIT, w.C2_GOAL) if cfgname == "c2" else _cfgs(cfgname)
    ref = K.KGMT(**cfg, seed=23, device=0); ref.set_obstacles(obs)
    want = ref.plan(init, goal)
    ranks = []
    for g in range(G):
        p = K.KGMT(**cfg, seed=23, device=g); p.set_obstacles(obs)
        ranks.append(p)
    for g, p in enumerate(ranks):
        p.peer_attach_local(g, ranks)
    got, errs = [None] * G, []

    def run(g):
        try:
            got[g] = ranks[g].peer_plan(init, goal)
        except Exception as e:            # noqa: BLE001
            errs.append((g, repr(e)))
    for rep in range(2):                                  # the second plan re-uses the attached state (sequence numbers go on)
        th = [threading.Thread(target=run, args=(g,)) for g in range(G)]
        for t in th:
            t.start()
        for t in th:
            t.join(60)
        assert not errs, errs
        for g in range(G):
            for k in ("stop", "iterations", "tree_size", "cost_to_goal", "goal_index", "expansions"):
                assert want[k] == got[g][k], (g, k, want[k], got[g][k])
    for p in ranks:
        _same_state(ref, p, want["tree_size"])
        if want["stop"] == 1:
            np.testing.assert_array_equal(ref.extract_path(), p.extract_path())
    for p in ranks:
        p.peer_detach()
