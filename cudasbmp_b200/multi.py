"""Multi-GPU modes of the KGMT path on one 8xB200 box: one process per GPU, torch.distributed for the plumbing
(NCCL over NVLink/NVSwitch on GPUs; the same code runs over gloo in the CPU tests with a stand-in planner).

The reference is single-GPU; these modes are the ones BASELINE.json's north_star names:

  plan_batch      primary: independent planning queries (or seeds) are sharded over the ranks.  No data-path
                  collective; one all-gather of the fixed-size result rows at the end.            (config 4)
  plan_portfolio  the same query with a different seed per rank; every `check_every` iterations one 8-byte
                  all-reduce(MIN) of (cost bits << 8 | rank) tells every rank whether somebody has reached the goal:
                  first-solution termination.  The winner's path is broadcast.
  A single query's tree does not partition across GPUs (insertion and the region maps are global state touched by
  every candidate): replicas only.  The sharded-expansion mode (candidates of one iteration split over the ranks) lives
  in cudasbmp_b200/sharded.py.
"""
import numpy as np
import torch
import torch.distributed as dist

RESULT_COLS = ("query", "rank", "stop", "iterations", "tree_size", "cost_to_goal", "expansions", "device_ms")


def shard_range(n_items, rank, world):
    """Contiguous, balanced shard [lo, hi) of n_items for `rank`."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _device_for(group):
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def plan_batch(planner, inits, goals, seeds, group=None, batched=True, cluster_size=2):
    """Shard Q queries over the ranks; every rank returns the full [Q, 8] float64 result table (RESULT_COLS).

    planner: cudasbmp_b200.KGMT with the obstacles already set.  batched=True plans the rank's shard in one launch
    (kgmt_plan_batch); batched=False calls plan() once per query.  Either way every query's result is the same.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    Q = len(seeds)
    lo, hi = shard_range(Q, rank, world)
    rows = np.zeros((Q, len(RESULT_COLS)), dtype=np.float64)
    if hi > lo and batched and hasattr(planner, "plan_batch"):
        # the rank's whole shard in ONE launch: thread-block clusters plan one query each (kgmt_plan_batch)
        res, ms, _, _ = planner.plan_batch(inits[lo:hi], goals[lo:hi], np.asarray(seeds[lo:hi]), cluster_size=cluster_size)
        for i, r in enumerate(res):
            rows[lo + i] = (lo + i, rank, r["stop"], r["iterations"], r["tree_size"], r["cost_to_goal"], r["expansions"], ms)
    else:
        for q in range(lo, hi):
            planner.set_seed(int(seeds[q]))
            r = planner.plan(inits[q], goals[q])
            rows[q] = (q, rank, r["stop"], r["iterations"], r["tree_size"], r["cost_to_goal"], r["expansions"], r["device_ms"])
    if world == 1:
        return rows
    # disjoint shards: a SUM all-reduce of the zero-padded table is the gather (one collective, Q x 64 bytes)
    t = torch.from_numpy(rows).to(_device_for(group))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def _encode(cost, rank):
    """(cost, rank) -> int64 whose MIN picks the lowest cost, ties to the lowest rank; 'no solution' sorts last."""
    if cost <= 0.0 or not np.isfinite(cost):
        return (0x7F800000 << 8) | 0xFF
    return (int(np.float32(cost).view(np.uint32)) << 8) | (rank & 0xFF)


def plan_portfolio(planner, init7, goal7, base_seed, check_every=2, max_checks=10_000, group=None):
    """Same query, seed base_seed+rank on every rank; stops all ranks as soon as one has a solution.

    planner: begin(init, goal), iterate_many(k) -> stats dict with 'stop' and 'cost_to_goal', extract_path().
    Returns dict(winner, cost, checks, iterations (this rank), stop (this rank), path (np [L,7], on every rank)).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = _device_for(group) if world > 1 else torch.device("cpu")
    planner.set_seed(int(base_seed) + rank)
    planner.begin(init7, goal7)
    st = {"stop": 0, "cost_to_goal": 0.0, "iteration": 0}
    checks, winner, cost = 0, -1, 0.0
    while checks < max_checks:
        if st["stop"] == 0:
            st = planner.iterate_many(check_every)
        checks += 1
        mine = _encode(st["cost_to_goal"] if st["stop"] == 1 else 0.0, rank)
        # word 0: best (cost, rank); word 1: 0 iff somebody is still running
        flag = torch.tensor([mine, 0 if st["stop"] == 0 else 1], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        best, all_stopped = int(flag[0]), int(flag[1])
        if (best >> 8) != 0x7F800000:
            winner = best & 0xFF
            cost = float(np.uint32(best >> 8).view(np.float32))
            break
        if all_stopped:
            break
    # the winner's path to every rank (length first, then the rows)
    path = np.zeros((0, 7), dtype=np.float32)
    if winner >= 0:
        if rank == winner:
            path = planner.extract_path()
        n = torch.tensor([len(path)], dtype=torch.int64, device=dev)
        if world > 1:
            dist.broadcast(n, src=winner, group=group)
        buf = torch.zeros((int(n[0]), 7), dtype=torch.float32, device=dev)
        if rank == winner:
            buf.copy_(torch.from_numpy(np.ascontiguousarray(path)))
        if world > 1:
            dist.broadcast(buf, src=winner, group=group)
        path = buf.cpu().numpy()
    return dict(winner=winner, cost=cost, checks=checks, iterations=st.get("iteration", 0), stop=st["stop"], path=path)
