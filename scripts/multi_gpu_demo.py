"""torchrun --nproc-per-node N scripts/multi_gpu_demo.py [Q]: config 4 (batched planning, Q independent queries on the C1
map sharded over N GPUs) and the portfolio mode (same query, one seed per GPU, NCCL first-solution termination)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from cudasbmp_b200 import kgmt as K, workloads as w, multi

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
inits, goals = w.random_queries(Q, w.C1_OBSTACLES)
plan = K.KGMT(**w.C1, seed=1, device=local)
plan.set_obstacles(w.C1_OBSTACLES)
multi.plan_batch(plan, inits, goals, seeds=list(range(Q)))          # warm-up (allocates the per-query workspaces)
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
table = multi.plan_batch(plan, inits, goals, seeds=list(range(Q)))
torch.cuda.synchronize()
if world > 1: dist.barrier()
dt = time.perf_counter() - t0
solved = table[:, 2] == 1
res = dict(mode="batch (one launch per rank)", gpus=world, queries=Q, wall_s=dt, queries_per_s=Q / dt, expansions_per_s=table[:, 6].sum() / dt,
           solved=int(solved.sum()), device_ms_median=float(np.median(table[:, 7])), device_ms_p95=float(np.percentile(table[:, 7], 95)),
           stops={int(k): int((table[:, 2] == k).sum()) for k in np.unique(table[:, 2])})
if rank == 0: print(json.dumps(res))
# portfolio on the C2 query
obs2 = w.c2_obstacles(1000)
p2 = K.KGMT(**w.C2, seed=1, device=local); p2.set_obstacles(obs2)
for trial in range(3):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    r = multi.plan_portfolio(p2, w.C2_INIT, w.C2_GOAL, base_seed=100 + 10 * trial, check_every=4)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps(dict(mode="portfolio", gpus=world, trial=trial, wall_ms=dt * 1e3, winner=r["winner"], cost=r["cost"],
                              checks=r["checks"], path_len=len(r["path"]))))
if world > 1: dist.destroy_process_group()
