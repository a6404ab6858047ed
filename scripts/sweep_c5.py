"""BASELINE config 5: standalone propagate+collision(+maps+insertion) throughput sweep, M = 2^20 .. 2^26 candidate
expansions in ONE iteration, sharded-expansion mode with NCCL all-gather / all-reduce at 1/2/4/8 GPUs.

    python scripts/sweep_c5.py [--max-log2 26]                                  # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/sweep_c5.py                                                     # N GPUs

32 768 parents uniform in the free space of the map (theta in (-pi, pi], v in [-2, 2]) are seeded as the frontier, each
is expanded M / 32 768 times; K in {5 (reference demo map), 1 000 (config-2 map)}; numDisc = 10; N = 16, n = 8.
Per (K, M): best of 3 rounds; compute time (expand + pack + commit kernels) and communication time (count all-gather,
row all-gather, delta all-reduce) are CUDA-event times on the shared stream, max over ranks.  One JSON line per point.
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from cudasbmp_b200 import kgmt as K, workloads as w
from cudasbmp_b200.sharded import PeerExpander, ShardedExpander

ap = argparse.ArgumentParser()
ap.add_argument("--min-log2", type=int, default=20)
ap.add_argument("--max-log2", type=int, default=26)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--parents", type=int, default=32768)
ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer"], help="nccl: pack -> all-gather/all-reduce -> commit; peer: the library's own kernels over peer memory")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P = args.parents
for Kobs, obs in ((5, w.C1_OBSTACLES), (1000, w.c2_obstacles(1000))):
    parents = w.random_parents(P, obs, seed=7)
    for lg in range(args.min_log2, args.max_log2 + 1, 2):
        M = 1 << lg
        children = M // P
        cfg = dict(w.C1, maxTreeSize=M + P, numIterations=4)
        p = K.KGMT(**cfg, seed=5, device=local, max_candidates=M)
        p.set_obstacles(obs)
        ex = ShardedExpander(p, timing=True) if args.exchange == "nccl" else PeerExpander(p, timing=True)
        best = None
        for rep in range(args.reps + 1):                       # first round is the warm-up
            p.set_seed(5 + rep)
            p.seed_frontier(parents, w.C2_GOAL)
            p.set_children(children)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            st = ex.iterate()
            if args.exchange == "peer":                        # one fused sequence: no separate communication time
                st.update(compute_ms=st["total_ms"], comm_ms=0.0, comm_bytes=0, expand_ms=st["total_ms"], pack_ms=0.0, commit_ms=0.0)
            t = torch.tensor([st["compute_ms"], st["comm_ms"], st["compute_ms"] + st["comm_ms"]], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            comp, comm, tot = (float(v) for v in t.tolist())
            if rep > 0 and (best is None or tot < best["total_ms"]):
                best = dict(exchange=args.exchange, K=Kobs, log2M=lg, M=M, gpus=world, accepted=st["accepted"], accept_ratio=st["accepted"] / M,
                            compute_ms=comp, comm_ms=comm, total_ms=tot,
                            rank0_split_ms=[round(st[k], 4) for k in ("expand_ms", "pack_ms", "commit_ms")], comm_bytes=st["comm_bytes"],
                            expansions_per_s=M / tot * 1e3, expansions_per_s_compute_only=M / comp * 1e3)
        # single-GPU cooperative kernel on the same iteration, for reference (world 1 only)
        if world == 1:
            ms = []
            for rep in range(args.reps):
                p.set_seed(5 + rep + 1)
                p.seed_frontier(parents, w.C2_GOAL); p.set_children(children)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); p.iterate(); e1.record(); torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            best["cooperative_kernel_ms"] = min(ms)
            best["cooperative_expansions_per_s"] = M / min(ms) * 1e3
        if rank == 0:
            print(json.dumps(best), flush=True)
        if args.exchange == "peer":
            if world > 1:
                dist.barrier()
            ex.close()
        p.close(); del ex, p
        torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
