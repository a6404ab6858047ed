"""Per-iteration device timeline of ONE fused sharded plan (kgmt_plan_sharded) next to the single-GPU loop, config 2.
Single process: world 1; under torchrun: every rank, rank 0 prints.   python scripts/fused_timeline.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cudasbmp_b200 import kgmt as K, workloads as w
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
box = [K.KGMT.comm_unique_id() if rank == 0 else None]
if dist is not None:
    dist.broadcast_object_list(box, src=0)
obs = w.c2_obstacles(1000)
p = K.KGMT(**w.C2, seed=1, device=local); p.set_obstacles(obs)
p.comm_init(rank, world, box[0])
p.iteration_log(True)
ms = []
for s in range(4):
    p.set_seed(1 + s); p.comm_barrier(); r = p.plan_sharded(w.C2_INIT, w.C2_GOAL); ms.append(r["device_ms"])
log = p.iteration_log()
if rank == 0:
    print("fused sharded plan, world", world, r, "ms", ms)
    for i, row in enumerate(log):
        r_ = row.astype(np.int64)
        ph = [(r_[k] - r_[2]) / 1e3 if r_[k] else float("nan") for k in (3, 4, 5, 6, 0, 7)]
        print("itr %2d  M %8d acc %7d | A done %6.1f  counts in %6.1f  pack+reduce done %6.1f  goal in %6.1f  finish done %6.1f  barrier 3 %6.1f us" % (
            i + 1, r_[1] >> 32, r_[1] & 0xFFFFFFFF, *ph))
if world == 1:
    q = K.KGMT(**w.C2, seed=1, device=local); q.set_obstacles(obs)
    print("single-GPU loop:", [round(q.plan(w.C2_INIT, w.C2_GOAL)["device_ms"], 3) for _ in range(4)])
p.comm_destroy()
if dist is not None:
    dist.destroy_process_group()
