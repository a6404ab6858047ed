"""torchrun --nproc-per-node N scripts/peer_check.py: the peer-memory sharded expansion across N real GPUs (cudaIpc) gives on
every rank exactly the tree a single GPU builds (checksums compared through an all-gather)."""
import os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from cudasbmp_b200 import kgmt as K, workloads as w
from cudasbmp_b200.sharded import PeerExpander
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def crc(p, T):
    h = zlib.crc32(p.export(K.ARR_SAMPLES)[:T].tobytes()); h = zlib.crc32(p.export(K.ARR_PARENT)[:T].tobytes(), h)
    h = zlib.crc32(p.export(K.ARR_COSTS)[:T].tobytes(), h)
    for m in (K.ARR_R1, K.ARR_R1VALID, K.ARR_R1INVALID, K.ARR_R1AVAIL, K.ARR_R2, K.ARR_R2VALID, K.ARR_R2INVALID, K.ARR_R2AVAIL, K.ARR_R1SCORE):
        h = zlib.crc32(p.export(m).tobytes(), h)
    return h
ok = True
for name, cfg, obs, init, goal in (("c1", w.C1, w.C1_OBSTACLES, w.C1_INIT, w.C1_GOAL), ("c2", w.C2, w.c2_obstacles(1000), w.C2_INIT, w.C2_GOAL)):
    ref = K.KGMT(**cfg, seed=17, device=local); ref.set_obstacles(obs); want = ref.plan(init, goal)
    p = K.KGMT(**cfg, seed=17, device=local); p.set_obstacles(obs); p.begin(init, goal)
    ex = PeerExpander(p)
    hist = ex.run()
    got = p.result()
    same = all(got[k] == want[k] for k in ("stop", "iterations", "tree_size", "expansions", "cost_to_goal", "goal_index")) and \
        crc(p, got["tree_size"]) == crc(ref, want["tree_size"])
    t = torch.tensor([1 if same else 0], device="cuda")
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0: print(name, "gpus", world, "iterations", len(hist), "tree", got["tree_size"], "IDENTICAL on every rank" if int(t) else "DIFFERENT", flush=True)
    ok &= bool(int(t))
    if world > 1: dist.barrier()
    ex.close(); p.close(); ref.close()
if world > 1: dist.destroy_process_group()
sys.exit(0 if ok else 1)
