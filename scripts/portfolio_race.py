"""torchrun --nproc-per-node N scripts/portfolio_race.py: median time-to-first-solution with a portfolio of N seeds, one per
GPU, first-solution termination through peer-memory flags (kgmt_peer_race).  Config 1 (reference demo) and config 2.
Per race: every rank plans the same query with seed base + rank; host wall clock from a common barrier to the return of
the rank's call; time-to-first-solution = the earliest return among the ranks that solved; 'all stopped' = the latest."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from cudasbmp_b200 import kgmt as K, workloads as w
from cudasbmp_b200.sharded import PeerExpander
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
races = int(sys.argv[1]) if len(sys.argv) > 1 else 101
race_id = 0
for name, cfg, obs, init, goal in (("c1", w.C1, w.C1_OBSTACLES, w.C1_INIT, w.C1_GOAL), ("c2", w.C2, w.c2_obstacles(1000), w.C2_INIT, w.C2_GOAL)):
    p = K.KGMT(**cfg, seed=1, device=local); p.set_obstacles(obs)
    ex = PeerExpander(p)                                   # exchanges the cudaIpc handles, attaches
    rows = []
    for q in range(races + 3):
        race_id += 1
        p.set_seed(1000 * q + rank + 1)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t0 = time.perf_counter()
        r = p.peer_race(init, goal, race_id)
        dt = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([dt if r["stop"] == 1 else 1e9, dt, float(r["stop"] == 1), r["device_ms"]], dtype=torch.float64, device="cuda")
        if world > 1:
            g = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(g, t)
            g = torch.stack(g).cpu().numpy()
        else:
            g = t.cpu().numpy()[None]
        if q >= 3:
            rows.append((g[:, 0].min(), g[:, 1].max(), g[:, 2].sum(), g[:, 3].max()))
    rows = np.array(rows)
    solved = rows[rows[:, 0] < 1e8]
    if rank == 0:
        print(json.dumps(dict(config=name, gpus=world, races=races, solved=int(len(solved)),
                              ttfs_median_ms=float(np.median(solved[:, 0])), ttfs_p95_ms=float(np.percentile(solved[:, 0], 95)),
                              all_stopped_median_ms=float(np.median(rows[:, 1])), winners_per_race_mean=float(rows[:, 2].mean()),
                              device_ms_max_median=float(np.median(rows[:, 3])))), flush=True)
    if world > 1: dist.barrier()
    ex.close(); p.close()
if world > 1: dist.destroy_process_group()
