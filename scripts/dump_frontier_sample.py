"""Makes bench_data/c2_frontier_sample.npz (run on a B200): 30 000 frontier nodes of a REAL config-2 plan (seed 1), taken
from the frontiers of several iterations — the parents whose children the planner actually expands.  bench.py's reference
arms (the reference's host loop and its CUDA kernel) and the same_population section expand these parents x 32 children,
so every arm sees the same candidate population as the GPU arm.   python scripts/dump_frontier_sample.py [out.npz]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cudasbmp_b200 import kgmt as K, workloads as w

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/c2_frontier_sample.npz"
p = K.KGMT(**w.C2, seed=1)
p.set_obstacles(w.c2_obstacles(1000))
p.begin(w.C2_INIT, w.C2_GOAL)
take = {6: 6000, 9: 6000, 12: 6000, 15: 6000, 18: 6000}
rows, its = [], []
rng = np.random.default_rng(0)
prev = 1
for it in range(1, 40):
    st = p.iterate()
    if it in take and st["accepted"] > 0:
        lo, hi = st["tree_size"] - st["accepted"], st["tree_size"]          # the frontier the NEXT iteration expands
        tree = p.export(K.ARR_SAMPLES)[lo:hi]
        pick = rng.choice(len(tree), size=min(take[it], len(tree)), replace=False)
        rows.append(tree[np.sort(pick)].copy()); its.append(it)
    if st["stop"] != 0:
        break
parents = np.concatenate(rows).astype(np.float32)
parents[:, 4:] = 0.0
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
np.savez_compressed(out, parents=parents, seed=np.int32(1), iterations=np.array(its, dtype=np.int32))
print("wrote", out, parents.shape, "iterations", its, "speed |v| mean %.2f max %.2f" % (np.abs(parents[:, 3]).mean(), np.abs(parents[:, 3]).max()))
