#!/bin/bash
# ncu evidence for the kernels besides the C2 cooperative kernel: C3 tile-streamed exhaustive back end (TMA), the sharded /
# peer-memory sequence (launch list), batch kernel.  usage: bash scripts/gpu_profiles2.sh <tag>
TAG=${1:-r01f}; O=gpurun_out; mkdir -p $O
cat > /tmp/c3_small.py <<'PY'
import sys; sys.path.insert(0, '.')
from cudasbmp_b200 import kgmt as K, workloads as w
cfg = dict(w.C3, maxTreeSize=20000)
p = K.KGMT(**cfg, seed=9, collision_mode=K.COLLIDE_BRUTE); p.set_obstacles(w.c3_obstacles(10000))
for s in (1, 2): p.set_seed(s); print(p.plan(w.C2_INIT, w.C2_GOAL), p.config())
PY
timeout 120 python /tmp/c3_small.py > $O/${TAG}_c3_plain.log 2>&1; echo "c3 plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:expand_kernel -c 1 -s 1 -f -o $O/${TAG}_c3stream python /tmp/c3_small.py > $O/${TAG}_c3_ncu.log 2>&1; echo "c3 ncu rc=$?"
cat > /tmp/peer_local.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from cudasbmp_b200 import kgmt as K, workloads as w
obs = w.c2_obstacles(1000); P, M = 32768, 1 << 22
parents = w.random_parents(P, obs, seed=7)
ranks = []
for g in range(2):
    p = K.KGMT(**dict(w.C1, maxTreeSize=M + P, numIterations=3), seed=5, max_candidates=M); p.set_obstacles(obs)
    p.seed_frontier(parents, w.C2_GOAL); p.set_children(M // P); ranks.append(p)
for g, p in enumerate(ranks): p.peer_attach_local(g, ranks)
for p in ranks: p.peer_expand_begin()
print([p.peer_expand_end() for p in ranks])
PY
timeout 120 python /tmp/peer_local.py > $O/${TAG}_peer_plain.log 2>&1; echo "peer plain rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${TAG}_launches_peer.csv python /tmp/peer_local.py > $O/${TAG}_peer_ncu.log 2>&1; echo "peer launches rc=$?"
