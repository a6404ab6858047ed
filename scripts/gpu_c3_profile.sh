#!/bin/bash
# ncu captures of FULL-SIZE config-3 plans (10 000 obstacles, numDisc 40, maxTree 2^20): culled back end (--set full) and the
# TMA tile-streamed exhaustive back end (one 2.4 s launch: a reduced section list keeps the replay passes to a minute).
#   usage: bash scripts/gpu_c3_profile.sh <tag>
TAG=${1:-r02F}; O=gpurun_out; mkdir -p $O
cat > /tmp/c3_full.py <<'PY'
import sys; sys.path.insert(0, '.')
from cudasbmp_b200 import kgmt as K, workloads as w
mode = K.COLLIDE_BRUTE if sys.argv[1] == "stream" else K.COLLIDE_GRID
p = K.KGMT(**w.C3, seed=9, collision_mode=mode); p.set_obstacles(w.c3_obstacles(10000))
for s in ((1,) if sys.argv[1] == "stream" else (1, 2)):
    p.set_seed(s); print(p.plan(w.C2_INIT, w.C2_GOAL), p.config(), flush=True)
PY
timeout 120 python /tmp/c3_full.py grid > $O/${TAG}_c3_grid_plain.log 2>&1; echo "c3 grid plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:expand_kernel -c 1 -s 1 -f -o $O/${TAG}_c3grid python /tmp/c3_full.py grid > $O/${TAG}_c3_grid_ncu.log 2>&1; echo "c3 grid ncu rc=$?"
timeout 120 python /tmp/c3_full.py stream > $O/${TAG}_c3_stream_plain.log 2>&1; echo "c3 stream plain rc=$?"
timeout 600 ncu --section SpeedOfLight --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats \
    --section SourceCounters --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis \
    --clock-control none --import-source on -k regex:expand_kernel -c 1 -f -o $O/${TAG}_c3stream python /tmp/c3_full.py stream > $O/${TAG}_c3_stream_ncu.log 2>&1; echo "c3 stream ncu rc=$?"
tail -n 2 $O/${TAG}_c3_grid_plain.log $O/${TAG}_c3_stream_plain.log
