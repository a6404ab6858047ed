#!/bin/bash
# The bounds-checked twin of the library over every kernel family (the stand-in for compute-sanitizer, which this GPU
# pool refuses: profiles/r02b_sanitizer_refused.txt).   usage: bash scripts/bounds_check.sh <tag>
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
S=$O/${TAG}_bounds_check.txt
echo "# KGMT_LIB=cudasbmp_b200/libkgmt_b200_check.so (-DKGMT_BOUNDS_CHECK) python scripts/sanitize_targets.py <target>" > $S
for tgt in c1 c2 c3s batch peer2 fused1 stage; do
  echo "==== $tgt" >> $S
  KGMT_LIB=$PWD/cudasbmp_b200/libkgmt_b200_check.so timeout 600 python scripts/sanitize_targets.py $tgt > $O/${TAG}_bc_$tgt.log 2>&1
  echo "rc=$?" >> $S
  grep -E "bounds checks|target ok|Error|Traceback|assert" $O/${TAG}_bc_$tgt.log | head -6 >> $S
done
# and the whole GPU parity suite on the checked library
KGMT_LIB=$PWD/cudasbmp_b200/libkgmt_b200_check.so timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_facade.py > $O/${TAG}_bc_pytest.log 2>&1
echo "==== pytest -m gpu on the checked library: rc=$?" >> $S
tail -2 $O/${TAG}_bc_pytest.log >> $S
KGMT_LIB=$PWD/cudasbmp_b200/libkgmt_b200_check.so python - >> $S 2>&1 <<'PY'
import sys; sys.path.insert(0, ".")
from cudasbmp_b200 import kgmt as K, workloads as w
p = K.KGMT(**w.C2, seed=1); p.set_obstacles(w.c2_obstacles(1000))
r = p.plan(w.C2_INIT, w.C2_GOAL)
print("full-size config-2 plan on the checked library:", r["tree_size"], "nodes,", r["expansions"], "expansions ->", p.debug_checks())
PY
cat $S
