#!/bin/bash
# compute-sanitizer over every kernel family (VERDICT r01 "next" 8).  usage: bash scripts/gpu_sanitize.sh <tag>
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
S=$O/${TAG}_sanitizer.txt
: > $S
for tool in memcheck racecheck initcheck synccheck; do
  for tgt in c1 c2 c3s batch peer2 fused1 stage; do
    echo "==== compute-sanitizer --tool $tool  python scripts/sanitize_targets.py $tgt" >> $S
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_targets.py $tgt > $O/${TAG}_san_${tool}_${tgt}.log 2>&1
    echo "rc=$?" >> $S
    grep -E "target ok|ERROR SUMMARY|RACECHECK SUMMARY|Error:|Race reported|hazard|Uninitialized|Invalid|AssertionError|Traceback" $O/${TAG}_san_${tool}_${tgt}.log | head -12 >> $S
  done
done
tail -80 $S
