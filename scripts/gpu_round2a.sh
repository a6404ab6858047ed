#!/bin/bash
# Round-2 first GPU pass: parity tests, issue-rate microbenchmark, lane-refill A/B, bench, ncu captures for G=1 and G=3.
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> $O/${TAG}_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
timeout 120 scripts/bin/issue_rate_bench > $O/${TAG}_issue_rates.json 2>&1; echo "issue rc=$?"
timeout 600 python scripts/ab_groups.py 31 > $O/${TAG}_ab_groups.log 2>&1; echo "ab rc=$?"
timeout 600 python bench.py --no-cpu-baseline > $O/${TAG}_bench_g1.log 2>&1; echo "bench rc=$?"
KGMT_GROUP=3 timeout 600 python bench.py --no-cpu-baseline > $O/${TAG}_bench_g3.log 2>&1; echo "bench g3 rc=$?"
timeout 300 python scripts/iter_profile.py timeline > $O/${TAG}_timeline.log 2>&1; echo "timeline rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:expand_kernel -c 2 -f -o $O/${TAG}_prof \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_f.log 2>&1; echo "ncu full rc=$?"
KGMT_GROUP=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:expand_kernel -c 2 -f -o $O/${TAG}_prof_g3 \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_f_g3.log 2>&1; echo "ncu full g3 rc=$?"
tail -5 $O/${TAG}_pytest.log
cat $O/${TAG}_ab_groups.log | head -12
