#!/bin/bash
# One GPU-box pass: parity tests, bench (ours + reference arm), baseline A, iteration timeline, ncu launch list + full capture.
# usage (from the repo root on the GPU box):  bash scripts/gpu_round.sh <tag>
TAG=${1:-r01x}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> $O/${TAG}_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
timeout 600 python bench.py > $O/${TAG}_bench.log 2>&1; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.log 2>&1; echo "bench ref rc=$?"
timeout 300 python scripts/iter_profile.py timeline > $O/${TAG}_timeline.log 2>&1; echo "timeline rc=$?"
timeout 600 python scripts/baseline_ref_gpu.py > $O/${TAG}_baselineA.log 2>&1; echo "baselineA rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_l.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:expand_kernel -c 2 -f -o $O/${TAG}_prof \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_f.log 2>&1; echo "ncu full rc=$?"
tail -3 $O/${TAG}_pytest.log
tail -1 $O/${TAG}_bench.log | cut -c1-600
