#!/bin/bash
# Round-2 re-entry pass (1 GPU): parity tests, headline bench, ncu launch list + full capture with source.
TAG=${1:-r02f}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> $O/${TAG}_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
timeout 600 python bench.py --only-headline > $O/${TAG}_bench.log 2> $O/${TAG}_bench.err; echo "bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:expand_kernel -c 1 -f -o $O/${TAG}_prof \
    python bench.py --steps 1 --warmup 1 --plans-per-step 2 --only-headline > $O/${TAG}_ncu_f.log 2>&1; echo "ncu full rc=$?"
tail -5 $O/${TAG}_pytest.log
tail -c 1500 $O/${TAG}_bench.log
