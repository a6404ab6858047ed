#!/bin/bash
# Round-2 second GPU pass (1 GPU): parity tests, C++ demos, frontier sample, full bench (both arms), timeline, ncu, sanitizer.
TAG=${1:-r02b}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> $O/${TAG}_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
timeout 300 python scripts/dump_frontier_sample.py $O/c2_frontier_sample.npz > $O/${TAG}_sample.log 2>&1; echo "sample rc=$?"
mkdir -p bench_data; cp $O/c2_frontier_sample.npz bench_data/ 2>/dev/null
timeout 900 python bench.py > $O/${TAG}_bench.log 2> $O/${TAG}_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.log 2>&1; echo "bench ref rc=$?"
timeout 300 python scripts/iter_profile.py timeline > $O/${TAG}_timeline.log 2>&1; echo "timeline rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --plans-per-step 2 --only-headline > $O/${TAG}_ncu_l.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:expand_kernel -c 2 -f -o $O/${TAG}_prof \
    python bench.py --steps 1 --warmup 1 --plans-per-step 2 --only-headline > $O/${TAG}_ncu_f.log 2>&1; echo "ncu full rc=$?"
[ -n "$SKIP_SANITIZER" ] || bash scripts/gpu_sanitize.sh $TAG > $O/${TAG}_sanitize_tail.log 2>&1; echo "sanitize rc=$?"
tail -5 $O/${TAG}_pytest.log
tail -c 1500 $O/${TAG}_bench.log
[ -n "$SKIP_SANITIZER" ] || tail -30 $O/${TAG}_sanitizer.txt
