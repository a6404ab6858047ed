#!/bin/bash
# full single-GPU pass: parity tests, complete bench line (all sections), reference arm
TAG=${1:-full}
O=gpurun_out
mkdir -p $O
lscpu | grep -E "Model name|^CPU\(s\)" > $O/${TAG}_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
timeout 900 python bench.py > $O/${TAG}_bench.log 2> $O/${TAG}_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.log 2>&1; echo "bench ref rc=$?"
tail -4 $O/${TAG}_pytest.log
tail -c 600 $O/${TAG}_bench.log
tail -3 $O/${TAG}_bench.err
tail -c 400 $O/${TAG}_bench_ref.log
