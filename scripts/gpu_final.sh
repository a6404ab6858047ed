#!/bin/bash
# what the driver runs at round end, on one GPU: smoke(), pytest -m gpu, bench.py (both arms)
TAG=${1:-final}; O=gpurun_out; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.log 2>&1; echo "bench ref rc=$?"
timeout 900 python bench.py > $O/${TAG}_bench.log 2> $O/${TAG}_bench.err; echo "bench rc=$?"
tail -2 $O/${TAG}_smoke.log; tail -3 $O/${TAG}_pytest.log; tail -c 300 $O/${TAG}_bench_ref.log; echo; head -c 700 $O/${TAG}_bench.log | tail -c 650
