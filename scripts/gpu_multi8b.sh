#!/bin/bash
# 8-GPU pass (round 2, final kernels): multi-GPU parity tests, C++ demo, fused timeline, bench at N.  bash scripts/gpu_multi8b.sh <tag> <n>
TAG=${1:-r02p8}
N=${2:-8}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $O/${TAG}_smi.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -k "several_gpus or all_gpus" -rs > $O/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a $O/${TAG}_pytest_multi.log
( cd /tmp && timeout 300 $OLDPWD/cudasbmp_b200/bin/kgmt_multi_demo $N ) > $O/${TAG}_demo.log 2>&1; echo "demo rc=$?" | tee -a $O/${TAG}_demo.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 scripts/fused_timeline.py > $O/${TAG}_fused_timeline.log 2>&1; echo "timeline rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_n$N.log 2> $O/${TAG}_bench_n$N.err; echo "bench n=$N rc=$?"
tail -4 $O/${TAG}_pytest_multi.log
tail -9 $O/${TAG}_demo.log
grep -v "OMP\|\*\*\*\|^$" $O/${TAG}_fused_timeline.log | grep -v "^itr 1[0-9]\|^itr  [6-9]" | tail -9
tail -1 $O/${TAG}_bench_n$N.log | cut -c1-200
tail -3 $O/${TAG}_bench_n$N.err
