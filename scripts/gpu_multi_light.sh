#!/bin/bash
# light multi-GPU confirmation after host-side changes: multi-GPU parity tests, C++ demo, a short bench under torchrun
TAG=${1:-ml}; N=${2:-2}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -k "several_gpus or all_gpus or one_rank" -rs > $O/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"
( cd /tmp && timeout 300 $OLDPWD/cudasbmp_b200/bin/kgmt_multi_demo $N ) > $O/${TAG}_demo.log 2>&1; echo "demo rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --skip c5 > $O/${TAG}_bench_n$N.log 2> $O/${TAG}_bench_n$N.err; echo "bench n=$N rc=$?"
tail -3 $O/${TAG}_pytest_multi.log; tail -4 $O/${TAG}_demo.log; tail -1 $O/${TAG}_bench_n$N.log | cut -c1-300; tail -2 $O/${TAG}_bench_n$N.err
