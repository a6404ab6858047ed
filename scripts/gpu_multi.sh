#!/bin/bash
# Multi-GPU pass on one box: bash scripts/gpu_multi.sh <tag> <ngpus>
TAG=${1:-r02m}
N=${2:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $O/${TAG}_smi.log 2>&1
nvidia-smi topo -m >> $O/${TAG}_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "several_gpus or all_gpus or one_rank" -rs > $O/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a $O/${TAG}_pytest_multi.log
( cd /tmp && timeout 300 $OLDPWD/cudasbmp_b200/bin/kgmt_multi_demo $N ) > $O/${TAG}_demo.log 2>&1; echo "demo rc=$?" | tee -a $O/${TAG}_demo.log
for n in $(seq 1 $N); do
  if [ $n -eq 1 ] || [ $n -eq 2 ] || [ $n -eq 4 ] || [ $n -eq 8 ]; then
    if [ $n -eq 1 ]; then
      timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --skip c3,same,fp32 > $O/${TAG}_bench_n$n.log 2> $O/${TAG}_bench_n$n.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 \
        bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_n$n.log 2> $O/${TAG}_bench_n$n.err
    fi
    echo "bench n=$n rc=$?"
  fi
done
tail -3 $O/${TAG}_pytest_multi.log
cat $O/${TAG}_demo.log | tail -12
for n in 1 2 4 8; do [ -f $O/${TAG}_bench_n$n.log ] && tail -1 $O/${TAG}_bench_n$n.log | cut -c1-300; done
tail -5 $O/${TAG}_bench_n$N.err
