#!/bin/bash
# sharded-expansion tests + config-5 sweep on the GPUs of this box.  usage: bash scripts/gpu_shard.sh <tag> <ngpus> [max-log2]
TAG=${1:-r01s}; N=${2:-1}; MAXLG=${3:-26}
O=gpurun_out; mkdir -p $O
if [ "$N" = "1" ]; then
  timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q > $O/${TAG}_pytest_sharded.log 2>&1; echo "pytest sharded rc=$?"; tail -15 $O/${TAG}_pytest_sharded.log
  timeout 900 python scripts/sweep_c5.py --max-log2 $MAXLG > $O/${TAG}_c5_n1.log 2>&1; echo "sweep rc=$?"; tail -12 $O/${TAG}_c5_n1.log
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      scripts/sweep_c5.py --max-log2 $MAXLG > $O/${TAG}_c5_n$N.log 2>&1; echo "sweep rc=$?"; tail -12 $O/${TAG}_c5_n$N.log
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      scripts/multi_gpu_demo.py ${Q:-1024} > $O/${TAG}_c4_n$N.log 2>&1; echo "c4 rc=$?"; tail -6 $O/${TAG}_c4_n$N.log
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_bench_n$N.log 2>&1; echo "bench rc=$?"; tail -1 $O/${TAG}_bench_n$N.log | cut -c1-400
fi
