#!/bin/bash
# A/B on the GPU box: default build (3 CTAs/SM register bound) against -DKGMT_EXPAND_MIN_CTAS=4, cull-grid sweep for each
O=gpurun_out; mkdir -p $O
echo "== default build" > $O/$1_ab_ctas.log
python scripts/sweep_cull.py 36,40,44,48 0 >> $O/$1_ab_ctas.log 2>&1
python scripts/ttfs_c1.py >> $O/$1_ab_ctas.log 2>&1
touch cudasbmp_b200/csrc/kgmt_capi.cu
make -s -C cudasbmp_b200/csrc EXTRA="-DKGMT_EXPAND_MIN_CTAS=4" 2>&1 | grep -v "^$" | head -3
echo "== -DKGMT_EXPAND_MIN_CTAS=4" >> $O/$1_ab_ctas.log
cuobjdump -res-usage cudasbmp_b200/libkgmt_b200.so 2>/dev/null | grep -A1 "expand_kernelILi0ELb0" | tail -1 | cut -c1-60 >> $O/$1_ab_ctas.log
python scripts/sweep_cull.py 32,36,40,44,48 0,3 >> $O/$1_ab_ctas.log 2>&1
python scripts/ttfs_c1.py >> $O/$1_ab_ctas.log 2>&1
cat $O/$1_ab_ctas.log
