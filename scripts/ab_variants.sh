#!/bin/bash
# A/B of compile-time variants ON the GPU box: bash scripts/ab_variants.sh <tag> "<EXTRA 1>" "<EXTRA 2>" ...
TAG=$1; shift
O=gpurun_out; mkdir -p $O; L=$O/${TAG}_ab.log; : > $L
for v in "$@"; do
  touch cudasbmp_b200/csrc/kgmt_capi.cu
  make -s -C cudasbmp_b200/csrc EXTRA="$v" 2>&1 | grep -v "^$\|warning\|\^\|Remark" | head -3
  echo "== variant [$v] $(cuobjdump -res-usage cudasbmp_b200/libkgmt_b200.so 2>/dev/null | grep -A1 'expand_kernelILi0ELb0' | tail -1 | cut -c1-40)" >> $L
  for cull in ${CULLS:-0}; do
    KGMT_CULL_CELLS=$cull python bench.py --only-headline --steps 10 | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('  cull $cull: %.3f G exp/s  %.4f ms/plan  %s' % (d['value']/1e9, d['plan']['device_ms_per_plan'], d['collide_backend']))" >> $L
  done
  python scripts/ttfs_c1.py >> $L 2>&1
done
cat $L
