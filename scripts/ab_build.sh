#!/bin/bash
# A/B of compile-time variants of the library ON the GPU box (same image, nvcc present).
# usage: bash scripts/ab_build.sh "<EXTRA flags variant 1>" "<variant 2>" ...
for v in "$@"; do
  touch cudasbmp_b200/csrc/kgmt_capi.cu
  make -s -C cudasbmp_b200/csrc EXTRA="$v" 2>&1 | grep -v "^$" | head -3
  cuobjdump -res-usage cudasbmp_b200/libkgmt_b200.so 2>/dev/null | grep -A1 "expand_kernelILi0ELb0" | tail -1 | cut -c1-40
  for i in 1 2; do
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('variant [$v] run $i: %.3f G exp/s  %.4f ms/plan  ttfs_c1 %.4f ms grid %d' % (d['value']/1e9, d['ms_per_step'], d['ttfs']['device_median_ms'], d['collide_backend']['grid']))"
  done
done
touch cudasbmp_b200/csrc/kgmt_capi.cu; make -s -C cudasbmp_b200/csrc
