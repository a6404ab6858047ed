#!/bin/bash
for v in "$@"; do
  touch cudasbmp_b200/csrc/kgmt_capi.cu
  make -s -C cudasbmp_b200/csrc EXTRA="$v" 2>&1 | grep -v "^$" | head -3
  cuobjdump -res-usage cudasbmp_b200/libkgmt_b200.so 2>/dev/null | grep -A1 "batch_kernelILi0" | tail -1 | cut -c1-40
  python scripts/batch_bench.py 1024 2>&1 | grep '"cluster": [12],' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('variant [$v] cluster %d ws %d device_ms %.3f q/s %.0f' % (d['cluster'], d['workspaces'], d['device_ms'], d['queries_per_s_device']))"
done
touch cudasbmp_b200/csrc/kgmt_capi.cu; make -s -C cudasbmp_b200/csrc
