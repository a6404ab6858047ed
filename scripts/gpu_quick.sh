#!/bin/bash
# quick GPU pass (1 GPU): parity tests + headline bench (+ optional timeline / ncu with NCU=1)
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
timeout 600 python bench.py --only-headline > $O/${TAG}_bench.log 2> $O/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python scripts/iter_profile.py timeline > $O/${TAG}_timeline.log 2>&1; echo "timeline rc=$?"
timeout 300 python scripts/ttfs_c1.py > $O/${TAG}_ttfs_c1.log 2>&1; echo "ttfs rc=$?"
if [ -n "$NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:expand_kernel -c 1 -f -o $O/${TAG}_prof \
    python bench.py --steps 1 --warmup 1 --plans-per-step 2 --only-headline > $O/${TAG}_ncu_f.log 2>&1; echo "ncu full rc=$?"
fi
tail -8 $O/${TAG}_pytest.log
python - <<P
import json
for l in open("$O/${TAG}_bench.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value %.4g exp/s  e2e %.4g  ms/plan %.4f  clocks %s  cfg %s" % (d["value"], d["e2e"]["value"], d["plan"]["device_ms_per_plan"], d["clocks"], d["collide_backend"]))
P
tail -3 $O/${TAG}_ttfs_c1.log
