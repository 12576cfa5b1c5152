#!/bin/bash
# A/B with per-kernel times: full default bench (no cpu baseline, no extras) per variant; prints kernel_ms_per_pass
cd $GRAFT_REPO_ROOT
O=gpurun_out/abp; mkdir -p $O; : > $O/ab.txt
for round in 1 2; do
for v in "$@"; do
  CT_B200_LIB=$PWD/tools/scratch/libs/libct_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --e2e-frames 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$v round $round', round(d['ms_per_step']/d['passes_per_step'],4), 'ms/pass', r['kernel_ms_per_pass'], 'step_frac', round(r['step_frac'],3), d['clocks']['sm_mhz'])" >> $O/ab.txt
done; done
cat $O/ab.txt
