#!/bin/bash
# A/B of the linear section (1035 pairs of 960x540) per variant library
cd $GRAFT_REPO_ROOT
for round in 1 2; do
for v in "$@"; do
  CT_B200_LIB=$PWD/tools/scratch/libs/libct_$v.so timeout 300 python bench.py --linear-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); l=d.get('linear', d)
print('$v round $round', {k.split('/')[0]+('/n' if 'noise' in k and 'smooth' not in k else ''): round(v['frac_of_hbm'],3) for k,v in l.items() if isinstance(v,dict) and 'frac_of_hbm' in v})"
done; done
