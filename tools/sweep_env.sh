#!/bin/bash
# Tuning aid: time the linear extras under different run-time settings, e.g.
#   tools/sweep_env.sh CT_LINEAR_CHUNK_PAIRS 0 2 4 8
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --linear-only 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
c2 = d.get('config2_1035_pairs_960x540', {})
print('$var=$v', 'reinhard %.1f Gpix/s %.3f  mkl %.1f %.3f' % (d['reinhard_f32']['Mpix/s'] / 1e3, d['reinhard_f32']['frac_of_hbm'], d['mkl_f32_to_f64']['Mpix/s'] / 1e3, d['mkl_f32_to_f64']['frac_of_hbm']), ' | 1035 pairs:', ' '.join('%s %.3f' % (k.split('/')[0][:8] + ('*' if 'uniform' in k else ''), v['frac_of_hbm']) for k, v in c2.items()))"
done
