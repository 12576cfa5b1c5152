#!/bin/bash
# Tuning aid: rebuild the library on the GPU box with different -D settings and time the linear extras.
# usage: tools/sweep_linear.sh "-DCT_STATS_FMA_SEEDS=1" "-DCT_REINHARD_FMA_SEEDS=2" ...
for defs in "$@"; do
  CT_NVCC_DEFS="$defs" python color-transfer_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $defs"; continue; }
  python bench.py --linear-only 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$defs', 'reinhard %.1f Gpix/s %.3f  mkl %.1f %.3f' % (d['reinhard_f32']['Mpix/s'] / 1e3, d['reinhard_f32']['frac_of_hbm'], d['mkl_f32_to_f64']['Mpix/s'] / 1e3, d['mkl_f32_to_f64']['frac_of_hbm']))"
done
python color-transfer_b200/build.py --force > /dev/null 2>&1
