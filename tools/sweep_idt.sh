#!/bin/bash
# Tuning aid: rebuild the library on the GPU box with different -D settings and time the IDT step.
# usage: tools/sweep_idt.sh "-DCT_RANGES_CHUNK=2" "-DCT_HIST_STAGES=5" ...
for defs in "$@"; do
  CT_NVCC_DEFS="$defs" python color-transfer_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $defs"; continue; }
  python bench.py --steps 40 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$defs', '%.0f Mpix/s step %.3f' % (d['value'], d['roofline']['step_frac']), d['roofline']['kernel_ms'], d['clocks']['sm_mhz'])"
done
python color-transfer_b200/build.py --force > /dev/null 2>&1
