for lib in default v1 v2; do
  if [ "$lib" = default ]; then unset CT_B200_LIB; else export CT_B200_LIB=$PWD/tools/scratch/libct_$lib.so; fi
  python bench.py --no-cpu-baseline --no-extras --steps 5 --e2e-frames 1 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', round(d['value']), d['roofline']['kernel_ms'])"
done
