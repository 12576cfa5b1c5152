for lib in default a34m2 a33m3 a34m3 a44m4; do
  if [ "$lib" = default ]; then unset CT_B200_LIB; else export CT_B200_LIB=$PWD/tools/scratch/libct_$lib.so; fi
  python bench.py --linear-only 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', {k:(round(v['Mpix/s']), round(v['frac_of_hbm'],3)) for k,v in d.items() if 'Mpix/s' in v})"
done
