for v in default mb4 mb6 t256mb2 t256mb3 t64mb10 gu2 gu6 gu8; do
  if [ $v = default ]; then unset MGB_B200_LIB; else export MGB_B200_LIB=$PWD/scratch/variants/libmgb_$v.so; fi
  python bench.py --steps 40 --warmup 5 --cpu-reps 0 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v','value',round(d['value']*1e3,2),'elem',round(d['roofline']['kernel_ms']*1e3,2),'gather',round(d['roofline']['assembly']['gather_ms']*1e3,2))"
done
