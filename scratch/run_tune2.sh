for L in 8 10 7; do
for v in default mb7 mb8; do
  if [ $v = default ]; then unset MGB_B200_LIB; else export MGB_B200_LIB=$PWD/scratch/variants/libmgb_$v.so; fi
  python bench.py --steps 40 --warmup 5 --cpu-reps 0 --L $L 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('L$L $v','value',round(d['value']*1e3,2),'elem',round(d['roofline']['kernel_ms']*1e3,2),'gather',round(d['roofline']['assembly']['gather_ms']*1e3,2),'f0',round(d['roofline']['assembly']['f0_ms']*1e3,2))"
done; done
