#!/bin/bash
# objective-only element instances: CTAs per SM via launch bounds (5 = as before, 6, 7 = default now, 8)
t() { python scratch/te_time.py 8 40 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read())['times']; print('full', round(d['full']['total_us'],2), 'f0', d['f0'])"; }
for v in f5 f6 f8; do echo "$v   $(MGB_B200_LIB=/root/repo/scratch/variants/libmgb_$v.so t)"; done
echo "f7   $(t)"
echo "coarse level 4, f7: $(python scratch/level_split.py 8 2>/dev/null | sed -n 5p)"
echo "coarse level 4, f8: $(MGB_B200_LIB=/root/repo/scratch/variants/libmgb_f8.so python scratch/level_split.py 8 2>/dev/null | sed -n 5p)"
