#!/bin/bash
# element-kernel launch-shape variants (registers / CTA size) at fem2d L=8
for v in "" mb6 t64 t64mb12 t256; do
  if [ -z "$v" ]; then unset MGB_B200_LIB; else export MGB_B200_LIB=/root/repo/scratch/variants/libmgb_$v.so; fi
  echo "variant ${v:-default}: $(python scratch/te_time.py 8 30 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read())['times']; print(d['full'], d['f0']['total_us'])")"
done
