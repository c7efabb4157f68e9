#!/bin/bash
# fine-level u-u block on the tensor cores: default library with the switch on / off, and static variants
python -m pytest tests/test_parity_gpu.py tests/test_golden.py -q -m gpu -x -k "not fem3d" 2>&1 | tail -2
t() { python scratch/te_time.py 8 40 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read())['times']; print(d['full'], d['f0']['total_us'])"; }
echo "on      $(t)"
echo "off     $(MGB_MMA_FINE=0 t)"
for v in st st6 st7; do echo "$v     $(MGB_B200_LIB=/root/repo/scratch/variants/libmgb_$v.so t)"; done
echo "on      $(t)"
