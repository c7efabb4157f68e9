#!/bin/bash
# ncu --set full of the fem3d (CSR path) kernels: one-stage replay and block-staged replay
set -x
for v in 0 512 128; do
  export MGB_HESS_BLOCK=$v
  python scratch/fem3d_time.py 5 > gpurun_out/fem3d_plain_$v.log 2>&1 || exit 1
  ncu --set full --clock-control none --import-source on -k regex:'csr_' -s 10 -c 6 -o gpurun_out/prof_fem3d_hb$v -f python scratch/fem3d_time.py 5 > gpurun_out/fem3d_ncu_$v.log 2>&1
done
ls -la gpurun_out | tail -5
