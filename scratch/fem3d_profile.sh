#!/bin/bash
# fem3d L=5 (CSR path): launch list + ncu --set full of the csr kernels
set -x
python scratch/fem3d_time.py 5 > gpurun_out/fem3d_plain_c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/fem3d_launches_r1c.csv python scratch/fem3d_time.py 5 > gpurun_out/fem3d_ncu_c0.log 2>&1
python scratch/fem3d_time.py 5 > gpurun_out/fem3d_plain_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'csr_' -s 8 -c 4 -o gpurun_out/prof_r1c_fem3d -f python scratch/fem3d_time.py 5 > gpurun_out/fem3d_ncu_c.log 2>&1
tail -1 gpurun_out/fem3d_plain_c.log
