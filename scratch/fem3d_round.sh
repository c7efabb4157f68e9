#!/bin/bash
set -x
python -m pytest tests/test_parity_gpu.py tests/test_parabolic_gpu.py tests/test_solver_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 500 python scratch/fem3d_tune.py 5 2>&1 | tail -12
python scratch/fem3d_time.py 5 > gpurun_out/fem3d_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/fem3d_launches3.csv python scratch/fem3d_time.py 5 > gpurun_out/fem3d_ncu.log 2>&1
tail -2 gpurun_out/fem3d_plain.log
