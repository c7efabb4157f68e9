#!/bin/bash
# round-1 session-3 measurement pass (one GPU): tests, headline bench, per-config times, launch list, fem3d ncu
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err && tail -c 900 gpurun_out/bench_r1c.json
python scratch/config_times.py > gpurun_out/config_times_r1c.log 2>&1; tail -3 gpurun_out/config_times_r1c.log
python bench.py --steps 5 --warmup 3 --cpu-reps 0 > gpurun_out/plain_c1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 5 --warmup 3 --cpu-reps 0 > gpurun_out/ncu_c1.log 2>&1
python scratch/fem3d_time.py 5 > gpurun_out/fem3d_plain_c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/fem3d_launches_r1c.csv python scratch/fem3d_time.py 5 > gpurun_out/fem3d_ncu_c0.log 2>&1
python scratch/fem3d_time.py 5 > gpurun_out/fem3d_plain_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'csr_' -s 14 -c 7 -o gpurun_out/prof_r1c_fem3d -f python scratch/fem3d_time.py 5 > gpurun_out/fem3d_ncu_c.log 2>&1
ls -la gpurun_out | tail -6
