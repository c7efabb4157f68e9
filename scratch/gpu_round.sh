#!/bin/bash
# one-GPU measurement pass: tests, bench (L=8 headline, L=10), ncu launch list + full capture
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err && tail -c 600 gpurun_out/bench_r1b.json
python bench.py --steps 30 --warmup 5 --L 10 --cpu-reps 0 > gpurun_out/bench_r1b_L10.json 2>> gpurun_out/bench_r1b.err
python bench.py --steps 5 --warmup 3 --cpu-reps 0 > gpurun_out/plain7.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 5 --warmup 3 --cpu-reps 0 > gpurun_out/ncu7.log 2>&1
python bench.py --steps 5 --warmup 3 --cpu-reps 0 > gpurun_out/plain8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'element_kernel|gather_kernel' -s 6 -c 2 -o gpurun_out/prof_r1b -f python bench.py --steps 5 --warmup 3 --cpu-reps 0 > gpurun_out/ncu8.log 2>&1
ls -la gpurun_out | tail -8
