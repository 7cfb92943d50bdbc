#!/bin/bash
# state of the tree on a B200: GPU parity tests, bench line, launch list, ncu full capture of the step's kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>gpurun_out/bench.err; tail -1 gpurun_out/bench.log | cut -c1-3000
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_" -s 8 -c 8 -o gpurun_out/prof_all \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_all.log 2>&1
tail -2 gpurun_out/ncu_all.log | cut -c1-200
