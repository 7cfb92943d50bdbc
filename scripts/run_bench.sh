#!/bin/bash
# usage: scripts/run_bench.sh [extra bench.py args]  — the default bench line on one GPU, wall time of the whole run noted
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python bench.py "$@" > gpurun_out/bench_1gpu.log 2> gpurun_out/bench_1gpu.err
echo "rc=$? wall=$(( $(date +%s) - t0 )) s"
tail -1 gpurun_out/bench_1gpu.log | cut -c1-6000
tail -5 gpurun_out/bench_1gpu.err
