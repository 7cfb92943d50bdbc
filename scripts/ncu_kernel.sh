#!/bin/bash
# usage: scripts/ncu_kernel.sh <kernel-regex> <tag> [batch]
K=$1; TAG=$2; B=${3:-1024}
ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch $B > gpurun_out/ncu_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}.log | cut -c1-200
