#!/bin/bash
# ncu --set full of the named kernels (one launch each), default layout
mkdir -p gpurun_out
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${K} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 2048 > gpurun_out/ncu_${K}.log 2>&1
  tail -1 gpurun_out/ncu_${K}.log | cut -c1-200
done
