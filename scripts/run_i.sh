#!/bin/bash
# ncu --set full of one kernel for both zone-spectrum layouts
mkdir -p gpurun_out
K=${1:-k_xill}
for G in conv table; do
  RELXILL_B200_XILL_GRID=$G ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${K}_$G \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 2048 > gpurun_out/ncu_${K}_$G.log 2>&1
  tail -2 gpurun_out/ncu_${K}_$G.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
