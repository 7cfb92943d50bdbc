#!/bin/bash
# ncu --set full capture of the three big kernels (second step of a short bench run)
ncu --set full --clock-control none --import-source on -k "regex:k_line|k_xill|k_conv|k_syspar|k_fine|k_dist" -s 7 -c 7 -o gpurun_out/prof_big3 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_big3.log 2>&1
tail -2 gpurun_out/ncu_big3.log | cut -c1-200
