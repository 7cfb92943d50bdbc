#!/bin/bash
# ncu --set full of one kernel of the relxilllpCp (6-D table) bench
mkdir -p gpurun_out
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/profcp_${K} \
    python bench.py --model relxilllpCp --steps 1 --warmup 1 --no-cpu-baseline --batch 2048 > gpurun_out/ncucp_${K}.log 2>&1
  tail -1 gpurun_out/ncucp_${K}.log | cut -c1-200
done
