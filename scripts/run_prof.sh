#!/bin/bash
# usage: scripts/run_prof.sh <tag> <model> <kernel-regex> [<kernel-regex> ...]
# one `ncu --set full --import-source on` capture per kernel of a bench step of $BATCH vectors (default 1024); source pages
# exported as CSV
TAG=$1; MODEL=$2; shift 2
mkdir -p gpurun_out
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_${K} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --batch ${BATCH:-1024} --model $MODEL > gpurun_out/ncu_${TAG}_${K}.log 2>&1
  tail -2 gpurun_out/ncu_${TAG}_${K}.log | cut -c1-200
  ncu -i gpurun_out/prof_${TAG}_${K}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src_${TAG}_${K}.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_${K}.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}_${K}.csv 2>/dev/null
  rm -f gpurun_out/prof_${TAG}_${K}.ncu-rep
done
