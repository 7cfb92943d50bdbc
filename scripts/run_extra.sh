#!/bin/bash
# the smaller deliverables: full-batch parity census, launch list of the headline workload only, single-call latencies
mkdir -p gpurun_out
rm -f gpurun_out/census.jsonl
RELXILL_B200_CENSUS_OUT=gpurun_out/census.jsonl timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "every_row or shard_vs_reference or full_sweep" 2>&1 | tail -2
cat gpurun_out/census.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/launches.log 2>&1
python scripts/latency_probe.py > gpurun_out/latency.json 2>/dev/null; cat gpurun_out/latency.json
