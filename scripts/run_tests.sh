#!/bin/bash
# GPU parity tests (with the outlier census written to gpurun_out/) + one bench line
mkdir -p gpurun_out
rm -f gpurun_out/census.jsonl
RELXILL_B200_CENSUS_OUT=gpurun_out/census.jsonl timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -30
cat gpurun_out/census.jsonl
RELXILL_B200_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench_quick.log | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['kernels_ms'], d['clocks'])"
grep "timing" gpurun_out/bench.err | tail -3
tail -5 gpurun_out/bench.err
