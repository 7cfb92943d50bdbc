#!/bin/bash
# GPU parity tests + one bench line (no CPU baseline) + the phase timings of the host-buffer call
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
RELXILL_B200_TIMING=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_quick.log | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['kernels_ms'], d['clocks'])"
grep "timing" gpurun_out/bench.err | tail -3
