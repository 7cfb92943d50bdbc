#!/bin/bash
# GPU tests + one bench line of the metric model (and optionally the Cp model)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for M in relxilllp "$@"; do
timeout 900 python bench.py --model $M --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_q_$M.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['kernels_ms']
print('$M value %.0f ms/step %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']), k)"
done
