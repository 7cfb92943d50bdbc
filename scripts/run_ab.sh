#!/bin/bash
# usage: scripts/run_ab.sh VAR v1 v2 ...   quick bench (no CPU baseline, no extras) once per value of an environment switch
VAR=$1; shift
mkdir -p gpurun_out
for V in "$@"; do
  env $VAR=$V RELXILL_B200_TIMING=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/ab_$V.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$VAR=$V value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['kernels_ms'])"
done
