#!/bin/bash
# A/B of an environment switch on the two walker benches: scripts/run_n.sh VAR v1 v2 ...
mkdir -p gpurun_out
VAR=$1; shift
for V in "$@"; do
  for M in relxilllp relxilllpCp; do
  env $VAR=$V timeout 900 python bench.py --model $M --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['kernels_ms']
print('$VAR=$V $M value %.0f ms/step %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']), k)"
  done
done
