#!/bin/bash
mkdir -p gpurun_out
for G in conv table; do
  RELXILL_B200_XILL_GRID=$G timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['kernels_ms']
print('$G relxilllp value %.0f ms/step %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']), k)"
done
bash scripts/run_k.sh
