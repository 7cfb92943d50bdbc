#!/bin/bash
# the 6-D (Cp) workloads in both zone-spectrum layouts: metric model relxilllpCp (50 zones, walkers) and the config-4 probe
mkdir -p gpurun_out
for G in conv table; do
  RELXILL_B200_XILL_GRID=$G timeout 900 python bench.py --model relxilllpCp --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/benchcp_$G.err | tee gpurun_out/benchcp_$G.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['kernels_ms']
print('$G relxilllpCp value %.0f ms/step %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']), k)"
  RELXILL_B200_XILL_GRID=$G timeout 900 python scripts/cfg4_probe.py 8192 2>gpurun_out/cfg4_$G.err | tee gpurun_out/cfg4_$G.json | cut -c1-600
done
