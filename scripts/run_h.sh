#!/bin/bash
# zone spectra on the convolution grid (default) against the table grid: GPU tests, then both variants of the bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
run() {
  env "$@" timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_$1.err | tee gpurun_out/bench_$1.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['kernels_ms']
print('$*', 'value %.0f ms/step %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']), k, d['roofline']['frac'], d['hbm_stage'].get('frac'))"
}
run RELXILL_B200_XILL_GRID=conv
run RELXILL_B200_XILL_GRID=table
run RELXILL_B200_XILL_GRID=conv
run RELXILL_B200_XILL_GRID=table
