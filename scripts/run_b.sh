#!/bin/bash
# quick GPU regression: parity tests + short bench; prints one summary line
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_tmp.json'))
    print('value %.0f  e2e %.0f  kernels %s' % (d['value'], d['e2e']['value'], d['kernels_ms']))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_tmp.err').read()[-2000:])
PY
