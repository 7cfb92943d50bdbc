#!/bin/bash
run() {
  env "$@" python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['kernels_ms']
print('$*', 'value %.0f ms/step %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
}
run RELXILL_B200_NO_AUX=1
run RELXILL_B200_X=1
run RELXILL_B200_NO_AUX=1
run RELXILL_B200_X=1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
