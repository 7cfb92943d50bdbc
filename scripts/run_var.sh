#!/bin/bash
# experiment: twiddle source of k_conv x lane split of k_xill
run() {
  env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['kernels_ms']
print('$*', 'value %.0f e2e %.0f conv %.2f xill %.2f line %.2f' % (d['value'], d['e2e']['value'], k['k_conv'], k['k_xill'], k['k_line']), d['clocks'])"
}
run RELXILL_B200_CONV_TW=0 RELXILL_B200_XILL_SPLIT=1
run RELXILL_B200_CONV_TW=1 RELXILL_B200_XILL_SPLIT=2
run RELXILL_B200_CONV_TW=2 RELXILL_B200_XILL_SPLIT=1
run RELXILL_B200_CONV_TW=3 RELXILL_B200_XILL_SPLIT=2
RELXILL_B200_CONV_TW=3 RELXILL_B200_XILL_SPLIT=2 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
RELXILL_B200_CONV_TW=2 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
