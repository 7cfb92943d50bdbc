#!/bin/bash
run() {
  env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['kernels_ms']
print('$*', 'value %.0f e2e %.0f conv %.2f xill %.2f line %.2f' % (d['value'], d['e2e']['value'], k['k_conv'], k['k_xill'], k['k_line']))"
}
run RELXILL_B200_CONV_MINB=2
run RELXILL_B200_CONV_MINB=1
