#!/bin/bash
# piece sizes of the pipelined host-buffer call (RELXILL_B200_PIPE / _PIPE_LAST; the convolution of a relxill batch is
# cut into pieces of half these sizes) against the end-to-end rate
run() {
  env "$@" python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$*', 'value %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"
}
run RELXILL_B200_PIPE=2072 RELXILL_B200_PIPE_LAST=888
run RELXILL_B200_PIPE=1184 RELXILL_B200_PIPE_LAST=592
run RELXILL_B200_PIPE=2368 RELXILL_B200_PIPE_LAST=592
run RELXILL_B200_PIPE=2368 RELXILL_B200_PIPE_LAST=1184
run RELXILL_B200_PIPE=1776 RELXILL_B200_PIPE_LAST=296
