#!/bin/bash
for P in 1000000 2072 1480; do
  RELXILL_B200_PIPE=$P python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('pipe $P value %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"
done
