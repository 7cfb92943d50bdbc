#!/bin/bash
# round deliverables: parity tests, full bench line, ncu launch list, one --set full capture of the three big kernels
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_line|k_xill|k_conv' -s 3 -c 3 -o gpurun_out/prof_big3 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_big3.log 2>&1
tail -2 gpurun_out/ncu_big3.log | cut -c1-200
python - <<'PY'
import json
for f in ('bench_1gpu',):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        print(f, 'value %.0f e2e %.0f n_gpus %d ms/step %.1f launches %d clocks %s' % (d['value'], d['e2e']['value'], d['n_gpus'], d['ms_per_step'], d['gpu_launches'], d['clocks']))
        print('   kernels', d['kernels_ms']); print('   roofline', d['roofline']); print('   cpu', d['cpu_baseline'])
    except Exception as e:
        print(f, 'failed', e); print(open('gpurun_out/%s.err'%f).read()[-1500:])
PY
