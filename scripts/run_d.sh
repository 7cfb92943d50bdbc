#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python - <<'PY'
import json
for f in ('bench_1gpu','bench_2gpu'):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        print(f, 'value %.0f e2e %.0f n_gpus %d ms/step %.1f launches %d clocks %s' % (d['value'], d['e2e']['value'], d['n_gpus'], d['ms_per_step'], d['gpu_launches'], d['clocks']))
        print('   kernels', d['kernels_ms']); print('   roofline', d['roofline']); print('   hbm_stage', d['hbm_stage']); print('   cpu', d['cpu_baseline'])
    except Exception as e:
        print(f, 'failed', e); print(open('gpurun_out/%s.err'%f).read()[-1500:])
PY
