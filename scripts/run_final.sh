#!/bin/bash
# round deliverables on one B200: parity tests, full bench line (both arms), launch list, ncu --set full of every kernel of
# one step, the 6-D (Cp) bench, the config-4 probe, the parity report
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/launches.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_' -s 9 -c 9 -f -o gpurun_out/prof_step \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_step.log 2>&1
tail -1 gpurun_out/ncu_step.log | cut -c1-200
python bench.py --model relxilllpCp --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cp.json 2> gpurun_out/bench_cp.err
python scripts/cfg4_probe.py 8192 > gpurun_out/cfg4.json 2> gpurun_out/cfg4.err
python scripts/parity_report.py > gpurun_out/parity.json 2> gpurun_out/parity.err
python - <<'PY'
import json
for f in ('bench_1gpu', 'bench_ref', 'bench_cp'):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        print(f, 'value %.0f e2e %s ms/step %.2f launches %s clocks %s' % (d['value'], d.get('e2e',{}).get('value'), d['ms_per_step'], d.get('gpu_launches'), d.get('clocks')))
        print('   kernels', d.get('kernels_ms')); print('   roofline', d.get('roofline')); print('   cpu', d.get('cpu_baseline'))
    except Exception as e:
        print(f, 'failed', e); print(open('gpurun_out/%s.err'%f).read()[-1500:])
print(open('gpurun_out/parity.json').read()[:600])
print(open('gpurun_out/cfg4.json').read()[:900])
PY
