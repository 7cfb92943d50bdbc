python scripts/gpu_check.py > gpurun_out/check2.log 2>&1; grep -E "final|==" gpurun_out/check2.log | awk '{print $1,$2,$3,$4}' | head -40
NZ=50 N=4 MODELS=relxilllp python scripts/gpu_check.py > gpurun_out/check3.log 2>&1; grep -E "final|==" gpurun_out/check3.log | cut -c1-60
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; python -c "
import json; d=json.load(open('gpurun_out/bench2.json')); print(d['value'], d['e2e']['value'], d['kernels_ms'])"
