#!/bin/bash
# usage: scripts/run_multi.sh N   — bench.py on N GPUs of one box, the way the driver launches it
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err
tail -1 gpurun_out/bench_${N}gpu.log | cut -c1-1500
tail -3 gpurun_out/bench_${N}gpu.err
