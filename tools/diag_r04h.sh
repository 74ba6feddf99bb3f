#!/bin/bash
# GPU call 61 (4 GPUs): the default bench command at N = 4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 ) > gpurun_out/r04h_bench_4gpu.json 2> gpurun_out/r04h_bench_4gpu.err
tail -n 4 gpurun_out/r04h_bench_4gpu.err; head -c 300 gpurun_out/r04h_bench_4gpu.json
