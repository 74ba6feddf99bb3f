#!/bin/bash
# GPU call 52 (8 GPUs): the default bench command at N = 8, both arms (the reference arm runs on rank 0 alone)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 ) > gpurun_out/r03z_bench_8gpu.json 2> gpurun_out/r03z_bench_8gpu.err
tail -n 4 gpurun_out/r03z_bench_8gpu.err; head -c 400 gpurun_out/r03z_bench_8gpu.json
