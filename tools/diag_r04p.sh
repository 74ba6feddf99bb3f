#!/bin/bash
# GPU call 71 (2 GPUs): the default bench command at N = 2 on the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 ) > gpurun_out/r04p_bench_2gpu.json 2> gpurun_out/r04p_bench_2gpu.err
tail -n 3 gpurun_out/r04p_bench_2gpu.err; head -c 200 gpurun_out/r04p_bench_2gpu.json
