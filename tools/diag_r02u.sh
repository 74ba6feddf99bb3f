#!/bin/bash
# round 2, GPU call 21 (8 GPUs): the bench lines of the final build at N = 8, 4, 2 as the driver launches them
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 8 4 2; do
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n ) > gpurun_out/r02u_bench_${n}gpu.json 2> gpurun_out/r02u_bench_${n}gpu.err
done
echo done
