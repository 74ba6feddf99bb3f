#!/bin/bash
# round 2, GPU call 15 (8 GPUs): the N = 8 and N = 4 bench lines as the driver launches them
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02o_gpus.txt
nvidia-smi topo -m > gpurun_out/r02o_topo.txt 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 ) > gpurun_out/r02o_bench_8gpu.json 2> gpurun_out/r02o_bench_8gpu.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 ) > gpurun_out/r02o_bench_4gpu.json 2> gpurun_out/r02o_bench_4gpu.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --impl reference --no-extras ) > gpurun_out/r02o_bench_ref_8gpu.json 2> gpurun_out/r02o_bench_ref_8gpu.err
echo done
