#!/bin/bash
# GPU call 49 (2 GPUs): the multi-process tests on two devices and the default bench command at N = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -5 ) > gpurun_out/r03w_pytest.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 ) > gpurun_out/r03w_bench_2gpu.json 2> gpurun_out/r03w_bench_2gpu.err
cat gpurun_out/r03w_pytest.log; tail -n 4 gpurun_out/r03w_bench_2gpu.err; head -c 600 gpurun_out/r03w_bench_2gpu.json
