#!/bin/bash
# round 2, GPU call 8 (2 GPUs): multi-process tests on two real devices, the N = 2 bench line of both arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02h_gpus.txt
( timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r02h_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 ) > gpurun_out/r02h_bench_2gpu.json 2> gpurun_out/r02h_bench_2gpu.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference ) > gpurun_out/r02h_bench_ref_2gpu.json 2> gpurun_out/r02h_bench_ref_2gpu.err
( timeout 300 python bench.py --workload c1 --steps 10 --no-extras --no-cpu-baseline --no-full-trips ) > gpurun_out/r02h_c1.json 2> gpurun_out/r02h_c1.err
( timeout 300 python bench.py --workload c3 --no-cpu-baseline ) > gpurun_out/r02h_c3.json 2> gpurun_out/r02h_c3.err
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload c3 ) > gpurun_out/r02h_c3_2gpu.json 2> gpurun_out/r02h_c3_2gpu.err
echo done
