#!/bin/bash
# round 2, GPU call 12 (2 GPUs): cross-GPU tile stealing: tests, then the N = 2 bench line (c4 block: with / without stealing)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r02l_pytest.log
( timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_api_gpu.py -m gpu -x -q 2>&1 | tail -5 ) >> gpurun_out/r02l_pytest.log
for s in 1 0; do
( CHAOS_STEAL=$s timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$s bench.py --gpus 2 --workload c4 --steps 10 --no-extras ) > gpurun_out/r02l_c4_2gpu_steal$s.json 2> gpurun_out/r02l_c4_2gpu_steal$s.err
done
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 ) > gpurun_out/r02l_bench_2gpu.json 2> gpurun_out/r02l_bench_2gpu.err
( timeout 300 python bench.py --workload c4 --steps 10 --no-extras --no-cpu-baseline --no-full-trips ) > gpurun_out/r02l_c4_1gpu.json 2> gpurun_out/r02l_c4_1gpu.err
echo done
