#!/bin/bash
# round 2, GPU call 16 (8 GPUs): what costs a rank of an 8-GPU c2 frame 1.4 ms when the same bands take 0.93 ms alone
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
i=0
for set in "X=0" "CHAOS_STRAND_MIN_TILES=0" "CHAOS_LONG_OCC=0,0,0" "CHAOS_BENCH_BAND_ROWS=16"; do
  i=$((i+1))
  ( env $set timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953$i bench.py --gpus 8 --no-extras --steps 40 ) > gpurun_out/r02p_8gpu_$i.json 2> gpurun_out/r02p_8gpu_$i.err
done
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --no-extras --steps 40 ) > gpurun_out/r02p_4gpu.json 2> gpurun_out/r02p_4gpu.err
echo done
