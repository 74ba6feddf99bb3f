#!/bin/bash
# round 2, GPU call 23: where the tile-synchronous kernel spends c5 (julia 4K, maxIter 900, adaptive SS 8) and c1's single launch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --no-extras --no-cpu-baseline --no-full-trips"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fractalRenderMainDoubleSync" --launch-skip 3 --launch-count 1 -o gpurun_out/r02w_sync_c5 -f $B --workload c5 --steps 1 --warmup 3 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fractalRenderMainDouble" --launch-skip 3 --launch-count 1 -o gpurun_out/r02w_main_c1 -f $B --workload c1 --steps 1 --warmup 3 > /dev/null 2>&1
echo done
