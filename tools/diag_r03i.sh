#!/bin/bash
# GPU call 35: the drain of the long kernel by pool threshold, with the length guess
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/quick.py --settings "CHAOS_STRANDS=1 CHAOS_STRANDS=1+CHAOS_POOL_MIN=4 CHAOS_STRANDS=1+CHAOS_POOL_MIN=12 CHAOS_STRANDS=1+CHAOS_POOL_MIN=31 X=0 CHAOS_POOL_MIN=4 CHAOS_POOL_MIN=12 CHAOS_POOL_MIN=31" --workloads "c2" --steps 8 2>&1 | tee gpurun_out/r03i_quick.txt
