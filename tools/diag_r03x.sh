#!/bin/bash
# GPU call 50: at most 3 / 2 long-kernel CTAs per SM across the strands' launches (CHAOS_LONG_SMEM), so that the chains' other kernels find a CTA slot
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/quick.py --settings "X=0 CHAOS_LONG_SMEM=73728 CHAOS_LONG_SMEM=73728+CHAOS_STRANDS=3 CHAOS_LONG_SMEM=73728+CHAOS_STRANDS=4 CHAOS_LONG_SMEM=110000 CHAOS_LONG_SMEM=110000+CHAOS_STRANDS=3 CHAOS_LONG_SMEM=73728+CHAOS_TIMELINE=gpurun_out/r03x_tl.txt" --workloads "c2" --steps 8 2>&1 | tee gpurun_out/r03x_quick.txt
cat gpurun_out/r03x_tl.txt
