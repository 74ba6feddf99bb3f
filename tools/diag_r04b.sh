#!/bin/bash
# GPU call 54: the shared-memory cap of the long kernels against the early compose into host memory (which needs its palette's shared memory next to them): end to end
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SETTINGS="CHAOS_LONG_SMEM=0 CHAOS_LONG_SMEM=73728 CHAOS_LONG_SMEM=58368 CHAOS_LONG_SMEM=58368+CHAOS_HOST_COMPOSE_BLOCKS=296" WORKLOADS="c2 c2f32" STEPS=20 tools/sweep_env.sh 2>&1 | tee gpurun_out/r04b.txt
