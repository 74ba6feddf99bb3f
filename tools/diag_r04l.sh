#!/bin/bash
# GPU call 66: parts per one-launch frame on its way to host memory (CHAOS_HOST_PARTS), now that the composes run next to the renders
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SETTINGS="CHAOS_HOST_PARTS=1 CHAOS_HOST_PARTS=2 CHAOS_HOST_PARTS=3 CHAOS_HOST_PARTS=4 CHAOS_HOST_PARTS=8" WORKLOADS="c5 c4" STEPS=10 tools/sweep_env.sh 2>&1 | tee gpurun_out/r04l.txt
