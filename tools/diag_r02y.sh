#!/bin/bash
# round 2, GPU call 25: timeline of a c2 frame's launches (CHAOS_TIMELINE), two strands and one
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/quick.py --settings "CHAOS_TIMELINE=gpurun_out/r02y_tl2.txt CHAOS_TIMELINE=gpurun_out/r02y_tl1.txt+CHAOS_STRANDS=1 CHAOS_TIMELINE=gpurun_out/r02y_tl2_ce32.txt+CHAOS_KERNELS_DIR=tools/variants/ce32" --workloads "c2" --steps 5 2>&1 | tee gpurun_out/r02y_quick.txt
for f in tl2 tl1 tl2_ce32; do echo "== $f"; cat gpurun_out/r02y_$f.txt; done
