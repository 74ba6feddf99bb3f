#!/bin/bash
# GPU call 32: timeline of a one-strand c2 frame with the marathon launches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/quick.py --settings "CHAOS_STRANDS=1+CHAOS_TIMELINE=gpurun_out/r03f_tl.txt CHAOS_STRANDS=1+CHAOS_EVICT_TRIPS=4096+CHAOS_TIMELINE=gpurun_out/r03f_tl4096.txt" --workloads "c2" --steps 4 2>&1 | tee gpurun_out/r03f_quick.txt
cat gpurun_out/r03f_tl.txt; echo; cat gpurun_out/r03f_tl4096.txt
