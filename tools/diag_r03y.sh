#!/bin/bash
# GPU call 51: long-kernel CTAs capped at 3 per SM across the strands' launches (default now), against off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/quick.py --settings "X=0 CHAOS_LONG_SMEM=0 X=1 CHAOS_LONG_SMEM=0+X=1" --workloads "c2 c2f32" --steps 10 2>&1 | tee gpurun_out/r03y_quick.txt
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "strands or exported or full_size_frame or engines_agree" 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r03y_pytest.log
