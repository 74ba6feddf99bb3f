#!/bin/bash
# GPU call 37: recurrence compare every 8 trips switched by the previous frame's proven share (CHAOS_DENSE_COMPARE forces), survivors carry their state
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/quick.py --settings "X=0 CHAOS_DENSE_COMPARE=0 CHAOS_DENSE_COMPARE=1" --workloads "c2 c2f32 c2ex2 c4 c5 c1" --steps 8 2>&1 | tee gpurun_out/r03k_quick.txt
( time timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "not c4_full" 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/r03k_pytest.log
