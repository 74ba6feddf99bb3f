#!/bin/bash
# GPU call 47: the build as it stands (compare every 16 / 8 in FP64 / FP32): all workloads, the whole GPU suite, smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/quick.py --settings "X=0" --workloads "c2 c2f32 c2ex2 c4 c5 c1" --steps 10 2>&1 | tee gpurun_out/r03u_quick.txt
( time timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/r03u_pytest.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) 2>&1 | tee gpurun_out/r03u_smoke.log
