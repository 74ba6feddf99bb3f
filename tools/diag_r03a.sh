#!/bin/bash
# GPU call 27: survivors carry their state through the long list (no restart); compare every 8 / 16 / 32 on top of it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python tools/quick.py --settings "X=0 CHAOS_KERNELS_DIR=tools/variants/ce32 CHAOS_KERNELS_DIR=tools/variants/ce16 CHAOS_ENGINE=2" --workloads "c2 c2f32 c2ex2 c4 c5 c1" --steps 8 ) > gpurun_out/r03a_quick.txt 2>&1
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "shortcuts or engines_agree or full_lists or exported or strands or full_size_frame" 2>&1 | tail -8 ) > gpurun_out/r03a_pytest.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) > gpurun_out/r03a_smoke.log
cat gpurun_out/r03a_quick.txt gpurun_out/r03a_pytest.log gpurun_out/r03a_smoke.log
