#!/bin/bash
# GPU call 33: lengths guessed from the convergence rate (CHAOS_PREDICT), top / hot / rest classes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "full_lists or shortcuts or engines_agree or exported or strands or full_size_frame" 2>&1 | tail -8 ) > gpurun_out/r03g_pytest.log
cat gpurun_out/r03g_pytest.log
timeout 600 python tools/quick.py --settings "X=0 CHAOS_PREDICT=0 CHAOS_STRANDS=1 CHAOS_STRANDS=1+CHAOS_PREDICT=0 CHAOS_HOT_FIRST=0 CHAOS_KERNELS_DIR=tools/variants/ce32 CHAOS_ENGINE=2" --workloads "c2 c2f32 c2ex2" --steps 8 2>&1 | tee gpurun_out/r03g_quick.txt
( CHAOS_STRANDS=1 LS_FRAMES=2 timeout 300 python tools/lane_stats.py c2 2>&1 | grep -v "pass main" | tail -6 ) 2>&1 | tee gpurun_out/r03g_ls.txt
