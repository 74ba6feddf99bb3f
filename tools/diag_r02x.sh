#!/bin/bash
# round 2, GPU call 24: the recurrence compare every 8 trips (CHAOS_COMPARE_EVERY) against group ends only (32) and every 4, all workloads;
# shortcut / engine-agreement / full-list parity tests of the new build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python tools/quick.py --settings "X=0 CHAOS_KERNELS_DIR=tools/variants/ce32 CHAOS_KERNELS_DIR=tools/variants/ce4" --workloads "c2 c2f32 c2ex2 c4 c5 c1" --steps 10 ) > gpurun_out/r02x_quick.txt 2>&1
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "shortcuts or engines_agree or full_lists or full_size_frame" 2>&1 | tail -8 ) > gpurun_out/r02x_pytest.log
cat gpurun_out/r02x_quick.txt gpurun_out/r02x_pytest.log
