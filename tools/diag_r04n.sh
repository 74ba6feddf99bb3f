#!/bin/bash
# GPU call 68: the exported tiles of a strand composed when its own pass D is over (CHAOS_LATE_BY_STRAND): device / end to end; parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SETTINGS="CHAOS_LATE_BY_STRAND=1 CHAOS_LATE_BY_STRAND=0 X=1 CHAOS_LATE_BY_STRAND=0+X=1" WORKLOADS="c2 c2ex2 c2f32" STEPS=20 tools/sweep_env.sh 2>&1 | tee gpurun_out/r04n.txt
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_api_gpu.py -m gpu -q -x -k "strands or exported or full_size_frame or engines_agree or visualise or device_output or partition" 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r04n_pytest.log
