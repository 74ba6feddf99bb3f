#!/bin/bash
# GPU call 59: pass D colours the tiles it finishes (CHAOS_FUSE_REPLAY), no second compose launch: device and end to end
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SETTINGS="X=0 CHAOS_FUSE_REPLAY=0 X=1 CHAOS_FUSE_REPLAY=0+X=1" WORKLOADS="c2 c2ex2" STEPS=20 tools/sweep_env.sh 2>&1 | tee gpurun_out/r04f.txt
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_api_gpu.py -m gpu -q -x 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r04f_pytest.log
