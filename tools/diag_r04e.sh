#!/bin/bash
# GPU call 57: stalled orbits finished by the long kernel's own warps on their way out (CHAOS_FINISH_IN_LONG), against the finish kernel alone
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/quick.py --settings "X=0 CHAOS_FINISH_IN_LONG=0 X=1 CHAOS_FINISH_IN_LONG=0+X=1 CHAOS_STRANDS=1 CHAOS_STRANDS=1+CHAOS_FINISH_IN_LONG=0 CHAOS_TIMELINE=gpurun_out/r04e_tl.txt" --workloads "c2 c2f32" --steps 10 2>&1 | tee gpurun_out/r04e_quick.txt
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "full_lists or shortcuts or engines_agree or exported or strands or full_size_frame" 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r04e_pytest.log
cat gpurun_out/r04e_tl.txt
