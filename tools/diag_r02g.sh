#!/bin/bash
# round 2, GPU call 7: full GPU suite, the default bench line (N = 1, everything in it) and the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r02g_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r02g_bench_ref.json 2> gpurun_out/r02g_bench_ref.err
SETTINGS="X=0 CHAOS_ENGINE=1 CHAOS_ENGINE=2" WORKLOADS="c2 c2ex2 c4 c5 c1 c2f32" STEPS=10 tools/sweep_env.sh > gpurun_out/r02g_engines.txt 2>&1
echo done
