#!/bin/bash
# round 2, GPU call 13: full suite after the sync-kernel changes and stealing (3 ranks on one device), c5 / default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -60 ) > gpurun_out/r02m_pytest.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/r02m_smoke.log 2>&1
SETTINGS="X=0" WORKLOADS="c5 c1 c2 c4" STEPS=20 tools/sweep_env.sh > gpurun_out/r02m_low.txt 2>&1
( time timeout 900 python bench.py ) > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
echo done
