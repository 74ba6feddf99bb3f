#!/bin/bash
# GPU call 62: final verification of the build as committed: whole GPU suite, smoke, both bench arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r04i_pytest.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/r04i_smoke.log 2>&1
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r04i_bench_ref.json 2> gpurun_out/r04i_bench_ref.err
( time timeout 900 python bench.py ) > gpurun_out/r04i_bench.json 2> gpurun_out/r04i_bench.err
cat gpurun_out/r04i_pytest.log; tail -n 2 gpurun_out/r04i_smoke.log
