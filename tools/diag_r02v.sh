#!/bin/bash
# round 2, GPU call 22: final verification of the build as committed: whole GPU suite, smoke, both bench arms, racecheck
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/r02v_pytest.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/r02v_smoke.log 2>&1
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r02v_bench_ref.json 2> gpurun_out/r02v_bench_ref.err
( time timeout 900 python bench.py ) > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err
( timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py 2>&1 | tail -25 ) > gpurun_out/r02v_racecheck.log
echo done
