#!/bin/bash
# round 2, GPU call 17: which rank of an 8-GPU c2 frame is the slow one (bands emulated one rank at a time), c2ex2 with the
# engine chosen from the previous frame, full suite, the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SETTINGS="CHAOS_EMULATE_PART=0:8 CHAOS_EMULATE_PART=1:8 CHAOS_EMULATE_PART=2:8 CHAOS_EMULATE_PART=3:8 CHAOS_EMULATE_PART=4:8 CHAOS_EMULATE_PART=5:8 CHAOS_EMULATE_PART=6:8 CHAOS_EMULATE_PART=7:8" WORKLOADS="c2" STEPS=10 tools/sweep_env.sh > gpurun_out/r02q_ranks.txt 2>&1
SETTINGS="X=0 CHAOS_STREAMS_ABOVE=0" WORKLOADS="c2ex2 c2 c2f32" STEPS=20 tools/sweep_env.sh > gpurun_out/r02q_heur.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/r02q_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r02q_bench_ref.json 2> gpurun_out/r02q_bench_ref.err
echo done
