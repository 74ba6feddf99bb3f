#!/bin/bash
# experiment: scheduling-pass thresholds (CHAOS_SCHED_IDLE=indep,rounds) and block length (CHAOS_BLOCK_ITERS)
cd "$(dirname "$0")/.."
for nb in 64 128; do
for si in 1,8 2,8 4,8 6,8 6,4 6,16; do
  for w in c2 c2ex2; do
    CHAOS_BLOCK_ITERS=$nb CHAOS_SCHED_IDLE=$si timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-full-trips > gpurun_out/ss_${nb}_${si}_$w.json 2> /dev/null
  done
done
done
