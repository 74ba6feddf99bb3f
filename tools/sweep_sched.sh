#!/bin/bash
# experiment: how many finished/empty lanes a warp lets accumulate before a scheduling pass (CHAOS_SCHED_IDLE=indep,rounds)
cd "$(dirname "$0")/.."
for si in 1,1 3,6 3,10 3,16 6,6 2,4 8,24; do
  for w in c2 c2ex2 c4 c1; do
    CHAOS_SCHED_IDLE=$si timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-full-trips > gpurun_out/si_${si}_$w.json 2> gpurun_out/si_${si}_$w.err
  done
done
