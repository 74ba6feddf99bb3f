#!/bin/bash
# experiment: trips per untested block (CHAOS_BLOCK_ITERS) on the headline workloads
cd "$(dirname "$0")/.."
for nb in 64 96 128 192 256; do
  for w in c2 c2ex2 c4; do
    CHAOS_BLOCK_ITERS=$nb timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-full-trips > gpurun_out/nb_${nb}_$w.json 2> gpurun_out/nb_${nb}_$w.err
  done
done
