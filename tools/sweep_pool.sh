#!/bin/bash
# experiment: orbit pool threshold (CHAOS_POOL_MIN) x CTA size of the pass kernels x strands
cd "$(dirname "$0")/.."
for pm in ${POOL:-0 12 20 28}; do
for t in ${THREADS:-256 64}; do
for g in ${STRANDS:-2}; do
  for w in ${WORKLOADS:-c2 c2ex2 c4}; do
    CHAOS_POOL_MIN=$pm CHAOS_PASS_THREADS=$t CHAOS_STRANDS=$g timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-full-trips > gpurun_out/pool_${pm}_${t}_${g}_$w.json 2> /dev/null
    python - <<P
import json
d=json.loads(open("gpurun_out/pool_${pm}_${t}_${g}_$w.json").read().strip().splitlines()[-1])
print("pool_min ${pm} threads ${t} strands ${g} $w ms %.3f e2e_ms %.3f frac %.3f"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]))
P
  done
done
done
done
