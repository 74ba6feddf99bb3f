#!/bin/bash
# experiment: strands of a multi-pass frame (CHAOS_STRANDS) x CTA size of the pass kernels (CHAOS_PASS_THREADS)
cd "$(dirname "$0")/.."
for t in ${THREADS:-256 128 64}; do
for g in ${STRANDS:-1 2 3 4}; do
  for w in ${WORKLOADS:-c2 c2ex2 c2f32}; do
    CHAOS_PASS_THREADS=$t CHAOS_STRANDS=$g timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-full-trips > gpurun_out/st_${t}_${g}_$w.json 2> /dev/null
    python - <<P
import json
d=json.loads(open("gpurun_out/st_${t}_${g}_$w.json").read().strip().splitlines()[-1])
print("threads ${t} strands ${g} $w ms %.3f e2e_ms %.3f frac %.3f"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]))
P
  done
done
done
