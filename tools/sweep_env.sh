#!/bin/bash
# experiment: bench a list of workloads under each of the given environment settings
# usage: SETTINGS="A=1+B=2 A=2+B=2,3" WORKLOADS="c2 c4" tools/sweep_env.sh     (variables of one setting are joined by +)
cd "$(dirname "$0")/.."
for set in ${SETTINGS:-X=0}; do
  for w in ${WORKLOADS:-c2 c2ex2 c2f32}; do
    tag=$(echo "$set" | tr '+,=/:' '_____')
    env $(echo "$set" | tr '+' ' ') timeout 300 python bench.py --workload $w --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline --no-full-trips --no-extras > gpurun_out/env_${tag}_$w.json 2> /dev/null
    python - <<P
import json
try:
    d=json.loads(open("gpurun_out/env_${tag}_$w.json").read().strip().splitlines()[-1])
    print("$set $w ms %.3f e2e_ms %.3f frac %.3f"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]))
except Exception as e:
    print("$set $w FAILED", e)
P
  done
done
