#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/quick.py --settings "X=0 CHAOS_LONG_SMEM=0" --workloads "c2" --steps 10 2>&1 | tee gpurun_out/r04d_quick.txt
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-full-trips 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench c2 ms %.3f e2e %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']))"
CHAOS_LONG_SMEM=0 timeout 300 python bench.py --no-extras --no-cpu-baseline --no-full-trips 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench c2 nocap ms %.3f e2e %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']))"
