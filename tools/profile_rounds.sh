#!/bin/bash
# experiment: scheduler statistics of the rounds engine (rebuilds the mandelbrot module with -DCHAOS_PROFILE on the GPU box)
cd "$(dirname "$0")/.."
CHAOS_NVCC_EXTRA="-DCHAOS_PROFILE" python -c "
import importlib,sys
sys.path.insert(0,'.')
b=importlib.import_module('chaos-ultra_b200.build')
from pathlib import Path
b.build_module(Path('chaos-ultra_b200/csrc/fractals/mandelbrot.cu'), force=True)
"
for w in c2 c2ex2; do
  CHAOS_PROFILE_PRINT=1 python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep "chaos profile" | tail -2 > gpurun_out/prof_$w.txt
done
