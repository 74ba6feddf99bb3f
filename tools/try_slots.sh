#!/bin/bash
# experiment: tile slots per warp in the rounds engine (rebuilds the mandelbrot module on the GPU box)
cd "$(dirname "$0")/.."
for K in 4 6 8 12; do
  CHAOS_NVCC_EXTRA="-DCHAOS_REFILL_SLOTS=$K" python -c "
import importlib,sys
sys.path.insert(0,'.')
b=importlib.import_module('chaos-ultra_b200.build')
from pathlib import Path
b.build_module(Path('chaos-ultra_b200/csrc/fractals/mandelbrot.cu'), force=True)
"
  for w in c2 c2ex2 c2f32; do
    python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/slots_${K}_$w.json 2> gpurun_out/slots_${K}_$w.err
  done
done
