#!/bin/bash
# experiment: build the mandelbrot and julia modules with extra nvcc flags into tools/variants/<name>/
# usage: tools/build_variant.sh <name> <nvcc flags...>
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p tools/variants/$name
for f in mandelbrot julia; do
  nvcc -cubin -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 "$@" -I chaos-ultra_b200/csrc chaos-ultra_b200/csrc/fractals/$f.cu -o tools/variants/$name/$f.cubin 2>&1 | grep -iE "error|spill" | grep -v " 0 bytes spill" 
done
