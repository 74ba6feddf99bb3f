/* what a reference-style module includes (src/main/cuda/fractals/fractal.cuh:1): the types and the three
 * declarations of the per-fractal contract; see chaos_compat_pre.cuh */
#include "chaos_compat_pre.cuh"
template <class Real> __device__ float computeFractal(unsigned int maxIterations, Point<Real> z);
__device__ __forceinline__ unsigned int colorize(cudaSurfaceObject_t colorPalette, unsigned int paletteLength, float computationResult);
__device__ void debugFractal();
