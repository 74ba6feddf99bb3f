/*
 * chaos_compat_post.cuh -- second half of the source-level compatibility layer (see chaos_compat_pre.cuh): wraps the
 * three free functions of a reference-style module (fractal.cuh:7-28) into the `struct Fractal` this backend's generic
 * device code is written against, then includes that code, as compile.sh:12-13 does with fractalRendererGeneric.cu.
 *
 * CHAOS_COMPAT_FUSED_PLANE_Y (default 1): how the pixel -> plane mapping of the reference build rounds c.y.  In the
 * reference's modules ptxas contracts  rt.y - pixelSize.y * (py + dy)  into one FMA unless a branch separates the two
 * operations (it does in mandelbrot.cu and julia.cu, whose loop condition is evaluated first); the shipped CUDA 9.2
 * modules fuse it everywhere.  Pass -DCHAOS_COMPAT_FUSED_PLANE_Y=0 for a module of the first kind.
 */
#ifndef CHAOS_COMPAT_POST_CUH
#define CHAOS_COMPAT_POST_CUH

#ifndef CHAOS_COMPAT_FUSED_PLANE_Y
#define CHAOS_COMPAT_FUSED_PLANE_Y 1
#endif

#include "../fractal.cuh"

struct ChaosCompatModule {
    template <class Real> static __device__ float compute(uint32_t maxIterations, Real px, Real py, uint32_t &trips)
    {
        trips = 0u;       /* an opaque function does not say how long it iterated */
        return computeFractal<Real>(maxIterations, Point<Real>(px, py));
    }
};

static __device__ __forceinline__ uint32_t chaos_compat_colorize(const uint32_t *palette, uint32_t len, float result)
{
    return colorize((cudaSurfaceObject_t)(uintptr_t)palette, len, result);
}
static __device__ __forceinline__ void chaos_compat_debug() { debugFractal(); }

struct Fractal {
    static constexpr bool kFusedPlaneY = CHAOS_COMPAT_FUSED_PLANE_Y != 0;
    template <class Real> using Orbit = ClassicOrbit<ChaosCompatModule, Real>;
    static __device__ __forceinline__ uint32_t colorize(const uint32_t *palette, uint32_t len, float result)
    {
        return chaos_compat_colorize(palette, len, result);
    }
    static __device__ void debugFractal() { chaos_compat_debug(); }
};

#include "../render_generic.cuh"
#endif
