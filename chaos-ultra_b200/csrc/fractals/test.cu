/*
 * test module: closed-form pattern (int)((|x|+|y|) * amplifier) -- a known-answer check of the
 * pixel->plane mapping and the constant plumbing.  Same results as
 * src/main/cuda/fractals/test.cu:6-28; host half modules/ModuleTest.java:9-23.
 * Written against the classic one-function contract (ClassicOrbit).
 */
#include "../fractal.cuh"

__constant__ int amplifier;

struct TestImpl {
    template <class Real> static __device__ __forceinline__ float compute(uint32_t, Real px, Real py, uint32_t &trips)
    {
        typedef real_ops<Real> op;
        trips = 0;
        Real m = op::add(op::abs(px), op::abs(py));
        Real v = op::mul(m, (Real)amplifier);
        return (float)(int)v;
    }
};

struct Fractal {
    /* in the reference build of this module ptxas contracts c.y = rt.y - psy*(py+dy) into one FMA
     * (SASS of oracle/_ref/test.src.cubin: DFMA R2, R2, -R16, UR6); mandelbrot/julia keep MUL + SUB */
    static constexpr bool kFusedPlaneY = true;
    template <class Real> using Orbit = ClassicOrbit<TestImpl, Real>;
    static __device__ __forceinline__ uint32_t colorize(const uint32_t *palette, uint32_t len, float result)
    {
        return chaos_default_colorize(palette, len, result, 128u);
    }
    static __device__ void debugFractal() { printf("hello from test\n"); }
};

#include "../render_generic.cuh"
