/*
 * "newton colored by iterations": the same Newton iteration as newton_generic, but the value is the number of
 * steps until any root is reached (tested after every step) and the colour comes from the palette, stretched by
 * `colorMagnifier`.  Same results as src/main/cuda/fractals/newton_iterations.cu:7-84; host half
 * modules/ModuleNewtonIterations.java:9-33.
 */
#include "newton_common.cuh"

__constant__ double roots[6];
__constant__ double coefficients[4];
__constant__ int colorMagnifier;

struct NewtonIterationsImpl {
    static constexpr bool kTestEveryStep = true;
    template <class Real> static __device__ __forceinline__ thrust::complex<Real> step(thrust::complex<Real> x) { return newton_step_cubic<Real>(coefficients, x); }
    template <class Real> static __device__ __forceinline__ unsigned int root_of(thrust::complex<Real> x)
    {
        typedef thrust::complex<Real> cplx;
        return newton_convergence_root<Real>(x, cplx(roots[0], roots[1]), cplx(roots[2], roots[3]), cplx(roots[4], roots[5]));
    }
};

struct Fractal {
    /* no branch separates c.y's multiply and subtract in the reference build of this module: ptxas contracts them
     * into one FMA (SASS of oracle/_ref/newton_iterations.src.cubin), see frame_map::plane_point */
    static constexpr bool kFusedPlaneY = true;
    template <class Real> using Orbit = NewtonOrbit<NewtonIterationsImpl, Real>;
    static __device__ __forceinline__ uint32_t colorize(const uint32_t *palette, uint32_t len, float result)
    {
        return chaos_default_colorize(palette, len, result, (uint32_t)colorMagnifier);
    }
    static __device__ void debugFractal() {}
};

#include "../render_generic.cuh"
