/*
 * "newton generic": Newton's method on a user-given cubic c0 + c1 x + c2 x^2 + c3 x^3 with its three roots,
 * pixel coloured by the root reached.  Same results as src/main/cuda/fractals/newton_generic.cu:7-117; constants
 * `roots` (3 x (re, im) doubles) and `coefficients` (4 doubles) are written by the host
 * (modules/ModuleNewtonGeneric.java:34-51).
 */
#include "newton_common.cuh"

__constant__ double roots[6];
__constant__ double coefficients[4];

struct NewtonGenericImpl {
    static constexpr bool kTestEveryStep = false;
    template <class Real> static __device__ __forceinline__ thrust::complex<Real> step(thrust::complex<Real> x) { return newton_step_cubic<Real>(coefficients, x); }
    template <class Real> static __device__ __forceinline__ unsigned int root_of(thrust::complex<Real> x)
    {
        typedef thrust::complex<Real> cplx;
        return newton_convergence_root<Real>(x, cplx(roots[0], roots[1]), cplx(roots[2], roots[3]), cplx(roots[4], roots[5]));
    }
};

struct Fractal {
    /* no branch separates c.y's multiply and subtract in the reference build of this module: ptxas contracts them
     * into one FMA (SASS of oracle/_ref/newton_generic.src.cubin), see frame_map::plane_point */
    static constexpr bool kFusedPlaneY = true;
    template <class Real> using Orbit = NewtonOrbit<NewtonGenericImpl, Real>;
    static __device__ __forceinline__ uint32_t colorize(const uint32_t *, uint32_t, float result) { return newton_root_colour(result); }
    static __device__ void debugFractal() {}
};

#include "../render_generic.cuh"
