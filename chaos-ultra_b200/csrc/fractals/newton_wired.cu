/*
 * "newton wired": Newton's method on x^3 - 1, pixel coloured by the root it converges to.
 * Same results as src/main/cuda/fractals/newton_wired.cu:7-85; host half modules/ModuleNewtonWired.java:7-20.
 */
#include "newton_common.cuh"

struct NewtonWiredImpl {
    static constexpr bool kTestEveryStep = false;
    template <class Real> static __device__ __forceinline__ thrust::complex<Real> step(thrust::complex<Real> x) { return newton_step_unity<Real>(x); }
    template <class Real> static __device__ __forceinline__ unsigned int root_of(thrust::complex<Real> x)
    {
        typedef thrust::complex<Real> cplx;
        return newton_convergence_root<Real>(x, cplx(1, 0), cplx(-0.5, 0.86602540378), cplx(-0.5, -0.86602540378));
    }
};

struct Fractal {
    /* no branch separates c.y's multiply and subtract in the reference build of this module: ptxas contracts them
     * into one FMA (SASS of oracle/_ref/newton_wired.src.cubin), see frame_map::plane_point */
    static constexpr bool kFusedPlaneY = true;
    template <class Real> using Orbit = NewtonOrbit<NewtonWiredImpl, Real>;
    static __device__ __forceinline__ uint32_t colorize(const uint32_t *, uint32_t, float result) { return newton_root_colour(result); }
    static __device__ void debugFractal() {}
};

#include "../render_generic.cuh"
