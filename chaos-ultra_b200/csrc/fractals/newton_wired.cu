/*
 * "newton wired": Newton's method on x^3 - 1, pixel coloured by the root it converges to.
 * Same results as src/main/cuda/fractals/newton_wired.cu:7-85; host half modules/ModuleNewtonWired.java:7-20.
 */
#include "newton_common.cuh"

struct NewtonWiredImpl {
    template <class Real> static __device__ float compute(uint32_t maxIterations, Real px, Real py, uint32_t &trips)
    {
        typedef thrust::complex<Real> cplx;
        auto root_of = [](cplx x) {
            return newton_convergence_root<Real>(x, cplx(1, 0), cplx(-0.5, 0.86602540378), cplx(-0.5, -0.86602540378));
        };
        return newton_root_search<Real>(maxIterations, px, py, trips, [](cplx x) { return newton_step_unity<Real>(x); }, root_of);
    }
};

struct Fractal {
    /* no branch separates c.y's multiply and subtract in the reference build of this module: ptxas contracts them
     * into one FMA (SASS of oracle/_ref/newton_wired.src.cubin), see frame_map::plane_point */
    static constexpr bool kFusedPlaneY = true;
    template <class Real> using Orbit = ClassicOrbit<NewtonWiredImpl, Real>;
    static __device__ __forceinline__ uint32_t colorize(const uint32_t *, uint32_t, float result) { return newton_root_colour(result); }
    static __device__ void debugFractal() {}
};

#include "../render_generic.cuh"
