/*
 * "newton wired": Newton's method on x^3 - 1, pixel coloured by the root it converges to.
 * Same results as src/main/cuda/fractals/newton_wired.cu:7-85; host half modules/ModuleNewtonWired.java:7-20.
 */
#include "newton_common.cuh"

struct NewtonWiredImpl {
    template <class Real> static __device__ __forceinline__ thrust::complex<Real> step(thrust::complex<Real> x)
    {
        thrust::complex<Real> x_pow_2 = x * x;
        thrust::complex<Real> x_pow_3 = x_pow_2 * x;
        thrust::complex<Real> f_eval_x = x_pow_3 - 1;
        thrust::complex<Real> f_derivative_eval_x = 3 * x_pow_2;
        return x - (f_eval_x / f_derivative_eval_x);
    }
    template <class Real> static __device__ __forceinline__ unsigned int root_of(thrust::complex<Real> x)
    {
        const thrust::complex<Real> root_a(1, 0);
        const thrust::complex<Real> root_b(-0.5, 0.86602540378);
        const thrust::complex<Real> root_c(-0.5, -0.86602540378);
        return newton_convergence_root<Real>(x, root_a, root_b, root_c);
    }
    /* convergence is tested after 10 steps and then every maxIterations/10 further steps (:52-66) */
    template <class Real> static __device__ float compute(uint32_t maxIterations, Real px, Real py, uint32_t &trips)
    {
        thrust::complex<Real> x(px, py);
        unsigned int i = 0;
        unsigned int check_at = 10;
        while (i < maxIterations) {
            x = step<Real>(x);
            ++i;
            if (i == check_at) {
                unsigned int root = root_of<Real>(x);
                if (root != 0) { trips = i; return root; }
                check_at += maxIterations / 10;
            }
        }
        trips = i;
        return root_of<Real>(x);
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
