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
    template <class Real> static __device__ __forceinline__ thrust::complex<Real> step(thrust::complex<Real> x)
    {
        thrust::complex<Real> x_pow_2 = x * x;
        thrust::complex<Real> x_pow_3 = x_pow_2 * x;
        thrust::complex<Real> f_eval_x = coefficients[0] +
                                         coefficients[1] * x +
                                         coefficients[2] * x_pow_2 +
                                         coefficients[3] * x_pow_3;
        thrust::complex<Real> f_derivative_eval_x = coefficients[1] +
                                                    coefficients[2] * 2 * x +
                                                    coefficients[3] * 3 * x_pow_2;
        return x - (f_eval_x / f_derivative_eval_x);
    }
    template <class Real> static __device__ __forceinline__ unsigned int root_of(thrust::complex<Real> x)
    {
        const thrust::complex<Real> root_a(roots[0], roots[1]);
        const thrust::complex<Real> root_b(roots[2], roots[3]);
        const thrust::complex<Real> root_c(roots[4], roots[5]);
        return newton_convergence_root<Real>(x, root_a, root_b, root_c);
    }
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
     * into one FMA (SASS of oracle/_ref/newton_generic.src.cubin), see frame_map::plane_point */
    static constexpr bool kFusedPlaneY = true;
    template <class Real> using Orbit = ClassicOrbit<NewtonGenericImpl, Real>;
    static __device__ __forceinline__ uint32_t colorize(const uint32_t *, uint32_t, float result) { return newton_root_colour(result); }
    static __device__ void debugFractal() {}
};

#include "../render_generic.cuh"
