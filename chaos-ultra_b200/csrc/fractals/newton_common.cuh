/*
 * newton_common.cuh -- shared pieces of the three Newton-method modules (cubic polynomials, thrust::complex).
 * Reference: src/main/cuda/fractals/newton_wired.cu, newton_generic.cu, newton_iterations.cu.  The arithmetic is
 * thrust::complex's (CCCL, shipped with the toolkit; its scaled division is thrust/detail/complex/arithmetic.h),
 * written as the same expressions so that the same toolkit produces the same operation sequence; the parity
 * tests run the reference's own modules next to these.
 */
#ifndef CHAOS_NEWTON_COMMON_CUH
#define CHAOS_NEWTON_COMMON_CUH

#include <thrust/complex.h>
#include <math.h>
#include "../fractal.cuh"

/* which of three roots x has reached (|re|,|im| of the difference both below 1e-4), 0 = none */
template <class Real>
static __device__ __forceinline__ unsigned int newton_convergence_root(thrust::complex<Real> x, thrust::complex<Real> root_a,
                                                                        thrust::complex<Real> root_b, thrust::complex<Real> root_c)
{
    const Real tolerance = 0.0001;
    thrust::complex<Real> difference;
    difference = x - root_a;
    if (abs(difference.real()) < tolerance && abs(difference.imag()) < tolerance) return 1;
    difference = x - root_b;
    if (abs(difference.real()) < tolerance && abs(difference.imag()) < tolerance) return 2;
    difference = x - root_c;
    if (abs(difference.real()) < tolerance && abs(difference.imag()) < tolerance) return 3;
    return 0;
}

/* fixed colours of the root-coloured modules (helpers.cuh:150-162, R in the low byte) */
static __device__ __forceinline__ uint32_t newton_root_colour(float result)
{
    switch (__float2uint_rz(roundf(result))) {
        case 1: return 0xff0000ffu;   /* RED */
        case 2: return 0xff00ff00u;   /* GREEN */
        case 3: return 0xffff0000u;   /* BLUE */
        default: return 0xff000000u;  /* BLACK */
    }
}

#endif
