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

/* one Newton step x - p(x)/p'(x) for p = c[0] + c[1] x + c[2] x^2 + c[3] x^3 (c in double, mixed with complex<Real>
 * exactly as newton_generic.cu:12-26 mixes them: products are promoted to complex<double>, the result is narrowed) */
template <class Real>
static __device__ __forceinline__ thrust::complex<Real> newton_step_cubic(const double *c, thrust::complex<Real> x)
{
    thrust::complex<Real> x_pow_2 = x * x;
    thrust::complex<Real> x_pow_3 = x_pow_2 * x;
    thrust::complex<Real> f_eval_x = c[0] +
                                     c[1] * x +
                                     c[2] * x_pow_2 +
                                     c[3] * x_pow_3;
    thrust::complex<Real> f_derivative_eval_x = c[1] +
                                                c[2] * 2 * x +
                                                c[3] * 3 * x_pow_2;
    return x - (f_eval_x / f_derivative_eval_x);
}

/* the same for the hard-wired polynomial x^3 - 1 (newton_wired.cu:8-16) */
template <class Real>
static __device__ __forceinline__ thrust::complex<Real> newton_step_unity(thrust::complex<Real> x)
{
    thrust::complex<Real> x_pow_2 = x * x;
    thrust::complex<Real> x_pow_3 = x_pow_2 * x;
    thrust::complex<Real> f_eval_x = x_pow_3 - 1;
    thrust::complex<Real> f_derivative_eval_x = 3 * x_pow_2;
    return x - (f_eval_x / f_derivative_eval_x);
}

/* Iterate Step until Root says "converged", testing after 10 steps and then every maxIterations/10 further steps
 * (newton_wired.cu:52-66 = newton_generic.cu:56-70); the value is the root index (0 = none) */
template <class Real, class Step, class Root>
static __device__ __forceinline__ float newton_root_search(uint32_t maxIterations, Real px, Real py, uint32_t &trips, Step step, Root root_of)
{
    thrust::complex<Real> x(px, py);
    unsigned int i = 0;
    unsigned int check_at = 10;
    while (i < maxIterations) {
        x = step(x);
        ++i;
        if (i == check_at) {
            unsigned int root = root_of(x);
            if (root != 0) { trips = i; return root; }
            check_at += maxIterations / 10;
        }
    }
    trips = i;
    return root_of(x);
}

/* The same search as an orbit that can be suspended and resumed (fractal.cuh, `struct Orbit`): its state is the iterate and
 * the trip at which convergence is tested next, so the lane-refill and stream engines can hand a pixel's Newton iteration
 * from block to block and from warp to warp like an escape orbit (kResumable).  Impl supplies
 *     static thrust::complex<Real> step(thrust::complex<Real>)      one Newton step (newton_step_cubic / newton_step_unity)
 *     static unsigned int root_of(thrust::complex<Real>)             which root has been reached, 0 = none
 *     static constexpr bool kTestEveryStep                           newton_iterations.cu:56-60: test after every step, the value
 *                                                                    is the step count; otherwise newton_generic.cu:56-70: test
 *                                                                    after 10 steps and then every maxIterations / 10, the value
 *                                                                    is the root
 * The step and the test are the functions above, called in the reference's order: splitting the loop changes no operation. */
template <class Impl, class Real> struct NewtonOrbit {
    static constexpr bool kResumable = true;
    thrust::complex<Real> x;
    uint32_t check_at, max_iter;
    float result;
    __device__ __forceinline__ void start(Real px, Real py, const orbit_ctx &ctx)
    {
        x = thrust::complex<Real>(px, py);
        check_at = 10u;
        max_iter = ctx.max_iter;
        result = 0.f;
    }
    /* iterate while i < limit; true = over (a root was reached at a test, or i = maxIterations) */
    __device__ __forceinline__ bool run(uint32_t &i, uint32_t limit, bool)
    {
        while (i < limit) {
            x = Impl::template step<Real>(x);
            ++i;
            if (Impl::kTestEveryStep) {
                if (Impl::template root_of<Real>(x) != 0u) { result = (float)i; return true; }
            } else if (i == check_at) {
                const unsigned int root = Impl::template root_of<Real>(x);
                if (root != 0u) { result = (float)root; return true; }
                check_at += max_iter / 10u;
            }
        }
        if (i >= max_iter) {
            result = Impl::kTestEveryStep ? (float)i : (float)Impl::template root_of<Real>(x);
            return true;
        }
        return false;
    }
    __device__ __forceinline__ void force_exact() {}
    __device__ __forceinline__ bool wants_tested() const { return false; }      /* one instruction stream: nothing is deferred */
    __device__ __forceinline__ uint32_t skipped() const { return 0u; }
    __device__ __forceinline__ void save(Real &ox, Real &oy) const { ox = x.real(); oy = x.imag(); }
    __device__ __forceinline__ void resume(Real ox, Real oy) { x = thrust::complex<Real>(ox, oy); }
    __device__ __forceinline__ uint32_t finish(uint32_t, uint32_t) const { return __float2uint_rz(result); }
};

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
