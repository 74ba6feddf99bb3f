/*
 * bench_kernels/peak.cu -- measurement only: peak issue rate of the FP64 and FP32 FMA pipes, the
 * denominator of the iteration kernels' roofline ("P is not in MEASURED_PEAKS.json ->
 * measure it on the box with an unrolled independent-DFMA/FFMA micro-kernel in the same run").
 * Each thread keeps 8 independent accumulators; one loop trip = 8 FMAs per lane.  The result is
 * written so the loop cannot be removed.  Launched by bench.py through cuda.bindings.driver.
 */
#include <stdint.h>

template <class T> __device__ __forceinline__ T fma_t(T a, T b, T c);
template <> __device__ __forceinline__ double fma_t(double a, double b, double c) { return __fma_rn(a, b, c); }
template <> __device__ __forceinline__ float fma_t(float a, float b, float c) { return __fmaf_rn(a, b, c); }

template <class T> __device__ __forceinline__ void peak_body(T *out, uint32_t trips, T seed)
{
    T a0 = seed, a1 = seed + (T)1, a2 = seed + (T)2, a3 = seed + (T)3, a4 = seed + (T)4, a5 = seed + (T)5, a6 = seed + (T)6,
      a7 = seed + (T)7;
    const T m = (T)0.999999, c = (T)1e-3;
#pragma unroll 4
    for (uint32_t i = 0; i < trips; ++i) {
        a0 = fma_t(a0, m, c); a1 = fma_t(a1, m, c); a2 = fma_t(a2, m, c); a3 = fma_t(a3, m, c);
        a4 = fma_t(a4, m, c); a5 = fma_t(a5, m, c); a6 = fma_t(a6, m, c); a7 = fma_t(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

extern "C" __global__ void __launch_bounds__(256) peak_fp64(double *out, uint32_t trips, double seed) { peak_body<double>(out, trips, seed); }
extern "C" __global__ void __launch_bounds__(256) peak_fp32(float *out, uint32_t trips, float seed) { peak_body<float>(out, trips, seed); }

/* The escape loop's own FP64 instruction mix (2 DMUL + 2 DADD + 2 DFMA per trip, scaled form of quadratic.cuh)
 * with 4 independent orbits per thread and no escape test: the issue rate the FP64 pipe reaches on THIS mix,
 * which is what bounds the iteration kernels (dependent chains of 3 inside each orbit, independent across orbits). */
extern "C" __global__ void __launch_bounds__(256) peak_mandel_mix(double *out, uint32_t trips, double seed)
{
    double x[4], y[4], cx[4], cy[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        cx[k] = -0.2 + 1e-3 * k + seed * 1e-9 + 1e-6 * threadIdx.x;   /* inside the main cardioid: never escapes */
        cy[k] = 0.1 + 1e-3 * k;
        x[k] = 0.0; y[k] = 0.0;
    }
#pragma unroll 2
    for (uint32_t i = 0; i < trips; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double xx = __dmul_rn(x[k], x[k]);
            double yy = __dmul_rn(y[k], y[k]);
            double xn = __fma_rn(__dsub_rn(xx, yy), 0.5, cx[k]);
            double s = __dadd_rn(xx, yy);
            y[k] = __fma_rn(x[k], y[k], cy[k]);
            x[k] = xn;
            cy[k] = (__double2hiint(s) < 0x40300000) ? cy[k] : 0.0;   /* keeps the sum live, integer pipe */
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (x[0] + x[1]) + (x[2] + x[3]) + (y[0] + y[1]) + (y[2] + y[3]);
}
