/*
 * fractal.cuh -- the per-fractal module contract (device half).
 *
 * Replaces the reference's src/main/cuda/fractals/fractal.cuh:7-28.  A module is one file
 * fractals/<name>.cu that defines `struct Fractal` and then includes render_generic.cuh; it is
 * compiled to <kernels_dir>/<name>.cubin and loaded by name (FractalRenderingModule.java:63).
 *
 * A module author supplies, as in the reference, three things -- but computeFractal is split so
 * the lane-refill scheduler can suspend and resume an orbit:
 *
 *   struct Fractal {
 *     template <class Real> struct Orbit {
 *       static constexpr bool kResumable = true;             // run() may be called repeatedly with growing limits
 *       __device__ void start(Real px, Real py, const orbit_ctx &ctx);   // point of the plane to evaluate
 *       __device__ bool run(uint32_t &i, uint32_t limit, bool tested);
 *                                     // iterate while i < limit; true = the loop is over: the orbit terminated at
 *                                     // trip i, or it PROVED that the reference's loop runs to ctx.max_iter and set
 *                                     // i to that (only if ctx.shortcuts allows it).  `tested` is uniform over the
 *                                     // warp.  An orbit may use a cheaper instruction stream when tested == false
 *                                     // that cannot tell the exact trip at which it ends; it then steps back,
 *                                     // returns false with i < limit, and reports wants_tested() until it is
 *                                     // called with tested == true (which must always be able to finish the job)
 *       __device__ bool wants_tested() const;
 *       __device__ void force_exact();                       // drop any exactly-equivalent fast form (may be a no-op)
 *       __device__ uint32_t skipped() const;                 // trips such a proof replaced (0 if none)
 *       __device__ void save(Real &x, Real &y) const;        // kResumable: the state after the trips run so far, and
 *       __device__ void resume(Real x, Real y);              // back into an orbit start()ed at the same point (engine 2
 *                                                            // carries an orbit from its long kernel to its finish kernel)
 *       __device__ uint32_t finish(uint32_t i, uint32_t maxIterations) const;
 *                                     // the value `uint escapeTime = computeFractal(..)` would take
 *     };
 *     static __device__ uint32_t colorize(const uint32_t *palette, uint32_t paletteLength, float result);
 *     static __device__ void debugFractal();
 *   };
 *
 * `palette` points at shared memory (staged by the compose kernel), R in the low byte.
 * Modules written in the reference's original style (one opaque computeFractal) can use
 * ClassicOrbit below; they run correctly but an orbit cannot be suspended mid-way.
 *
 * All value-path arithmetic must go through real_ops<Real> (explicit round-to-nearest
 * intrinsics): nvcc must not be free to contract or re-associate anything, because the
 * iteration counts have to match the reference build bit for bit.
 */
#ifndef CHAOS_FRACTAL_CUH
#define CHAOS_FRACTAL_CUH

#include <stdint.h>
#include "chaos_device.h"

/* what an orbit may know about the frame */
struct orbit_ctx {
    uint32_t max_iter;   /* maxIterations of the frame */
    uint32_t shortcuts;  /* CHAOS_SHORTCUT_* the host allows */
};

template <class Real> struct real_ops;

template <> struct real_ops<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float from_u32(uint32_t v) { return __uint2float_rn(v); }
    static __device__ __forceinline__ float from_f32(float v) { return v; }
    static __device__ __forceinline__ float from_f64(double v) { return __double2float_rn(v); }
    static __device__ __forceinline__ float to_f32(float v) { return v; }
    static __device__ __forceinline__ float abs(float v) { return fabsf(v); }
    /* s < 4 for s >= +0 or NaN: plain ordered compare (FSETP issues on the ALU side) */
    static __device__ __forceinline__ bool below4(float s) { return s < 4.0f; }
};

template <> struct real_ops<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double from_u32(uint32_t v) { return __uint2double_rn(v); }
    static __device__ __forceinline__ double from_f32(float v) { return (double)v; }
    static __device__ __forceinline__ double from_f64(double v) { return v; }
    static __device__ __forceinline__ float to_f32(double v) { return __double2float_rn(v); }
    static __device__ __forceinline__ double abs(double v) { return fabs(v); }
    /* s is a sum of two squares: +0 <= s, +inf, or NaN.  For those, "s < 4.0" is exactly
     * "high word < 0x40100000" as unsigned (NaN and inf have a larger high word, a sign bit
     * only appears on NaN and makes it larger still).  This keeps the test off the FP64 pipe. */
    static __device__ __forceinline__ bool below4(double s) { return (uint32_t)__double2hiint(s) < 0x40100000u; }
};

/* Adapter for modules written as one opaque function, the reference's original contract
 * (fractal.cuh:7-8): Impl::template compute<Real>(maxIterations, px, py, trips) -> float */
template <class Impl, class Real> struct ClassicOrbit {
    static constexpr bool kResumable = false; /* the engine calls run() once with limit = maxIterations */
    Real px, py;
    float result;
    __device__ __forceinline__ void start(Real x, Real y, const orbit_ctx &) { px = x; py = y; result = 0.f; }
    __device__ __forceinline__ void force_exact() {}
    __device__ __forceinline__ uint32_t skipped() const { return 0u; }
    __device__ __forceinline__ bool wants_tested() const { return false; }
    __device__ __forceinline__ void save(Real &, Real &) const {}
    __device__ __forceinline__ void resume(Real, Real) {}

    __device__ __forceinline__ bool run(uint32_t &i, uint32_t limit, bool)
    {
        uint32_t trips = 0;
        result = Impl::template compute<Real>(limit, px, py, trips);
        i = trips;
        return true;
    }
    __device__ __forceinline__ uint32_t finish(uint32_t, uint32_t) const { return __float2uint_rz(result); }
};

/* the default palette lookup every shipped module uses (fractal.cuh:20-26) */
static __device__ __forceinline__ uint32_t chaos_default_colorize(const uint32_t *palette, uint32_t len,
                                                                   float result, uint32_t scale = 1u)
{
    uint32_t k = __float2uint_rz(roundf(result)) * scale;
    uint32_t idx = len - (k % len) - 1u;
    return palette[idx];
}

#endif
