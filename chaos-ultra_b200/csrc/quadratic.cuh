/*
 * quadratic.cuh -- the escape loop of z <- z^2 + c, shared by the mandelbrot and julia modules.
 *
 * Reference: src/main/cuda/fractals/mandelbrot.cu:15-20 (= julia.cu:9-14).  As built by nvcc 12.9 for
 * sm_100a (and as in the shipped PTX, mandelbrot.ptx:982-996) one trip is
 *     xx = rn(x*x); yy = rn(y*y); leave unless rn(xx+yy) < 4;
 *     xn = rn(cx + rn(xx-yy)); y = fma(rn(x+x), y, cy); x = xn            -- 2 MUL + 4 ADD + 1 FMA
 * and the trip COUNT must come out bit-identical.  Two things are done differently here, neither of
 * which can change a count:
 *
 * 1. Trips run in unrolled groups (eight in the FP64 loop, four elsewhere) with the "still below 4" predicates
 *    tested once per group.  After the first failing test the remaining steps of the group compute garbage
 *    (inf/NaN, no traps) that is never looked at; the count is the index of the first failure.
 *
 * 2. FP64 only: the orbit is carried as X = 2x, Y = 2y with CX = 2cx, CY = 2cy:
 *        XX = rn(X*X) = 4xx;  YY = rn(Y*Y) = 4yy;  leave unless rn(XX+YY) < 16;
 *        X' = fma(rn(XX-YY), 0.5, CX) = rn(CX + 2d) = 2xn;   Y' = fma(X, Y, CY) = rn(4xy + 2cy) = 2y'
 *    -- 2 MUL + 2 ADD + 2 FMA, one FP64-pipe instruction fewer per trip.  Scaling by a power of two
 *    commutes with round-to-nearest as long as no value is subnormal or overflows.  Overflow cannot
 *    happen before the escape test fails (|z| < 2).  Subnormals are excluded by construction: the
 *    scaled form is used only when cx and cy are non-zero with 2^-400 <= |c| <= 2^400 and the start
 *    point's components are zero or in the same range.  Then every later x is 0 or >= ulp(cx)/2 >=
 *    2^-453 (a sum of two doubles one of which is cx), every later y is 0 or >= 2^-106 |cy| >= 2^-506
 *    (an exactly formed product-plus-cy), so xx, yy are 0 or >= 2^-1012: normal.  Differences of
 *    nearby normal numbers (xx-yy) are exact in both forms.  Any other orbit takes the 7-operation
 *    form.  The GPU parity tests compare both engines against the oracle and the reference kernels.
 *
 * The "below 4" test itself reads the high word of the sum on the integer pipe (see real_ops).
 */
#ifndef CHAOS_QUADRATIC_CUH
#define CHAOS_QUADRATIC_CUH

#include "fractal.cuh"

template <class Real> struct quadratic_orbit {
    typedef real_ops<Real> op;
    static constexpr bool kResumable = true;
    Real x, y, cx, cy;

    __device__ __forceinline__ void init(Real zx, Real zy, Real pcx, Real pcy)
    {
        x = zx; y = zy; cx = pcx; cy = pcy;
    }
    __device__ __forceinline__ void force_exact() {}
    __device__ __forceinline__ bool step()
    {
        Real xx = op::mul(x, x);
        Real yy = op::mul(y, y);
        bool below = op::below4(op::add(xx, yy));
        Real xn = op::add(cx, op::sub(xx, yy));
        y = op::fma(op::add(x, x), y, cy);
        x = xn;
        return below;
    }
    /* Two orbits stepped together (their dependency chains interleave).  An orbit that ended keeps being stepped
     * -- its state is garbage nobody reads -- so that all lanes of a warp stay in this one loop instead of
     * diverging into a single-orbit loop; the pair is left only when both ended or a live one reaches its limit
     * (the caller finishes leftovers with run()). */
    static __device__ __forceinline__ void run_pair(quadratic_orbit &a, uint32_t &ia, uint32_t la, bool &ea,
                                                    quadratic_orbit &b, uint32_t &ib, uint32_t lb, bool &eb)
    {
        /* groups of four trips both orbits can still take inside their limits; counts are settled once at the end
         * (or at the group in which an orbit ended), the hot loop only counts groups */
        const uint32_t n = min((la - ia) >> 2, (lb - ib) >> 2);
        uint32_t g = 0;
        while (g < n && (!ea | !eb)) {
            bool a0 = a.step(), b0 = b.step(), a1 = a.step(), b1 = b.step(), a2 = a.step(), b2 = b.step(), a3 = a.step(), b3 = b.step();
            const bool fa = !(a0 & a1 & a2 & a3) & !ea, fb = !(b0 & b1 & b2 & b3) & !eb;
            if (fa | fb) {
                if (fa) { ia += 4u * g + (a0 ? (a1 ? (a2 ? 3u : 2u) : 1u) : 0u); ea = true; }
                if (fb) { ib += 4u * g + (b0 ? (b1 ? (b2 ? 3u : 2u) : 1u) : 0u); eb = true; }
            }
            ++g;
        }
        if (!ea) ia += 4u * g;
        if (!eb) ib += 4u * g;
    }
    /* advance while i < limit; true = the escape test failed at trip i (i is the exact count) */
    __device__ __forceinline__ bool run(uint32_t &i, uint32_t limit)
    {
        while (i + 8u <= limit) {
            bool p0 = step(), p1 = step(), p2 = step(), p3 = step(), p4 = step(), p5 = step(), p6 = step(), p7 = step();
            if (!(p0 & p1 & p2 & p3 & p4 & p5 & p6 & p7)) {
                i += p0 ? (p1 ? (p2 ? (p3 ? (p4 ? (p5 ? (p6 ? 7u : 6u) : 5u) : 4u) : 3u) : 2u) : 1u) : 0u;
                return true;
            }
            i += 8u;
        }
        while (i + 4u <= limit) {
            bool p0 = step(), p1 = step(), p2 = step(), p3 = step();
            if (!(p0 & p1 & p2 & p3)) {
                i += p0 ? (p1 ? (p2 ? 3u : 2u) : 1u) : 0u;
                return true;
            }
            i += 4u;
        }
        while (i < limit) {
            if (!step()) return true;
            ++i;
        }
        return false;
    }
};

/* FP64: adds the exactly-scaled 6-operation form */
template <> struct quadratic_orbit<double> {
    static constexpr bool kResumable = true;
    double x, y, cx, cy;   /* scaled: 2x, 2y, 2cx, 2cy */
    bool scaled;

    static __device__ __forceinline__ bool in_safe_range(double v)
    {
        uint32_t e = ((uint32_t)__double2hiint(v) >> 20) & 0x7ffu;
        return (e - (1023u - 400u)) <= 800u;
    }
    static __device__ __forceinline__ bool zero_or_safe(double v) { return v == 0.0 || in_safe_range(v); }

    __device__ __forceinline__ void init(double zx, double zy, double pcx, double pcy)
    {
        scaled = in_safe_range(pcx) && in_safe_range(pcy) && zero_or_safe(zx) && zero_or_safe(zy);
        const double k = scaled ? 2.0 : 1.0;   /* exact */
        x = __dmul_rn(zx, k); y = __dmul_rn(zy, k); cx = __dmul_rn(pcx, k); cy = __dmul_rn(pcy, k);
    }
    /* back to the reference's 7-operation form (halving is exact); the tile-synchronous engine uses this so
     * that it stays an independent implementation to test the scaled form against */
    __device__ __forceinline__ void force_exact()
    {
        if (scaled) {
            x = __dmul_rn(x, 0.5); y = __dmul_rn(y, 0.5); cx = __dmul_rn(cx, 0.5); cy = __dmul_rn(cy, 0.5);
            scaled = false;
        }
    }
    __device__ __forceinline__ bool step_exact()
    {
        double xx = __dmul_rn(x, x);
        double yy = __dmul_rn(y, y);
        bool below = (uint32_t)__double2hiint(__dadd_rn(xx, yy)) < 0x40100000u;   /* < 4.0 */
        double xn = __dadd_rn(cx, __dsub_rn(xx, yy));
        y = __fma_rn(__dadd_rn(x, x), y, cy);
        x = xn;
        return below;
    }
    __device__ __forceinline__ bool step_scaled()
    {
        double xx = __dmul_rn(x, x);
        double yy = __dmul_rn(y, y);
        bool below = (uint32_t)__double2hiint(__dadd_rn(xx, yy)) < 0x40300000u;   /* < 16.0 */
        double xn = __fma_rn(__dsub_rn(xx, yy), 0.5, cx);
        y = __fma_rn(x, y, cy);
        x = xn;
        return below;
    }
    template <bool kScaled> __device__ __forceinline__ bool step() { return kScaled ? step_scaled() : step_exact(); }
    template <bool kScaled> __device__ __forceinline__ bool run_as(uint32_t &i, uint32_t limit)
    {
        while (i + 8u <= limit) {      /* groups of eight trips: one branch per 48 FP64 instructions (16 measured no better) */
            bool p0 = step<kScaled>(), p1 = step<kScaled>(), p2 = step<kScaled>(), p3 = step<kScaled>();
            bool p4 = step<kScaled>(), p5 = step<kScaled>(), p6 = step<kScaled>(), p7 = step<kScaled>();
            if (!(p0 & p1 & p2 & p3 & p4 & p5 & p6 & p7)) {
                i += p0 ? (p1 ? (p2 ? (p3 ? (p4 ? (p5 ? (p6 ? 7u : 6u) : 5u) : 4u) : 3u) : 2u) : 1u) : 0u;
                return true;
            }
            i += 8u;
        }
        while (i + 4u <= limit) {
            bool p0 = step<kScaled>(), p1 = step<kScaled>(), p2 = step<kScaled>(), p3 = step<kScaled>();
            if (!(p0 & p1 & p2 & p3)) {
                i += p0 ? (p1 ? (p2 ? 3u : 2u) : 1u) : 0u;
                return true;
            }
            i += 4u;
        }
        while (i < limit) {
            if (!step<kScaled>()) return true;
            ++i;
        }
        return false;
    }
    __device__ __forceinline__ bool run(uint32_t &i, uint32_t limit)
    {
        return scaled ? run_as<true>(i, limit) : run_as<false>(i, limit);
    }
    template <bool kScaled>
    static __device__ __forceinline__ void run_pair_as(quadratic_orbit &a, uint32_t &ia, uint32_t la, bool &ea,
                                                       quadratic_orbit &b, uint32_t &ib, uint32_t lb, bool &eb)
    {
        /* groups of four trips both orbits can still take inside their limits; counts are settled once at the end
         * (or at the group in which an orbit ended), the hot loop only counts groups */
        const uint32_t n = min((la - ia) >> 2, (lb - ib) >> 2);
        uint32_t g = 0;
        while (g < n && (!ea | !eb)) {
            bool a0 = a.step<kScaled>(), b0 = b.step<kScaled>(), a1 = a.step<kScaled>(), b1 = b.step<kScaled>();
            bool a2 = a.step<kScaled>(), b2 = b.step<kScaled>(), a3 = a.step<kScaled>(), b3 = b.step<kScaled>();
            const bool fa = !(a0 & a1 & a2 & a3) & !ea, fb = !(b0 & b1 & b2 & b3) & !eb;
            if (fa | fb) {
                if (fa) { ia += 4u * g + (a0 ? (a1 ? (a2 ? 3u : 2u) : 1u) : 0u); ea = true; }
                if (fb) { ib += 4u * g + (b0 ? (b1 ? (b2 ? 3u : 2u) : 1u) : 0u); eb = true; }
            }
            ++g;
        }
        if (!ea) ia += 4u * g;
        if (!eb) ib += 4u * g;
    }
    /* two orbits stepped together; a pair in different forms is brought to the 7-operation form first (exact) */
    static __device__ __forceinline__ void run_pair(quadratic_orbit &a, uint32_t &ia, uint32_t la, bool &ea,
                                                    quadratic_orbit &b, uint32_t &ib, uint32_t lb, bool &eb)
    {
        if (a.scaled != b.scaled) { a.force_exact(); b.force_exact(); }
        if (a.scaled) run_pair_as<true>(a, ia, la, ea, b, ib, lb, eb);
        else run_pair_as<false>(a, ia, la, ea, b, ib, lb, eb);
    }
};

#endif
