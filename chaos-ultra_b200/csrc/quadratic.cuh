/*
 * quadratic.cuh -- the escape loop of z <- z^2 + c, shared by the mandelbrot and julia modules.
 *
 * Reference: src/main/cuda/fractals/mandelbrot.cu:15-20 (= julia.cu:9-14).  As built by nvcc 12.9 for
 * sm_100a (and as in the shipped PTX, mandelbrot.ptx:982-996) one trip is
 *     xx = rn(x*x); yy = rn(y*y); leave unless rn(xx+yy) < 4;
 *     xn = rn(cx + rn(xx-yy)); y = fma(rn(x+x), y, cy); x = xn            -- 2 MUL + 4 ADD + 1 FMA
 * and the trip COUNT must come out bit-identical.  Four things are done differently here, none of which can
 * change a count; (3) and (4) are switched by chaos_render_args::shortcuts and are off in the differential
 * check (force_exact), so every one of them is tested against the plain form on full frames.
 *
 * 1. Trips run in unrolled groups of eight (four near a limit) with the "still below 4" predicates tested once
 *    per group.  After the first failing test the remaining steps of the group compute garbage (inf/NaN, no
 *    traps) that is never looked at; the count is the index of the first failure.
 *
 * 2. FP64 only: the orbit is carried as X = 2x, Y = 2y with CX = 2cx, CY = 2cy:
 *        XX = rn(X*X) = 4xx;  YY = rn(Y*Y) = 4yy;  leave unless rn(XX+YY) < 16;
 *        X' = fma(rn(XX-YY), 0.5, CX) = rn(CX + 2d) = 2xn;   Y' = fma(X, Y, CY) = rn(4xy + 2cy) = 2y'
 *    -- 2 MUL + 2 ADD + 2 FMA, one FP64-pipe instruction fewer per trip.  Scaling by a power of two
 *    commutes with round-to-nearest as long as no value is subnormal or overflows.  Overflow cannot
 *    happen before the escape test fails (|z| < 2).  Subnormals are excluded by construction: the
 *    scaled form is used only when cx and cy are non-zero with 2^-400 <= |c| <= 2^400 (or cy = 0 with
 *    y0 = 0: then y is +0 throughout in either form) and the start point's components are zero or in
 *    the same range.  Then every later x is 0 or >= ulp(cx)/2 >=
 *    2^-453 (a sum of two doubles one of which is cx), every later y is 0 or >= 2^-107 |cy| >= 2^-507
 *    (an exactly formed product-plus-cy), so xx, yy are 0 or >= 2^-1014: normal.  Differences of
 *    nearby normal numbers (xx-yy) are exact in both forms.  Any other orbit takes the 7-operation form.
 *
 * 3. Deferred escape test (kDeferTest).  For |c|^2 < 3.6 the test is monotone along the COMPUTED orbit: if it
 *    fails at trip k it fails at every later trip.  Proof sketch (u = unit round-off): a failing test means
 *    rn(rn(x^2)+rn(y^2)) >= 4 or NaN.  NaN/inf are sticky through the update.  Otherwise r^2 = x^2+y^2 >=
 *    4(1-3u); the computed successor z' differs from z^2+c by at most 3u r^2 + 2u|z'| (three roundings in x',
 *    one in y'), hence |z'| >= r^2(1-3u) - |c| - tiny >= 4(1-3u)^2 - 1.8974 > 2.10, and the next test sees
 *    |z'|^2 (1-u)^2 > 4.4 >= 4 -- by induction for all later trips (overflow gives inf/NaN: fails too; an
 *    underflowing square contributes an absolute error < 2^-148, nothing against a margin of 0.1).
 *    So a group of kGroup = 32 trips runs WITHOUT the sum and the compare (5 FP64 instructions per trip in the
 *    scaled form, 6 in the plain one) and only the state after the group is tested: passing proves all skipped
 *    tests passed; failing restores the state saved before the group, and the group is replayed with a test per
 *    trip, which finds the exact trip.  The squares computed for the test are the first two operations of the
 *    next group.
 *    Two hardware facts shape this (tools/loopbench.cu, measured on B200):
 *    - the end of a group costs about 35 integer-side instructions (test, compare with the kept state,
 *      bookkeeping) and the FP64 pipe takes a warp instruction every other issue slot, so with groups of 8 the
 *      issue port, not the pipe, is the limit: 2.56 T trips/s at 8 warps per scheduler against 3.08 T with groups
 *      of 32 (2.50 T with a test per trip);
 *    - a replay executed by one lane while the other 31 wait costs the whole warp its length.  Replaying inside
 *      run() made c4 (an orbit of some lane ends every ~47 trips) 70 % SLOWER with groups of 32.  Hence run() is
 *      PHASED: the caller says, uniformly for the warp, whether a call is a tested phase or an untested one.  In an
 *      untested phase an orbit whose group fails only steps back and raises wants_tested(); the engines answer
 *      with a short tested phase for the whole warp, in which every lane keeps advancing (with tests), so the
 *      replay is ordinary productive work of one instruction stream.  New orbits start in a tested phase: most
 *      orbits of a frame end within a few trips.
 *
 * 4. Exact recurrence (kDetectCycle).  The update is a pure function of (x, y) for a fixed c.  If the state after
 *    a group equals, bit for bit, a state seen earlier on the same orbit, and every test in between passed, the
 *    computed orbit is periodic and every future test passes: the reference's loop would run to maxIterations.
 *    The orbit reports i = maxIterations at once -- the very value the reference arrives at, not an
 *    approximation; an orbit that does not close exactly (chaotic boundary points, |multiplier| close to 1)
 *    simply runs on.  One earlier state is kept, replaced at trip counts growing by 1.25x (Brent's scheme), and
 *    compared every kCompareEvery = 16 (FP32: 8) trips on the integer pipe when the host sets
 *    CHAOS_SHORTCUT_DENSE_COMPARE, else once per group (the outcome is looked at once per group either way, after the
 *    group's test).  A cycle of period p is seen when the distance to the kept state is a common multiple of p and
 *    the compare distance.  skipped() tells the engine how many trips were proven
 *    instead of executed, so executed work is reported separately from the reference-equivalent count.
 *
 * The "below 4" test itself reads the high word of the sum on the integer pipe (see real_ops).
 */
#ifndef CHAOS_QUADRATIC_CUH
#define CHAOS_QUADRATIC_CUH

#include "fractal.cuh"

#ifndef CHAOS_GROUP_UNROLL
#define CHAOS_GROUP_UNROLL 4   /* how far the eight-trip body of an untested group is unrolled further: 1, 2 or 4 (= the whole
                               * group: 160 FP instructions without a branch; c2 3.72 / 3.68 / 3.61 ms, c4 35.7 / 35.4 / 35.1 ms) */
#endif
#ifndef CHAOS_GROUP
#define CHAOS_GROUP 32u       /* untested trips per group: the escape test, the recurrence compare and their bookkeeping run once per group */
#endif
#ifndef CHAOS_GROUP_LOOP_UNROLL
#define CHAOS_GROUP_LOOP_UNROLL 1   /* groups per iteration of the untested loop (the copies that keep the state before a group
                                     * alive are made once per iteration) */
#endif
/* With CHAOS_SHORTCUT_DENSE_COMPARE the state is compared with the kept state every so many trips of a group (a power of
 * two up to the group).  The kept state is taken at a group boundary, so a cycle of period p shows when the distance to it
 * is a common multiple of p and this number -- and the cycles that computed orbits end in have all sorts of periods
 * (rounding noise around an attracting point), so comparing at group ends only waits for a distance of lcm(p, 32).
 * Executed trips of c2 by this number: 32: 5.35 G, 16: 4.86 G, 8: 4.62 G (host model: 4: -2 %, 1: -2.5 % more); frame 3.03 /
 * 2.92 / 2.91 ms; FP32 2.12 / 1.96 / 1.88 ms.  A compare is 4 LOP3 + 1 min (FP32: 2 + 1) on the integer side of a loop that
 * is bound by the FP64 pipe only when every trip is needed: frames without provable orbits (c4, c2 "M ex 2") lose 2 % / 5 %
 * with 16 / 8, so the host drops the flag when the renderer's previous frame proved next to nothing. */
#ifndef CHAOS_COMPARE_EVERY_F64
#define CHAOS_COMPARE_EVERY_F64 16
#endif
#ifndef CHAOS_COMPARE_EVERY_F32
#define CHAOS_COMPARE_EVERY_F32 8
#endif
#ifndef CHAOS_SAVE_SHIFT
#define CHAOS_SAVE_SHIFT 2   /* the kept state of the recurrence check is replaced at trip counts growing by 1 + 2^-shift.
                              * Executed trips of c2 by shift: 0: 5.98 G, 1: 5.62 G, 2: 5.63 G, 3: 5.97 G, 4: 6.88 G, 5: 8.39 G
                              * (a short distance to the kept state cannot span the longer periods) */
#endif

template <class Real> struct quad_bits;
template <> struct quad_bits<float> {
    static constexpr bool kCanScale = false;
    static __device__ __forceinline__ bool in_safe_range(float) { return false; }
    static __device__ __forceinline__ bool below16(float s) { return s < 16.0f; }
    static __device__ __forceinline__ bool same(float a, float b) { return __float_as_uint(a) == __float_as_uint(b); }
    /* 0 iff (x, y) and (sx, sy) are the same bit patterns (integer pipe) */
    static __device__ __forceinline__ uint32_t differs(float x, float y, float sx, float sy)
    {
        return (__float_as_uint(x) ^ __float_as_uint(sx)) | (__float_as_uint(y) ^ __float_as_uint(sy));
    }
    static __device__ __forceinline__ float never() { return __uint_as_float(0x7fc00000u); }
};
template <> struct quad_bits<double> {
    static constexpr bool kCanScale = true;
    static __device__ __forceinline__ bool in_safe_range(double v)
    {
        uint32_t e = ((uint32_t)__double2hiint(v) >> 20) & 0x7ffu;
        return (e - (1023u - 400u)) <= 800u;
    }
    static __device__ __forceinline__ bool below16(double s) { return (uint32_t)__double2hiint(s) < 0x40300000u; }
    static __device__ __forceinline__ bool same(double a, double b)
    {
        return ((__double2hiint(a) ^ __double2hiint(b)) | (__double2loint(a) ^ __double2loint(b))) == 0;
    }
    static __device__ __forceinline__ uint32_t differs(double x, double y, double sx, double sy)
    {
        return (uint32_t)((__double2hiint(x) ^ __double2hiint(sx)) | (__double2loint(x) ^ __double2loint(sx)) |
                          (__double2hiint(y) ^ __double2hiint(sy)) | (__double2loint(y) ^ __double2loint(sy)));
    }
    static __device__ __forceinline__ double never() { return __hiloint2double(0x7ff80000, 0); }
};

template <class Real> struct quadratic_orbit {
    typedef real_ops<Real> op;
    typedef quad_bits<Real> qb;
    static constexpr bool kResumable = true;
    static constexpr uint32_t kScaled = 1u, kDeferTest = 2u, kDetectCycle = 4u, kPeriodic = 8u, kReplay = 16u, kDense = 32u;
    static constexpr uint32_t kGroup = CHAOS_GROUP;   /* untested trips per group (power of two) */
    static constexpr int kGroupUnroll = CHAOS_GROUP_UNROLL;
    static constexpr int kLoopUnroll = CHAOS_GROUP_LOOP_UNROLL;
    static constexpr int kCompareEvery = sizeof(Real) == 8 ? CHAOS_COMPARE_EVERY_F64 : CHAOS_COMPARE_EVERY_F32;

    Real x, y, cx, cy;      /* kScaled: 2x, 2y, 2cx, 2cy */
    Real sx, sy;            /* the earlier state the orbit is compared with (kDetectCycle) */
    uint32_t next_save;     /* trip count at which (sx, sy) is replaced next; after kPeriodic: the trip of the proof */
    uint32_t mode;
    uint32_t max_iter;

    static __device__ __forceinline__ bool zero_or_safe(Real v) { return v == (Real)0 || qb::in_safe_range(v); }

    __device__ __forceinline__ void init(Real zx, Real zy, Real pcx, Real pcy, const orbit_ctx &ctx)
    {
        mode = 0u;
        max_iter = ctx.max_iter;
        /* (cy == 0 with y0 == 0, the mandelbrot set's real axis: y stays +0 in both forms -- fma(X, +-0, +0) = +0 -- so every
         * y-term is an exact zero and the argument about x is untouched.  It matters for speed, not for results: a warp that
         * holds scaled and unscaled orbits runs both instruction streams one after the other, and the axis row of the full
         * view is 3840 orbits that never escape and never recur -- the launch's tail.) */
        const bool cy_ok = qb::in_safe_range(pcy) || (pcy == (Real)0 && zy == (Real)0);
        if (qb::kCanScale && qb::in_safe_range(pcx) && cy_ok && zero_or_safe(zx) && zero_or_safe(zy)) mode |= kScaled;
        /* |c|^2 < 3.6 (any rounding of this sum is fine, the proof has 5 % to spare) */
        if ((ctx.shortcuts & CHAOS_SHORTCUT_DEFER_TEST) && op::fma(pcx, pcx, op::mul(pcy, pcy)) < (Real)3.6) {
            mode |= kDeferTest;
            if (ctx.shortcuts & CHAOS_SHORTCUT_RECURRENCE) mode |= kDetectCycle;
            if ((mode & kDetectCycle) && (ctx.shortcuts & CHAOS_SHORTCUT_DENSE_COMPARE)) mode |= kDense;
        }
        const Real k = (mode & kScaled) ? (Real)2 : (Real)1;   /* exact */
        x = op::mul(zx, k); y = op::mul(zy, k); cx = op::mul(pcx, k); cy = op::mul(pcy, k);
        sx = sy = qb::never();
        next_save = 0u;
    }
    /* back to the reference's 7-operation trip with a test per trip (halving is exact); the differential check
     * (engine 0 with force_exact) runs this so that it stays an independent implementation */
    __device__ __forceinline__ void force_exact()
    {
        if (mode & kScaled) {
            x = op::mul(x, (Real)0.5); y = op::mul(y, (Real)0.5); cx = op::mul(cx, (Real)0.5); cy = op::mul(cy, (Real)0.5);
        }
        mode = 0u;
    }
    /* the state an orbit is carried through memory with (engine 2: long list -> finish list); start() has set the mode,
     * which depends on c and the start point only */
    __device__ __forceinline__ void save(Real &ox, Real &oy) const { ox = x; oy = y; }
    __device__ __forceinline__ void resume(Real ox, Real oy) { x = ox; y = oy; mode &= ~kReplay; }
    /* trips proven instead of executed (exact recurrence) */
    __device__ __forceinline__ uint32_t skipped() const { return (mode & kPeriodic) ? max_iter - next_save : 0u; }

    template <bool kS> __device__ __forceinline__ bool below(Real s) const { return kS ? qb::below16(s) : op::below4(s); }
    /* one trip with its test */
    template <bool kS> __device__ __forceinline__ bool step()
    {
        Real xx = op::mul(x, x);
        Real yy = op::mul(y, y);
        bool ok = below<kS>(op::add(xx, yy));
        Real xn = kS ? op::fma(op::sub(xx, yy), (Real)0.5, cx) : op::add(cx, op::sub(xx, yy));
        y = kS ? op::fma(x, y, cy) : op::fma(op::add(x, x), y, cy);
        x = xn;
        return ok;
    }
    /* one trip without its test; xx, yy are the squares of the current state on entry and on exit */
    template <bool kS> __device__ __forceinline__ void advance(Real &xx, Real &yy)
    {
        Real xn = kS ? op::fma(op::sub(xx, yy), (Real)0.5, cx) : op::add(cx, op::sub(xx, yy));
        y = kS ? op::fma(x, y, cy) : op::fma(op::add(x, x), y, cy);
        x = xn;
        xx = op::mul(x, x);
        yy = op::mul(y, y);
    }

    /* trips with a test each, while i < stop; true = a test failed at trip i */
    template <bool kS> __device__ __forceinline__ bool run_tested(uint32_t &i, uint32_t stop)
    {
        while (i + 8u <= stop) {      /* one branch per group of eight tested trips (16 measured no better) */
            bool p0 = step<kS>(), p1 = step<kS>(), p2 = step<kS>(), p3 = step<kS>();
            bool p4 = step<kS>(), p5 = step<kS>(), p6 = step<kS>(), p7 = step<kS>();
            if (!(p0 & p1 & p2 & p3 & p4 & p5 & p6 & p7)) {
                i += p0 ? (p1 ? (p2 ? (p3 ? (p4 ? (p5 ? (p6 ? 7u : 6u) : 5u) : 4u) : 3u) : 2u) : 1u) : 0u;
                return true;
            }
            i += 8u;
        }
        while (i + 4u <= stop) {
            bool p0 = step<kS>(), p1 = step<kS>(), p2 = step<kS>(), p3 = step<kS>();
            if (!(p0 & p1 & p2 & p3)) {
                i += p0 ? (p1 ? (p2 ? 3u : 2u) : 1u) : 0u;
                return true;
            }
            i += 4u;
        }
        while (i < stop) {
            if (!step<kS>()) return true;
            ++i;
        }
        return false;
    }
    template <bool kS> __device__ __forceinline__ bool run_as(uint32_t &i, uint32_t limit, bool tested)
    {
        if (tested || !(mode & kDeferTest)) {
            mode &= ~kReplay;
            return run_tested<kS>(i, limit);
        }
        if (mode & kReplay) return false;           /* waits for a tested phase */
        /* (the same for every orbit of a launch: one instruction stream per warp either way) */
        return (mode & kDense) ? run_groups<kS, kCompareEvery>(i, limit) : run_groups<kS, (int)kGroup>(i, limit);
    }
    /* untested groups while a whole one fits below `limit`; the state is compared with the kept one every kEvery trips */
    template <bool kS, int kEvery> __device__ __forceinline__ bool run_groups(uint32_t &i, uint32_t limit)
    {
        static_assert(kEvery > 0 && (kEvery & (kEvery - 1)) == 0 && kEvery <= (int)kGroup, "compare points: a power of two up to the group");
        Real xx = op::mul(x, x), yy = op::mul(y, y);
#pragma unroll kLoopUnroll
        while (i + kGroup <= limit) {
            const Real bx = x, by = y;
            uint32_t apart = 0xffffffffu;           /* 0 once the state was the kept state at one of the group's compare points */
#pragma unroll kGroupUnroll
            for (uint32_t r = 0; r < kGroup / 8u; ++r) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    advance<kS>(xx, yy);
                    if ((r * 8u + (uint32_t)k + 1u) % (uint32_t)kEvery == 0u) apart = min(apart, qb::differs(x, y, sx, sy));
                }
            }
            if (!below<kS>(op::add(xx, yy))) {      /* a test among trips i .. i+kGroup fails: back to the state before the group */
                x = bx; y = by;
                mode |= kReplay;
                return false;
            }
            i += kGroup;
            if (mode & kDetectCycle) {
                if (apart == 0u) {          /* exactly periodic (and the group's test passed): the loop runs to maxIterations */
                    mode |= kPeriodic;
                    next_save = i;
                    i = max_iter;
                    return true;
                }
                if (i >= next_save) {
                    sx = x; sy = y;
                    next_save = i + max(kGroup, (i >> CHAOS_SAVE_SHIFT) & ~(kGroup - 1u));
                }
            }
        }
        if (i < limit) mode |= kReplay;             /* a tail shorter than a group needs its tests */
        return false;
    }
    /* One phase of the loop, while i < limit.  `tested` must be uniform over the warp: all lanes then execute ONE
     * instruction stream.  tested = true: every trip with its test (the first trips of an orbit, the replay of a
     * failed group, tails).  tested = false: untested groups; an orbit whose group fails goes back to the state
     * before that group, asks for a tested phase (wants_tested()) and does nothing until it gets one.
     * true = the loop is over: the test failed at trip i, or i = maxIterations was proven. */
    __device__ __forceinline__ bool run(uint32_t &i, uint32_t limit, bool tested)
    {
        if (qb::kCanScale && (mode & kScaled)) return run_as<true>(i, limit, tested);
        return run_as<false>(i, limit, tested);
    }
    __device__ __forceinline__ bool wants_tested() const { return (mode & kReplay) != 0u; }
};

#endif
