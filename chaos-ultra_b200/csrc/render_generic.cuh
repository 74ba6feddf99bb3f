/*
 * render_generic.cuh -- generic device code of a fractal module, B200 (sm_100a).
 *
 * Included at the end of every fractals/<name>.cu after `struct Fractal` is defined; the
 * resulting cubin exports the entry names the reference host looks up
 * (FractalRenderingModule.java:91-97): fractalRenderMain{Float,Double},
 * fractalRenderAdvanced{Float,Double}, compose, fractalRenderUnderSampled, debug, init, and
 * the constant VISUALIZE_SAMPLE_COUNT.  Each takes one by-value struct (chaos_device.h).
 *
 * Behavioural contract (what must equal the reference, not how it is computed):
 *   sampleTheFractal      src/main/cuda/fractalRendererGeneric.cu:85-155
 *   computeDispersion     :24-34         fractalRenderMain      :168-181
 *   foveation             :209-257       reuse / reprojection   :194-203, :259-302
 *   fractalRenderAdvanced :307-368       compose                :455-475, :67-77
 * The reference decides "stop supersampling" with warp votes over the 8x4 pixel rectangle
 * one warp covers (:36-53, :140-148).  Here a *vote tile* is that same aligned 8x4 rectangle,
 * but tiles are handed out by a persistent work-stealing scheduler instead of being tied to
 * blockIdx, and (engine 1) the orbits of a tile are not tied to fixed lanes.
 */
#ifndef CHAOS_RENDER_GENERIC_CUH
#define CHAOS_RENDER_GENERIC_CUH

#include <stdio.h>
#include <float.h>
#include <math.h>
#include "chaos_device.h"
#include "fractal.cuh"

__constant__ bool VISUALIZE_SAMPLE_COUNT = false;          /* fractalRendererGeneric.cu:14 */
__constant__ uint32_t CHAOS_MODULE_ABI_VERSION = CHAOS_MODULE_ABI;

#define CHAOS_FULL_MASK 0xffffffffu
#define CHAOS_RENDER_THREADS 256
#define CHAOS_ADAPTIVE_THRESHOLD 10u                       /* :96 adaptiveTreshold */
#define CHAOS_MAX_SAMPLES 64u                              /* :12 MAX_SUPER_SAMPLING */

/* ------------------------------------------------------------------------------------------
 * frame constants every orbit needs, computed once per thread exactly like the reference does
 * per thread (:97): pixelSize = (rt - lb) / (Real) gridSize with IEEE division.
 * ---------------------------------------------------------------------------------------- */
template <class Real> struct frame_map {
    typedef real_ops<Real> op;
    Real lbx, rty, psx, psy;
    __device__ __forceinline__ void init(const chaos_render_args &a);
    /* c = image_left_top + (1,-1) * (pixel + delta) * pixelSize  (:119-124) as nvcc 12.9 compiles it:
     * c.x = fma(psx, dx + px, lb.x);  c.y = rt.y - rn(psy * (dy + py)).  kFusedY: modules in whose reference build
     * ptxas contracted that mul+sub into one FMA (the `test` module; see its SASS) say so with
     * `static constexpr bool kFusedPlaneY = true` */
    template <bool kFusedY>
    __device__ __forceinline__ void plane_point(uint32_t px, uint32_t py, Real dx, Real dy, Real &cx, Real &cy) const
    {
        Real ax = op::add(dx, op::from_u32(px));
        Real ay = op::add(dy, op::from_u32(py));
        cx = op::fma(psx, ax, lbx);
        cy = kFusedY ? op::fma(psy, -ay, rty) : op::sub(rty, op::mul(psy, ay));
    }
};
template <> __device__ __forceinline__ void frame_map<double>::init(const chaos_render_args &a)
{
    lbx = a.image[0]; rty = a.image[3];
    psx = __ddiv_rn(__dsub_rn(a.image[2], a.image[0]), __uint2double_rn(a.width));
    psy = __ddiv_rn(__dsub_rn(a.image[3], a.image[1]), __uint2double_rn(a.height));
}
template <> __device__ __forceinline__ void frame_map<float>::init(const chaos_render_args &a)
{
    lbx = a.imagef[0]; rty = a.imagef[3];
    psx = __fdiv_rn(__fsub_rn(a.imagef[2], a.imagef[0]), __uint2float_rn(a.width));
    psy = __fdiv_rn(__fsub_rn(a.imagef[3], a.imagef[1]), __uint2float_rn(a.height));
}

/* does the module declare kFusedPlaneY? (default: no) */
template <class F, class = void> struct fused_plane_y { static constexpr bool value = false; };
template <class F> struct fused_plane_y<F, decltype((void)F::kFusedPlaneY)> { static constexpr bool value = F::kFusedPlaneY; };

/* sample offset inside the pixel for sample index i (:103-117).  scf is the float sample
 * budget the call started with (it is NOT the clamped integer count). */
template <class Real>
static __device__ __forceinline__ void sample_delta(uint32_t i, float spr /* sqrtf(scf - 2) */, Real &dx, Real &dy)
{
    typedef real_ops<Real> op;
    if (i <= 2u) {
        dx = op::div(op::from_u32(i), (Real)3);
        dy = dx;
    } else {
        uint32_t nrow = __float2uint_rz(roundf(spr));
        uint32_t a = i - 2u;
        uint32_t q = a / nrow;
        uint32_t r = a - q * nrow;
        dx = op::from_f32(__fdiv_rn(__uint2float_rn(r), spr));
        dy = op::from_f32(__fdiv_rn(__uint2float_rn(q), spr));
    }
}

/* index of dispersion over the first n samples (:24-34), n < 10 here */
static __device__ __forceinline__ float dispersion(const float *s, uint32_t n, float mean)
{
    float var = 0.f;
    for (uint32_t k = 0; k < n; ++k) {
        float d = __fsub_rn(s[k], mean);
        var = __fmaf_rn(d, d, var);
    }
    var = __fdiv_rn(var, __uint2float_rn(n - 1u));
    return __fdiv_rn(var, mean);
}

/* per-pixel predicates of the decision block (:128-150) after sample i, for a pixel with the
 * running sum `sum` and stored samples s[0..min(i,9)] */
struct vote_preds { bool eq, lt, le, zero_mean; };
static __device__ __forceinline__ vote_preds decision_preds(const float *s, uint32_t i, uint32_t sum)
{
    vote_preds p;
    float mean = __uint2float_rn(sum / (i + 1u));
    /* the reference would read past samples[10] when i > 10 (only at i == S/2 >= 11); those
     * reads are undefined there and are taken as 0 here ("reference UB") */
    float var = 0.f;
    uint32_t n = i < CHAOS_ADAPTIVE_THRESHOLD ? i : CHAOS_ADAPTIVE_THRESHOLD;
    for (uint32_t k = 0; k < n; ++k) {
        float d = __fsub_rn(s[k], mean);
        var = __fmaf_rn(d, d, var);
    }
    for (uint32_t k = n; k < i; ++k) {
        float d = __fsub_rn(0.f, mean);
        var = __fmaf_rn(d, d, var);
    }
    float disp = __fdiv_rn(__fdiv_rn(var, __uint2float_rn(i - 1u)), mean);
    p.zero_mean = mean == 0.f;      /* every dispersion of this pixel is NaN or inf as long as its mean stays 0: it vetoes both stop rules */
    p.eq = (i == 1u) && (fabsf(__fsub_rn(s[0], s[1])) < FLT_EPSILON);
    p.lt = disp < 0.01f;
    p.le = disp <= 1.0f;
    return p;
}
/* the tile-uniform update of the sample bound S after sample i, given the three ALL votes */
static __device__ __forceinline__ uint32_t decision_update(uint32_t i, uint32_t S, bool all_eq, bool all_lt, bool all_le)
{
    if (i == 1u && all_eq) return 2u;
    if (all_lt) return i + 1u;
    if (i >= (S >> 1) && all_le) return i + 1u;
    return S;
}
static __device__ __forceinline__ bool decision_entered(bool adaptive, uint32_t i, uint32_t S)
{
    return adaptive ? (((i - 1u) < 9u) || (i == (S >> 1))) : (i == (S >> 1));
}

/* vote tile t of this launch -> pixel origin, honouring the row-band partition */
static __device__ __forceinline__ void tile_origin_of(uint32_t tiles_x, uint32_t part_index, uint32_t part_count, uint32_t band_tile_rows,
                                                      uint32_t t, uint32_t &x0, uint32_t &y0)
{
    uint32_t row = t / tiles_x;
    uint32_t col = t - row * tiles_x;
    if (part_count > 1u) {
        /* local tile row -> global: bands of band_tile_rows rows are dealt round-robin */
        uint32_t band_local = row / band_tile_rows;
        uint32_t in_band = row - band_local * band_tile_rows;
        row = (band_local * part_count + part_index) * band_tile_rows + in_band;
    }
    x0 = col * 8u;
    y0 = row * 4u;
}
static __device__ __forceinline__ void tile_origin(const chaos_render_args &a, uint32_t t, uint32_t &x0, uint32_t &y0)
{
    tile_origin_of(a.tiles_x, a.part_index, a.part_count, a.band_tile_rows, t, x0, y0);
}

static __device__ __forceinline__ chaos_pixel_info *record_at(chaos_pixel_info *base, uint64_t pitch, uint32_t x, uint32_t y)
{
    return (chaos_pixel_info *)((char *)base + (size_t)y * pitch) + x;
}
static __device__ __forceinline__ const chaos_pixel_info *record_at(const chaos_pixel_info *base, uint64_t pitch, uint32_t x, uint32_t y)
{
    return (const chaos_pixel_info *)((const char *)base + (size_t)y * pitch) + x;
}
static __device__ __forceinline__ void store_record(chaos_pixel_info *p, float value, float weight, uint32_t reused, float wnew)
{
    /* one 128-bit store per pixel */
    float4 v = make_float4(value, weight, __uint_as_float(reused), wnew);
    *reinterpret_cast<float4 *>(p) = v;
}

/* exact work counters: warp-reduce, then one atomic per warp and counter */
static __device__ __forceinline__ void flush_counters(const chaos_render_args &a, unsigned long long iters, unsigned long long nsamples,
                                                      unsigned long long skipped)
{
    for (int o = 16; o; o >>= 1) {
        iters += __shfl_xor_sync(CHAOS_FULL_MASK, iters, o);
        nsamples += __shfl_xor_sync(CHAOS_FULL_MASK, nsamples, o);
        skipped += __shfl_xor_sync(CHAOS_FULL_MASK, skipped, o);
    }
    if ((threadIdx.x & 31u) == 0 && nsamples) {
        atomicAdd(&a.counters->pixel_iterations, iters);
        atomicAdd(&a.counters->samples, nsamples);
        if (skipped) atomicAdd(&a.counters->skipped_iterations, skipped);
    }
}

/* A whole orbit in warp-uniform phases (see Orbit::run in fractal.cuh): the first trips tested (most orbits end
 * there), then untested until every lane is through or stuck on a failed group, then one tested phase in which
 * all stuck lanes replay their group together.  Lanes that are done sit the phases out. */
#define CHAOS_SYNC_TESTED_HEAD 64u
template <class Orbit>
static __device__ __forceinline__ void run_whole(Orbit &o, uint32_t &it, uint32_t max_iter)
{
    bool ended = o.run(it, Orbit::kResumable ? min(max_iter, CHAOS_SYNC_TESTED_HEAD) : max_iter, true);
    if (!ended && it < max_iter) ended = o.run(it, max_iter, false);
    if (!ended && it < max_iter) o.run(it, max_iter, true);
}

/* ==========================================================================================
 * Engine 0: tile-synchronous.  One warp = one vote tile, lane = pixel (lane = 8*row + col),
 * all lanes step through the sample rounds together.  Simple; kept as the differential
 * reference for engine 1 and used by the advanced kernel for its (rare) sampled tiles.
 * ======================================================================================== */
/* the sample offsets of a call with budget scf, for all 64 possible sample indices: a division or three each (:103-117), so
 * a kernel whose tiles all sample with the same budget computes them once per CTA (shared memory) instead of per sample */
template <class Real> struct sample_offsets {
    Real dx[CHAOS_MAX_SAMPLES], dy[CHAOS_MAX_SAMPLES];
    float scf;      /* the budget the table was made for */
    __device__ __forceinline__ void fill(float budget)     /* all threads of the CTA; ends with a barrier */
    {
        if (threadIdx.x < CHAOS_MAX_SAMPLES) sample_delta<Real>(threadIdx.x, sqrtf(__fadd_rn(budget, -2.0f)), dx[threadIdx.x], dy[threadIdx.x]);
        if (threadIdx.x == 0) scf = budget;
        __syncthreads();
    }
};

template <class Real, class FractalT>
static __device__ __forceinline__ uint32_t sample_tile_sync(const chaos_render_args &a, const frame_map<Real> &fm,
                                                            bool participate, uint32_t px, uint32_t py, float &scf,
                                                            unsigned long long &iters, unsigned long long &nsamples,
                                                            unsigned long long &skipped, const sample_offsets<Real> *table = nullptr)
{
    typedef typename FractalT::template Orbit<Real> Orbit;
    const orbit_ctx ctx = {a.max_iter, a.force_exact ? 0u : a.shortcuts};
    if (scf < 1.f) { scf = 0.f; return 0u; }                           /* :87-90 (tile-uniform) */
    uint32_t S = min(64u, __float2uint_rz(roundf(scf)));
    const float spr = sqrtf(__fadd_rn(scf, -2.0f));
    const bool adaptive = (a.flags & CHAOS_FLAG_ADAPTIVE_SS) != 0u;
    const bool tabled = table != nullptr && table->scf == scf;      /* (bit-identical values: the table holds what sample_delta returns) */
    float samples[CHAOS_ADAPTIVE_THRESHOLD];
    uint32_t sum = 0;
    uint32_t i = 0;
    do {
        if (participate) {
            Real dx, dy, cx, cy;
            if (tabled) { dx = table->dx[i]; dy = table->dy[i]; }
            else sample_delta<Real>(i, spr, dx, dy);
            fm.template plane_point<fused_plane_y<FractalT>::value>(px, py, dx, dy, cx, cy);
            Orbit o;
            o.start(cx, cy, ctx);
            if (a.force_exact) o.force_exact();             /* engine 0 as differential check: the reference's own operation sequence */
            uint32_t it = 0;
            run_whole(o, it, a.max_iter);
            uint32_t et = o.finish(it, a.max_iter);
            iters += it;
            skipped += o.skipped();
            nsamples += 1;
            sum += et;
            if (i < CHAOS_ADAPTIVE_THRESHOLD) {
                /* static indexing keeps samples[] in registers */
#pragma unroll
                for (uint32_t k = 0; k < CHAOS_ADAPTIVE_THRESHOLD; ++k)
                    if (k == i) samples[k] = __uint2float_rn(et);
            }
        }
        if (decision_entered(adaptive, i, S)) {
            if (i == 1u) {
                /* after sample 1 the dispersion is a division by (uint) 1 - 1 = 0: +inf or NaN for every pixel, so neither
                 * `disp < 0.01` nor `disp <= 1` can hold and only the first rule (:140-142, both samples equal) can end
                 * the tile -- decision_preds/decision_update at i == 1, written out (most tiles of a frame end here) */
                const bool eq = !participate || fabsf(__fsub_rn(samples[0], samples[1])) < FLT_EPSILON;
                if (__all_sync(CHAOS_FULL_MASK, eq)) S = 2u;
            } else {
                vote_preds p = {true, true, true, false};
                if (participate) p = decision_preds(samples, i, sum);
                bool all_eq = __all_sync(CHAOS_FULL_MASK, p.eq);
                bool all_lt = __all_sync(CHAOS_FULL_MASK, p.lt);
                bool all_le = __all_sync(CHAOS_FULL_MASK, p.le);
                S = decision_update(i, S, all_eq, all_lt, all_le);
            }
        }
        ++i;
    } while (i < S);
    scf = __uint2float_rn(S);
    return sum / S;
}

template <class Real, class FractalT>
static __device__ void render_main_sync(const chaos_render_args &a)
{
    __shared__ sample_offsets<Real> offsets;
    offsets.fill(a.max_ss);
    frame_map<Real> fm;
    fm.init(a);
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long iters = 0, nsamples = 0, skipped = 0;
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(&a.counters->next_tile, 1u);
        t = __shfl_sync(CHAOS_FULL_MASK, t, 0);
        if (t >= a.n_tiles) break;
        uint32_t x0, y0;
        tile_origin(a, t, x0, y0);
        uint32_t px = x0 + (lane & 7u), py = y0 + (lane >> 3);
        bool inb = px < a.width && py < a.height;
        float scf = a.max_ss;
        uint32_t v = sample_tile_sync<Real, FractalT>(a, fm, inb, px, py, scf, iters, nsamples, skipped, &offsets);
        if (inb) store_record(record_at(a.out, a.out_pitch, px, py), __uint2float_rn(v), scf, 0u, 0.f);
    }
    flush_counters(a, iters, nsamples, skipped);     /* exact work counters: warp-reduce then one atomic per warp */
}

/* ------------------------------------------------------------------------------------------
 * foveation (:209-257).  Depends only on the tile origin, the focus and maxSuperSampling, so it
 * is tile-uniform.  atanf/sqrtf/roundf are the same libdevice routines the reference build
 * inlines (same toolkit), divisions are IEEE.
 * ---------------------------------------------------------------------------------------- */
static __device__ __forceinline__ void foveation(uint32_t x0, uint32_t y0, uint32_t fx, uint32_t fy, float max_ss,
                                                 float &advised, bool &inside)
{
    const float pixelRealWidthInCm = 0.02652f, screenDistance = 60.f;
    float dx = __fsub_rn(__uint2float_rn(fx), __uint2float_rn(x0));
    float dy = __fsub_rn(__uint2float_rn(fy), __uint2float_rn(y0));
    float dist = __fmul_rn(sqrtf(__fmaf_rn(dx, dx, __fmul_rn(dy, dy))), pixelRealWidthInCm);
    float angle = __fdiv_rn(__fmul_rn(atanf(__fdiv_rn(dist, screenDistance)), 180.f), 3.14159265358979f);
    float thr = max_ss >= 1.f ? 5.5f : __fmul_rn(max_ss, 5.5f);
    /* linearMapping<float>(angle, thr, 60, 1, 0): k = (0-1)/(60-thr), q = (60*1 - 0*thr)/(60-thr) */
    float den = __fsub_rn(60.f, thr);
    float k = __frcp_rn(den);
    float q = __fdiv_rn(__fmaf_rn(thr, -0.f, 60.f), den);
    float lin = __fmaf_rn(-k, angle, q);   /* one FFMA in the reference build (ptxas contracts k*x+q) */
    float rq = __double2float_rn(fmin((double)lin, 1.0));
    advised = __fmul_rn(max_ss, rq);
    inside = false;
    if (angle <= thr) {
        advised = __double2float_rn(fmax((double)advised, 1.0));
        inside = true;
    }
}

/* previous-frame pixel coordinate of output pixel (px,py) (:194-203) */
template <class Real>
static __device__ __forceinline__ void warp_origin(const chaos_render_args &a, uint32_t px, uint32_t py, float &ox, float &oy);
template <>
__device__ __forceinline__ void warp_origin<double>(const chaos_render_args &a, uint32_t px, uint32_t py, float &ox, float &oy)
{
    double rx = __ddiv_rn(__uint2double_rn(px), __uint2double_rn(a.width));
    double ry = __ddiv_rn(__uint2double_rn(a.height - py), __uint2double_rn(a.height));
    double planex = __fma_rn(__dsub_rn(a.image[2], a.image[0]), rx, a.image[0]);
    double planey = __fma_rn(__dsub_rn(a.image[3], a.image[1]), ry, a.image[1]);
    double relx = __ddiv_rn(__dsub_rn(planex, a.image_reused[0]), __dsub_rn(a.image_reused[2], a.image_reused[0]));
    double rely = __ddiv_rn(__dsub_rn(planey, a.image_reused[1]), __dsub_rn(a.image_reused[3], a.image_reused[1]));
    float fw = __uint2float_rn(a.width), fh = __uint2float_rn(a.height);
    ox = __fmul_rn(fw, __double2float_rn(relx));
    oy = __fmaf_rn(fh, -__double2float_rn(rely), fh);   /* fh - fh*y is one FFMA in the reference build */
}
template <>
__device__ __forceinline__ void warp_origin<float>(const chaos_render_args &a, uint32_t px, uint32_t py, float &ox, float &oy)
{
    float fw = __uint2float_rn(a.width), fh = __uint2float_rn(a.height);
    float rx = __fdiv_rn(__uint2float_rn(px), fw);
    float ry = __fdiv_rn(__uint2float_rn(a.height - py), fh);
    float planex = __fmaf_rn(__fsub_rn(a.imagef[2], a.imagef[0]), rx, a.imagef[0]);
    float planey = __fmaf_rn(__fsub_rn(a.imagef[3], a.imagef[1]), ry, a.imagef[1]);
    float relx = __fdiv_rn(__fsub_rn(planex, a.image_reusedf[0]), __fsub_rn(a.image_reusedf[2], a.image_reusedf[0]));
    float rely = __fdiv_rn(__fsub_rn(planey, a.image_reusedf[1]), __fsub_rn(a.image_reusedf[3], a.image_reusedf[1]));
    ox = __fmul_rn(fw, relx);
    oy = __fmaf_rn(fh, -rely, fh);
}

/* row j of the previous frame: one buffer, or (multi-GPU fast frames) the primary buffer of the rank whose slab holds it */
static __device__ __forceinline__ const char *previous_row(const chaos_render_args &a, uint32_t j)
{
    const chaos_pixel_info *base = a.in;
    if (a.slab_rows) base = a.in_peer[min(j / a.slab_rows, a.part_count - 1u)];
    return (const char *)base + (size_t)j * a.in_pitch;
}

/* 4-tap filter of value and weight at (ox,oy) (:259-302); only value/weight are read, as one
 * 64-bit load per tap */
static __device__ __forceinline__ void gather_bilinear(const chaos_render_args &a, float ox, float oy, float &value, float &weight)
{
    uint32_t i = __float2uint_rz(floorf(ox)), j = __float2uint_rz(floorf(oy));
    float al = __fsub_rn(ox, __uint2float_rn(i)), be = __fsub_rn(oy, __uint2float_rn(j));
    const char *row0 = previous_row(a, j);
    const char *row1 = previous_row(a, j + 1u);
    float2 t00 = __ldg(reinterpret_cast<const float2 *>(row0 + (size_t)i * 16u));
    float2 t10 = __ldg(reinterpret_cast<const float2 *>(row0 + (size_t)(i + 1u) * 16u));
    float2 t01 = __ldg(reinterpret_cast<const float2 *>(row1 + (size_t)i * 16u));
    float2 t11 = __ldg(reinterpret_cast<const float2 *>(row1 + (size_t)(i + 1u) * 16u));
    float na = __fsub_rn(1.f, al), nb = __fsub_rn(1.f, be);
    float w00 = __fmul_rn(na, nb), w10 = __fmul_rn(al, nb), w01 = __fmul_rn(na, be), w11 = __fmul_rn(al, be);
    value = __fmaf_rn(w11, t11.x, __fmaf_rn(w01, t01.x, __fmaf_rn(w00, t00.x, __fmul_rn(w10, t10.x))));
    weight = __fmaf_rn(w11, t11.y, __fmaf_rn(w01, t01.y, __fmaf_rn(w00, t00.y, __fmul_rn(w10, t10.y))));
}

/*
 * fractalRenderAdvanced (:307-368).  A fast frame is two very different kinds of work:
 *   - almost every pixel only reprojects: 4-tap gather from the previous frame + one 16-byte store (HBM-bound);
 *   - the foveal disc (resampled while zooming in) and the pixels without history (border ring) run escape loops.
 * They are split into two launches so that neither waits for the other:
 *   pass R (advanced_reuse_pass)  static grid-stride over all vote tiles, one warp per tile.  Tiles in which no
 *          pixel needs a sample are finished here; the others are appended to a worklist (tile_order[]) untouched.
 *   pass S (advanced_sample_pass) dynamic (one atomic per listed tile) over the worklist; does the whole
 *          per-tile procedure of the reference, both call sites of sampleTheFractal with their own vote groups.
 * kMode: 1 = pass R, 2 = pass S.
 */
template <class Real, class FractalT, int kMode>
static __device__ __forceinline__ void advanced_tile(const chaos_render_args &a, const frame_map<Real> &fm, uint32_t t, uint32_t lane,
                                                     bool use_fov, unsigned long long &iters, unsigned long long &nsamples,
                                                     unsigned long long &skipped)
{
    const uint32_t fl = a.flags;
    uint32_t x0, y0;
    tile_origin(a, t, x0, y0);
    const uint32_t px = x0 + (lane & 7u), py = y0 + (lane >> 3);
    const bool inb = px < a.width && py < a.height;

    float advised = a.max_ss;
    bool inside = false;
    if (use_fov) foveation(x0, y0, a.focus_x, a.focus_y, a.max_ss, advised, inside);

    bool reusing = false;
    float rv = 0.f, rw = 0.f;
    if (inb && (fl & CHAOS_FLAG_SAMPLE_REUSE)) {
        float ox, oy;
        warp_origin<Real>(a, px, py, ox, oy);
        int oix = __float2int_rz(roundf(ox)), oiy = __float2int_rz(roundf(oy));
        if (!(oix < 2 || (uint32_t)oix >= a.width - 2u || oiy < 2 || (uint32_t)oiy >= a.height - 2u)) {
            gather_bilinear(a, ox, oy, rv, rw);
            reusing = !((double)rw < 0.1);
        }
    }
    const bool resample = reusing && (fl & CHAOS_FLAG_ZOOMING_IN) && inside;    /* call site :351 */
    const bool fresh = inb && !reusing;                                          /* call site :361 */
    if (kMode == 1) {
        if (__any_sync(CHAOS_FULL_MASK, resample || fresh)) {                    /* leave the whole tile to pass S */
            if (lane == 0) a.tile_order[atomicAdd(&a.counters->bucket_count[0], 1u)] = t;
            return;
        }
        if (inb) store_record(record_at(a.out, a.out_pitch, px, py), rv, rw, 1u, 0.f);
        return;
    }
    float value = rv, weight = rw, wnew = 0.f;
    uint32_t reused_flag = reusing ? 1u : 0u;
    if (__any_sync(CHAOS_FULL_MASK, resample)) {
        float scf = advised;
        uint32_t s = sample_tile_sync<Real, FractalT>(a, fm, resample, px, py, scf, iters, nsamples, skipped);
        if (resample) {
            float wold = __fmul_rn(rw, 0.75f);
            weight = __fadd_rn(wold, scf);
            value = __fdiv_rn(__fmaf_rn(rv, wold, __fmul_rn(scf, __uint2float_rn(s))), weight);
            wnew = scf;
        }
    }
    if (__any_sync(CHAOS_FULL_MASK, fresh)) {
        float scf = advised < 1.f ? 1.f : advised;
        uint32_t s = sample_tile_sync<Real, FractalT>(a, fm, fresh, px, py, scf, iters, nsamples, skipped);
        if (fresh) { value = __uint2float_rn(s); weight = scf; }
    }
    if (inb) store_record(record_at(a.out, a.out_pitch, px, py), value, weight, reused_flag, wnew);
}

/*
 * pass R.  Memory-bound by design: per pixel 4 taps of 8 bytes gathered (about 16 B unique) and one 16-byte store.
 * To keep it that way the instruction count per pixel has to be small, so the warp does not walk 8x4 tiles:
 *   - a warp takes a strip of 4 horizontally adjacent vote tiles = 32 columns x 4 rows; lane = column;
 *   - the reprojected x of a pixel depends only on its column and the reprojected y only on its row
 *     (:194-203), so each lane computes ONE x (its column) and lanes 0..3 compute the four y, shared by shuffle:
 *     4 divisions per lane per strip instead of 16, same operations per value as the reference;
 *   - all 16 gathers of the lane's four pixels are issued before the first one is consumed;
 *   - foveation (:226-257) is only needed as the predicate "inside the focus area", a monotone function of the
 *     squared pixel distance d2 up to rounding; its ~150-instruction chain (sqrt, atanf, divisions) is evaluated
 *     only within +-2 % of the threshold d2, elsewhere the comparison with the threshold decides (libdevice atanf
 *     is accurate to 1 ulp and every other step to half an ulp, four orders of magnitude inside the guard band).
 * A vote tile in which any pixel needs a sample is appended to the worklist and left untouched for pass S.
 */
static __device__ __forceinline__ bool inside_focus_area(uint32_t x0, uint32_t y0, uint32_t fx, uint32_t fy, float max_ss, float d2_thr)
{
    float dx = __fsub_rn(__uint2float_rn(fx), __uint2float_rn(x0));
    float dy = __fsub_rn(__uint2float_rn(fy), __uint2float_rn(y0));
    float d2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
    if (d2 < 0.98f * d2_thr) return true;
    if (d2 > 1.02f * d2_thr) return false;
    float adv;
    bool inside;
    foveation(x0, y0, fx, fy, max_ss, adv, inside);
    return inside;
}

/* reprojected coordinate of column px (kY = false) or row py (kY = true), :194-203 */
template <class Real, bool kY> static __device__ __forceinline__ float warp_origin_1d(const chaos_render_args &a, uint32_t p);
template <> __device__ __forceinline__ float warp_origin_1d<double, false>(const chaos_render_args &a, uint32_t px)
{
    double rx = __ddiv_rn(__uint2double_rn(px), __uint2double_rn(a.width));
    double plane = __fma_rn(__dsub_rn(a.image[2], a.image[0]), rx, a.image[0]);
    double rel = __ddiv_rn(__dsub_rn(plane, a.image_reused[0]), __dsub_rn(a.image_reused[2], a.image_reused[0]));
    return __fmul_rn(__uint2float_rn(a.width), __double2float_rn(rel));
}
template <> __device__ __forceinline__ float warp_origin_1d<double, true>(const chaos_render_args &a, uint32_t py)
{
    double ry = __ddiv_rn(__uint2double_rn(a.height - py), __uint2double_rn(a.height));
    double plane = __fma_rn(__dsub_rn(a.image[3], a.image[1]), ry, a.image[1]);
    double rel = __ddiv_rn(__dsub_rn(plane, a.image_reused[1]), __dsub_rn(a.image_reused[3], a.image_reused[1]));
    float fh = __uint2float_rn(a.height);
    return __fmaf_rn(fh, -__double2float_rn(rel), fh);
}
template <> __device__ __forceinline__ float warp_origin_1d<float, false>(const chaos_render_args &a, uint32_t px)
{
    float fw = __uint2float_rn(a.width);
    float rx = __fdiv_rn(__uint2float_rn(px), fw);
    float plane = __fmaf_rn(__fsub_rn(a.imagef[2], a.imagef[0]), rx, a.imagef[0]);
    float rel = __fdiv_rn(__fsub_rn(plane, a.image_reusedf[0]), __fsub_rn(a.image_reusedf[2], a.image_reusedf[0]));
    return __fmul_rn(fw, rel);
}
template <> __device__ __forceinline__ float warp_origin_1d<float, true>(const chaos_render_args &a, uint32_t py)
{
    float fh = __uint2float_rn(a.height);
    float ry = __fdiv_rn(__uint2float_rn(a.height - py), fh);
    float plane = __fmaf_rn(__fsub_rn(a.imagef[3], a.imagef[1]), ry, a.imagef[1]);
    float rel = __fdiv_rn(__fsub_rn(plane, a.image_reusedf[1]), __fsub_rn(a.image_reusedf[3], a.image_reusedf[1]));
    return __fmaf_rn(fh, -rel, fh);
}

static __device__ __forceinline__ uint32_t colorize_sample_count(uint32_t cnt, uint32_t cnt100);

template <class Real>
static __device__ void advanced_reuse_pass(const chaos_render_args &a)
{
    /* fused colouring (chaos_render_args::fuse_rgba): the palette is staged in shared memory once per CTA, as compose does */
    extern __shared__ uint32_t s_fuse_palette[];
    const bool fuse = a.fuse_rgba != nullptr;
    if (fuse) {
        for (uint32_t k = threadIdx.x; k < a.fuse_palette_len; k += blockDim.x) s_fuse_palette[k] = a.fuse_palette[k];
        __syncthreads();
    }
    const bool visualize = VISUALIZE_SAMPLE_COUNT;
    const uint32_t cnt100 = __double2uint_rz(fmax((double)a.max_ss, 1.0));
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t fl = a.flags;
    const bool use_fov = (fl & CHAOS_FLAG_FOVEATION) && (fl & CHAOS_FLAG_IS_ZOOMING) && (fl & CHAOS_FLAG_ZOOMING_IN);
    const bool reuse = (fl & CHAOS_FLAG_SAMPLE_REUSE) != 0u;
    const bool resample_focus = use_fov && (fl & CHAOS_FLAG_ZOOMING_IN);
    const float d2_thr = a.focus_d2_thr;
    const uint32_t strips_x = (a.tiles_x + 3u) >> 2;
    const uint32_t rows_owned = a.n_tiles / a.tiles_x;
    const uint32_t n_strips = strips_x * rows_owned;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < n_strips; u += warps) {
        const uint32_t row = u / strips_x, g = u - row * strips_x;
        const uint32_t tile = row * a.tiles_x + 4u * g + (lane >> 3);      /* the lane's vote tile */
        const bool tile_ok = 4u * g + (lane >> 3) < a.tiles_x;
        uint32_t x0 = 0, y0 = 0;
        tile_origin(a, tile_ok ? tile : row * a.tiles_x, x0, y0);
        const uint32_t px = 32u * g + lane;
        const bool col_ok = px < a.width;
        const bool inside = resample_focus && tile_ok && inside_focus_area(x0, y0, a.focus_x, a.focus_y, a.max_ss, d2_thr);

        float ox = 0.f, oyv = 0.f;
        int oix = 0;
        if (reuse) {
            ox = warp_origin_1d<Real, false>(a, px);
            oyv = warp_origin_1d<Real, true>(a, y0 + (lane & 3u));          /* lanes 0..3 hold rows 0..3 */
            oix = __float2int_rz(roundf(ox));
        }
        const bool x_in = reuse && col_ok && !(oix < 2 || (uint32_t)oix >= a.width - 2u);
        const uint32_t i = __float2uint_rz(floorf(ox));
        const float al = __fsub_rn(ox, __uint2float_rn(i));

        bool tap[4];
        float be[4];
        float2 t00[4], t10[4], t01[4], t11[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float oy = __shfl_sync(CHAOS_FULL_MASK, oyv, r);
            const int oiy = __float2int_rz(roundf(oy));
            const uint32_t py = y0 + r;
            tap[r] = x_in && py < a.height && !(oiy < 2 || (uint32_t)oiy >= a.height - 2u);
            const uint32_t j = __float2uint_rz(floorf(oy));
            be[r] = __fsub_rn(oy, __uint2float_rn(j));
            t00[r] = t10[r] = t01[r] = t11[r] = make_float2(0.f, 0.f);
            if (tap[r]) {
                const char *row0 = previous_row(a, j) + (size_t)i * 16u;
                const char *row1 = previous_row(a, j + 1u) + (size_t)i * 16u;
                t00[r] = __ldg(reinterpret_cast<const float2 *>(row0));
                t10[r] = __ldg(reinterpret_cast<const float2 *>(row0 + 16));
                t01[r] = __ldg(reinterpret_cast<const float2 *>(row1));
                t11[r] = __ldg(reinterpret_cast<const float2 *>(row1 + 16));
            }
        }
        float rv[4], rw[4];
        bool needs = false;
        const float na = __fsub_rn(1.f, al);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float nb = __fsub_rn(1.f, be[r]);
            const float w00 = __fmul_rn(na, nb), w10 = __fmul_rn(al, nb), w01 = __fmul_rn(na, be[r]), w11 = __fmul_rn(al, be[r]);
            rv[r] = __fmaf_rn(w11, t11[r].x, __fmaf_rn(w01, t01[r].x, __fmaf_rn(w00, t00[r].x, __fmul_rn(w10, t10[r].x))));
            rw[r] = __fmaf_rn(w11, t11[r].y, __fmaf_rn(w01, t01[r].y, __fmaf_rn(w00, t00[r].y, __fmul_rn(w10, t10[r].y))));
            const bool in_frame = col_ok && (y0 + r) < a.height;
            const bool reusing = tap[r] && !((double)rw[r] < 0.1);
            needs |= in_frame && (!reusing || inside);
        }
        /* any pixel of an 8x4 vote tile needs a sample -> the whole tile goes to pass S */
        const uint32_t needs_mask = __ballot_sync(CHAOS_FULL_MASK, needs);
        const bool tile_needs = ((needs_mask >> (lane & 24u)) & 0xffu) != 0u;
        if (tile_needs) {
            if ((lane & 7u) == 0u && tile_ok) {
                a.tile_order[atomicAdd(&a.counters->bucket_count[0], 1u)] = tile;
                if (fuse) {          /* compose colours this tile after pass S */
                    const uint32_t gt = (y0 >> 2) * a.tiles_x + (x0 >> 3);
                    atomicOr(&a.late_tiles[gt >> 5], 1u << (gt & 31u));
                }
            }
        } else if (col_ok) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (y0 + r < a.height) {
                    store_record(record_at(a.out, a.out_pitch, px, y0 + r), rv[r], rw[r], 1u, 0.f);
                    if (fuse)        /* compose (:455-475) on the record just written: (value, weight, isReused = 1, 0) */
                        a.fuse_rgba[(size_t)(y0 + r) * a.width + px] =
                            visualize ? colorize_sample_count(0u, cnt100) : Fractal::colorize(s_fuse_palette, a.fuse_palette_len, rv[r]);
                }
        }
    }
}

template <class Real, class FractalT>
static __device__ void advanced_sample_pass(const chaos_render_args &a)
{
    frame_map<Real> fm;
    fm.init(a);
    const uint32_t lane = threadIdx.x & 31u;
    const bool use_fov = (a.flags & CHAOS_FLAG_FOVEATION) && (a.flags & CHAOS_FLAG_IS_ZOOMING) && (a.flags & CHAOS_FLAG_ZOOMING_IN);
    const uint32_t n_work = a.counters->bucket_count[0];       /* written by pass R, complete before this launch starts */
    unsigned long long iters = 0, nsamples = 0, skipped = 0;
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(&a.counters->next_tile_b, 1u);
        w = __shfl_sync(CHAOS_FULL_MASK, w, 0);
        if (w >= n_work) break;
        advanced_tile<Real, FractalT, 2>(a, fm, a.tile_order[w], lane, use_fov, iters, nsamples, skipped);
    }
    flush_counters(a, iters, nsamples, skipped);
}

/* ==========================================================================================
 * entry points
 * ======================================================================================== */
#include "render_refill.cuh"
#include "render_streams.cuh"

extern "C" __global__ void init() {}

/* engine 1 (default): lane-refill scheduler (render_refill.cuh).  One kernel per pass, each with its own register
 * allocation.  fractalRenderMain*: one sample per pixel, the whole frame in one launch; multi-sample frames run
 * chaosPassA* (sample 0), chaosPassB* (rounds with votes; dynamic shared memory = CHAOS_REFILL_SMEM_BYTES),
 * chaosPassC* (exported rounds as independent orbits) and chaosReplayExported. */
extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS, 4)
fractalRenderMainFloat(const __grid_constant__ chaos_render_args a) { render_main_independent<float, Fractal, 0>(a); }
extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS, 4)
fractalRenderMainDouble(const __grid_constant__ chaos_render_args a) { render_main_independent<double, Fractal, 0>(a); }
extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS, 4)
chaosPassAFloat(const __grid_constant__ chaos_render_args a) { render_main_independent<float, Fractal, 1>(a); }
extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS, 4)
chaosPassADouble(const __grid_constant__ chaos_render_args a) { render_main_independent<double, Fractal, 1>(a); }
extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS, 4)
chaosPassBFloat(const __grid_constant__ chaos_render_args a) { render_pass_b<float, Fractal>(a); }
extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS, 4)
chaosPassBDouble(const __grid_constant__ chaos_render_args a) { render_pass_b<double, Fractal>(a); }
extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS, 4)
chaosPassCFloat(const __grid_constant__ chaos_render_args a) { render_main_independent<float, Fractal, 2>(a); }
extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS, 4)
chaosPassCDouble(const __grid_constant__ chaos_render_args a) { render_main_independent<double, Fractal, 2>(a); }

/* engine 2: probe -> long -> finish (render_streams.cuh); args.phase says which pass the chain serves */
extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS)
chaosProbeFloat(const __grid_constant__ chaos_render_args a) { stream_probe<float, Fractal>(a); }
extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS)
chaosProbeDouble(const __grid_constant__ chaos_render_args a) { stream_probe<double, Fractal>(a); }
#ifndef CHAOS_LONG_MIN_BLOCKS
#define CHAOS_LONG_MIN_BLOCKS 4
#endif
extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS, CHAOS_LONG_MIN_BLOCKS)
chaosLongFloat(const __grid_constant__ chaos_render_args a) { stream_long<float, Fractal>(a); }
extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS, CHAOS_LONG_MIN_BLOCKS)
chaosLongDouble(const __grid_constant__ chaos_render_args a) { stream_long<double, Fractal>(a); }
extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS)
chaosFinishFloat(const __grid_constant__ chaos_render_args a) { stream_finish<float, Fractal>(a); }
extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS)
chaosFinishDouble(const __grid_constant__ chaos_render_args a) { stream_finish<double, Fractal>(a); }

/* engine 0: tile-synchronous, reference operation sequence; the differential check of engine 1 */
extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS)
fractalRenderMainFloatSync(const __grid_constant__ chaos_render_args a) { render_main_sync<float, Fractal>(a); }

extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS)
fractalRenderMainDoubleSync(const __grid_constant__ chaos_render_args a) { render_main_sync<double, Fractal>(a); }

/* between pass A and pass B of a two-pass render (render_refill.cuh) */
extern "C" __global__ void __launch_bounds__(256) chaosClassifyTiles(const __grid_constant__ chaos_render_args a) { classify_tiles(a); }
extern "C" __global__ void __launch_bounds__(256) chaosOrderTiles(const __grid_constant__ chaos_render_args a) { order_tiles(a); }
extern "C" __global__ void chaosWaitForeign(const __grid_constant__ chaos_render_args a) { wait_foreign(a); }
extern "C" __global__ void __launch_bounds__(256) chaosExportAll(const __grid_constant__ chaos_render_args a) { export_all_tiles(a); }
/* after pass C: the decisions of the tiles pass B exported */
extern "C" __global__ void __launch_bounds__(256) chaosReplayExported(const __grid_constant__ chaos_render_args a) { replay_exported(a); }

/* shared-memory need of the engine-1 kernels, read by the host at module load */
__constant__ uint32_t CHAOS_REFILL_SMEM = (uint32_t)CHAOS_REFILL_SMEM_BYTES;

/* fast frame: pass R (chaosReusePass*) then pass S (fractalRenderAdvanced*) */
extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS, 4)
chaosReusePassFloat(const __grid_constant__ chaos_render_args a) { advanced_reuse_pass<float>(a); }

extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS, 4)
chaosReusePassDouble(const __grid_constant__ chaos_render_args a) { advanced_reuse_pass<double>(a); }

extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS)
fractalRenderAdvancedFloat(const __grid_constant__ chaos_render_args a) { advanced_sample_pass<float, Fractal>(a); }

extern "C" __global__ void __launch_bounds__(CHAOS_RENDER_THREADS)
fractalRenderAdvancedDouble(const __grid_constant__ chaos_render_args a) { advanced_sample_pass<double, Fractal>(a); }

/* the reference's fractalRenderUnderSampled (:478-503) is looked up but never launched
 * (SURVEY.md 2.2); the symbol is kept so the module contract is complete */
extern "C" __global__ void fractalRenderUnderSampled(const __grid_constant__ chaos_render_args) {}

/* colorizeSampleCount (:67-77) */
static __device__ __forceinline__ uint32_t colorize_sample_count(uint32_t cnt, uint32_t cnt100)
{
    cnt = min(cnt, cnt100);
    float rel = __fdiv_rn(__uint2float_rn(cnt), __uint2float_rn(cnt100));
    uint32_t b = (uint32_t)__float2int_rz(__fmul_rn(rel, 255.f)) & 0xffu;
    return b | (b << 8) | (b << 16) | 0xff000000u;
}

/*
 * compose (:455-475): record -> RGBA8.  HBM-bound: 16 B read + 4 B written per pixel.  Each
 * thread handles 4 consecutive pixels: four 128-bit record loads, one 128-bit uchar4x4 store
 * (to device memory or straight into mapped pinned host memory); the palette is staged in
 * shared memory once per CTA.
 */
#define CHAOS_COMPOSE_THREADS 256
extern "C" __global__ void __launch_bounds__(CHAOS_COMPOSE_THREADS)
compose(const __grid_constant__ chaos_compose_args a)
{
    extern __shared__ uint32_t s_palette[];
    for (uint32_t k = threadIdx.x; k < a.palette_len; k += blockDim.x) s_palette[k] = a.palette[k];
    __syncthreads();
    const bool visualize = VISUALIZE_SAMPLE_COUNT;
    const uint32_t cnt100 = __double2uint_rz(fmax((double)a.max_ss, 1.0));
    const uint32_t quads_x = (a.width + 3u) >> 2;
    const uint64_t total = (uint64_t)quads_x * a.height;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t y = (uint32_t)(q / quads_x);
        uint32_t x = (uint32_t)(q - (uint64_t)y * quads_x) << 2;
        if (a.part_count > 1u && ((y / a.band_rows) % a.part_count) != a.part_index) continue;
        if (a.only_tiles) {      /* a quad of pixels lies within one 8x4 vote tile */
            const uint32_t gt = (y >> 2) * a.tiles_x + (x >> 3);
            if (!((a.only_tiles[gt >> 5] >> (gt & 31u)) & 1u)) continue;
        }
        const float4 *src = reinterpret_cast<const float4 *>((const char *)a.in + (size_t)y * a.in_pitch) + x;
        uint32_t n = min(4u, a.width - x);
        float4 rec[4];
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k)
            if (k < n) rec[k] = __ldcs(src + k);              /* streaming: records are not re-read by this pass */
        uint32_t col[4];
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
            if (k >= n) { col[k] = 0; continue; }
            if (visualize) {
                col[k] = colorize_sample_count(__float2uint_rz(rec[k].y), cnt100);
                if (__float_as_uint(rec[k].z) & 0xffu)
                    col[k] = colorize_sample_count(__float2uint_rz(rec[k].w), cnt100);
            } else {
                col[k] = Fractal::colorize(s_palette, a.palette_len, rec[k].x);
            }
        }
        uint32_t *dst = a.out_rgba + (size_t)y * a.width + x;
        if (n == 4u && ((a.width & 3u) == 0u)) {
            *reinterpret_cast<uint4 *>(dst) = make_uint4(col[0], col[1], col[2], col[3]);
        } else {
            for (uint32_t k = 0; k < n; ++k) dst[k] = col[k];
        }
    }
}

extern "C" __global__ void debug()                        /* :505-513 */
{
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) Fractal::debugFractal();
}

#endif
