/*
 * julia module: z <- z^2 + c, z0 = pixel, c = julia_c (host-written constant).
 * Same results as src/main/cuda/fractals/julia.cu:3-31 (nvcc 12.9, sm_100a); inside points
 * report maxIterations (not 0).  Host half: modules/ModuleJulia.java:9-49.
 */
#include "../quadratic.cuh"

__constant__ double julia_c[2];

struct Fractal {
    template <class Real> struct Orbit {
        typedef real_ops<Real> op;
        static constexpr bool kResumable = true;
        quadratic_orbit<Real> q;
        __device__ __forceinline__ void start(Real px, Real py, const orbit_ctx &ctx)
        {
            q.init(px, py, op::from_f64(julia_c[0]), op::from_f64(julia_c[1]), ctx);   /* julia.cu:7 */
        }
        __device__ __forceinline__ void force_exact() { q.force_exact(); }
        __device__ __forceinline__ bool run(uint32_t &i, uint32_t limit, bool tested) { return q.run(i, limit, tested); }
        __device__ __forceinline__ bool wants_tested() const { return q.wants_tested(); }
        __device__ __forceinline__ uint32_t skipped() const { return q.skipped(); }
        __device__ __forceinline__ void save(Real &x, Real &y) const { q.save(x, y); }
        __device__ __forceinline__ void resume(Real x, Real y) { q.resume(x, y); }
        __device__ __forceinline__ uint32_t finish(uint32_t i, uint32_t) const
        {
            return __float2uint_rz(__uint2float_rn(i));
        }
    };
    static __device__ __forceinline__ uint32_t colorize(const uint32_t *palette, uint32_t len, float result)
    {
        return chaos_default_colorize(palette, len, result);
    }
    static __device__ void debugFractal() { printf("hello from julia\n"); }
};

#include "../render_generic.cuh"
