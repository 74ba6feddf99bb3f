/*
 * mandelbrot module: z <- z^2 + c, z0 = 0, c = pixel.
 * Same results as the reference module src/main/cuda/fractals/mandelbrot.cu:11-39 as built by
 * nvcc 12.9 for sm_100a; the escape loop is in quadratic.cuh.
 */
#include "../quadratic.cuh"

struct Fractal {
    template <class Real> struct Orbit {
        static constexpr bool kResumable = true;
        quadratic_orbit<Real> q;
        __device__ __forceinline__ void start(Real px, Real py, const orbit_ctx &ctx) { q.init((Real)0, (Real)0, px, py, ctx); }
        __device__ __forceinline__ void force_exact() { q.force_exact(); }
        __device__ __forceinline__ bool run(uint32_t &i, uint32_t limit, bool tested) { return q.run(i, limit, tested); }
        __device__ __forceinline__ bool wants_tested() const { return q.wants_tested(); }
        __device__ __forceinline__ uint32_t skipped() const { return q.skipped(); }
        __device__ __forceinline__ void save(Real &x, Real &y) const { q.save(x, y); }
        __device__ __forceinline__ void resume(Real x, Real y) { q.resume(x, y); }
        /* mandelbrot.cu:22-24: points that never left report 0 */
        __device__ __forceinline__ uint32_t finish(uint32_t i, uint32_t max_iterations) const
        {
            return i == max_iterations ? 0u : __float2uint_rz(__uint2float_rn(i));
        }
    };
    static __device__ __forceinline__ uint32_t colorize(const uint32_t *palette, uint32_t len, float result)
    {
        return chaos_default_colorize(palette, len, result);
    }
    static __device__ void debugFractal() { printf("hello from mandelbrot\n"); }
};

#include "../render_generic.cuh"
