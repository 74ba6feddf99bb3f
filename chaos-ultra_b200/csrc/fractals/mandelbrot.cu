/*
 * mandelbrot module: z <- z^2 + c, z0 = 0, c = pixel.
 * Same results as the reference module src/main/cuda/fractals/mandelbrot.cu:11-39 as built by
 * nvcc 12.9 for sm_100a: per trip  xx=rn(x*x); yy=rn(y*y); leave unless rn(xx+yy) < 4;
 * xn=rn(cx+rn(xx-yy)); y=fma(rn(x+x), y, cy); x=xn   (SURVEY.md 8a row 1).
 */
#include "../fractal.cuh"

struct Fractal {
    template <class Real> struct Orbit {
        typedef real_ops<Real> op;
        static constexpr bool kResumable = true;
        Real x, y, cx, cy;
        __device__ __forceinline__ void start(Real px, Real py)
        {
            cx = px; cy = py;
            x = (Real)0; y = (Real)0;
        }
        __device__ __forceinline__ bool run(uint32_t &i, uint32_t limit)
        {
            while (i < limit) {
                Real xx = op::mul(x, x);
                Real yy = op::mul(y, y);
                if (!op::below4(op::add(xx, yy))) return true;
                Real xn = op::add(cx, op::sub(xx, yy));
                y = op::fma(op::add(x, x), y, cy);
                x = xn;
                ++i;
            }
            return false;
        }
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
