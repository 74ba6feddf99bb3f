/*
 * chaos_compat_pre.cuh -- source-level compatibility with the reference's per-fractal module contract
 * (src/main/cuda/fractals/fractal.cuh:7-28, built by src/main/cuda/compile.sh:10-15).
 *
 * A module written for the reference is ONE file that includes "fractal.cuh" and defines three free functions:
 *     template <class Real> __device__ float computeFractal(unsigned int maxIterations, Point<Real> z);
 *     __device__ unsigned int colorize(cudaSurfaceObject_t colorPalette, unsigned int paletteLength, float result);
 *     __device__ void debugFractal();
 * build.py compiles such a file, UNMODIFIED, into a module of this backend from a three-line translation unit
 * (the counterpart of the one compile.sh writes):
 *     #include "compat/chaos_compat_pre.cuh"      <- this file: the types helpers.cuh gives a module author
 *     #include "<path>/<name>.cu"                 <- the author's file
 *     #include "compat/chaos_compat_post.cuh"     <- adapts the three functions to `struct Fractal`
 * Its own `#include "fractal.cuh"` finds either compat/fractal.cuh (include path) or, when the file lies in the
 * reference's tree, the reference's header -- whose helpers.cuh is then skipped through its include guard, so the
 * types below are the ones in use either way.
 *
 * What an author can rely on: Point<T> with the operators of helpers.cuh:22-77, ColorsRGBA (:150-162), color_t,
 * ASSERT, and surf2Dread() on the palette.  The palette is not a surface here (compose stages it in shared memory):
 * the handle passed to colorize() carries the palette's address and surf2Dread is redirected to a plain load.
 * computeFractal() is an opaque call, so such a module runs through ClassicOrbit (fractal.cuh): correct records and
 * colours; the lane-refill machinery has nothing to suspend, and the work counters count samples but not trips.
 */
#ifndef CHAOS_COMPAT_PRE_CUH
#define CHAOS_COMPAT_PRE_CUH

#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#define HELPERS                 /* the reference's helpers.cuh is superseded by this header */
#define DEBUG_MODE
#define ASSERT(x) assert(x)
#define __ALL(predicate) __all_sync(__activemask(), predicate)
#define __ANY(predicate) __any_sync(__activemask(), predicate)

/* helpers.cuh:22-77 -- component-wise arithmetic on a pair */
#define CHAOS_COMPAT_POINT_OP(op)                                                                       \
    __device__ Point<T> operator op(const Point<T> &o) { return Point<T>(x op o.x, y op o.y); }         \
    __device__ const Point<T> operator op(const Point<T> &o) const { return Point<T>(x op o.x, y op o.y); }
template <class T> struct Point {
    T x, y;
    __device__ Point() {}
    __device__ Point(const T px, const T py) : x(px), y(py) {}
    __device__ Point(const T both) : x(both), y(both) {}
    CHAOS_COMPAT_POINT_OP(+)
    CHAOS_COMPAT_POINT_OP(-)
    CHAOS_COMPAT_POINT_OP(*)
    CHAOS_COMPAT_POINT_OP(/)
    CHAOS_COMPAT_POINT_OP(%)
    __device__ bool operator==(const Point<T> &o) const { return x == o.x && y == o.y; }
    __device__ bool operator!=(const Point<T> &o) const { return x != o.x || y != o.y; }
    __device__ T manhattanDistanceTo(const Point<T> &o) const { return abs(x - o.x) + abs(y - o.y); }
    __device__ T distanceTo(const Point<T> &o) const { return sqrt((x - o.x) * (x - o.x) + (y - o.y) * (y - o.y)); }
    template <class S> __device__ Point<S> cast() { return Point<S>((S)x, (S)y); }
    template <class S> __device__ const Point<S> cast() const { return Point<S>((S)x, (S)y); }
};
#undef CHAOS_COMPAT_POINT_OP

/* helpers.cuh:132-146 */
struct rgba { char r, g, b, a; };
typedef struct color_t {
    union {
        unsigned int intValue;
        struct rgba rgba;
    };
} color_t;

/* helpers.cuh:150-162 -- R in the low byte */
class ColorsRGBA {
public:
    static constexpr const unsigned int BLACK = 0xff000000, WHITE = 0xffffffff, PINK = 0xffb469ff, GOLD = 0xff00d7ff;
    static constexpr const unsigned int YELLOW = 0xff00ffff, BLUE = 0xffff0000, GREEN = 0xff00ff00, RED = 0xff0000ff;
};

/* the palette "surface": one row of RGBA8; the handle is the address of the staged palette */
template <class T>
static __device__ __forceinline__ void chaos_compat_surf2Dread(T *out, cudaSurfaceObject_t palette, int x_bytes, int /* y: the palette has one row */)
{
    *out = *reinterpret_cast<const T *>(reinterpret_cast<const char *>((uintptr_t)palette) + x_bytes);
}
#define surf2Dread chaos_compat_surf2Dread

#endif
