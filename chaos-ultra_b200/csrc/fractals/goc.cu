/*
 * "goc": z <- 42 z^-2 + c^7 for exactly maxIterations steps, value = |z|.  A toy of the reference
 * (src/main/cuda/fractals/goc.cu:8-44, modules/ModuleGoci.java:9-28), kept because the provider registers it.
 */
#include <thrust/complex.h>
#include "../fractal.cuh"

__constant__ int amplifier;

struct GocImpl {
    template <class Real> static __device__ __forceinline__ thrust::complex<Real> step(thrust::complex<Real> c, thrust::complex<Real> z)
    {
        return 42 * 1 / (z * z) + (c * c * c * c * c * c * c);
    }
    template <class Real> static __device__ float compute(uint32_t maxIterations, Real px, Real py, uint32_t &trips)
    {
        trips = 0;
        if (px == 0 || py == 0) return 0;
        thrust::complex<Real> z = thrust::complex<Real>(px, py);
        thrust::complex<Real> c = z;
        unsigned int i = 0;
        while (i < maxIterations) {
            z = step<Real>(c, z);
            i++;
        }
        trips = i;
        Real dx = z.real() - 0, dy = z.imag() - 0;
        return sqrt(dx * dx + dy * dy);
    }
};

struct Fractal {
    /* no branch separates c.y's multiply and subtract in the reference build of this module: ptxas contracts them
     * into one FMA (SASS of oracle/_ref/goc.src.cubin), see frame_map::plane_point */
    static constexpr bool kFusedPlaneY = true;
    template <class Real> using Orbit = ClassicOrbit<GocImpl, Real>;
    static __device__ __forceinline__ uint32_t colorize(const uint32_t *palette, uint32_t len, float result)
    {
        uint32_t colour = chaos_default_colorize(palette, len, result);
        if (result < 35) colour &= 0xff00ffffu;
        return colour;
    }
    static __device__ void debugFractal() { printf("hello from goci\n"); }
};

#include "../render_generic.cuh"
