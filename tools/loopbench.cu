/*
 * tools/loopbench.cu -- measurement only: how fast the FP64 escape loop of quadratic.cuh can go in isolation.
 * Every lane runs one (or two) bounded, non-closing orbits for a fixed number of trips; no scheduling, no
 * stores in the loop.  Variants: test per trip (6 ops), deferred test in groups of 8/16 (5 ops), with or without
 * the recurrence compare, one or two orbits per lane.  Swept over resident warps per SM sub-partition.
 *
 *   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I chaos-ultra_b200/csrc tools/loopbench.cu -o /tmp/loopbench
 */
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct orb {
    double x, y, cx, cy, sx, sy;
    uint32_t next_save;
};

__device__ __forceinline__ bool below16(double s) { return (uint32_t)__double2hiint(s) < 0x40300000u; }
__device__ __forceinline__ bool same(double a, double b)
{
    return ((__double2hiint(a) ^ __double2hiint(b)) | (__double2loint(a) ^ __double2loint(b))) == 0;
}
__device__ __forceinline__ bool step(orb &o)
{
    double xx = __dmul_rn(o.x, o.x), yy = __dmul_rn(o.y, o.y);
    bool ok = below16(__dadd_rn(xx, yy));
    double xn = __fma_rn(__dsub_rn(xx, yy), 0.5, o.cx);
    o.y = __fma_rn(o.x, o.y, o.cy);
    o.x = xn;
    return ok;
}
__device__ __forceinline__ void adv(orb &o, double &xx, double &yy)
{
    double xn = __fma_rn(__dsub_rn(xx, yy), 0.5, o.cx);
    o.y = __fma_rn(o.x, o.y, o.cy);
    o.x = xn;
    xx = __dmul_rn(o.x, o.x);
    yy = __dmul_rn(o.y, o.y);
}

/* V: 0 tested x8; 1 deferred; 2 deferred + recurrence compare.  G: group length.  U: orbits per lane. */
template <int V, int G, int U> __global__ void __launch_bounds__(256) loop_kernel(double *out, uint32_t trips, double seed)
{
    orb o[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        o[u].cx = 2.0 * (-1.8 + 1e-7 * (threadIdx.x + 256 * u) + seed * 1e-9);   /* chaotic real parameter: bounded, never closes */
        o[u].cy = 0.0;
        o[u].x = 0.0; o[u].y = 0.0;
        o[u].sx = o[u].sy = 123.0;
        o[u].next_save = 8u;
    }
    uint32_t i = 0;
    uint32_t flag = 0;
    if (V == 0) {
        while (i + 8u <= trips) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int u = 0; u < U; ++u) ok &= step(o[u]);
            if (!ok) { flag = 1; break; }
            i += 8u;
        }
    } else {
        double xx[U], yy[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { xx[u] = __dmul_rn(o[u].x, o[u].x); yy[u] = __dmul_rn(o[u].y, o[u].y); }
        while (i + G <= trips) {
            double bx[U], by[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { bx[u] = o[u].x; by[u] = o[u].y; }
#pragma unroll 1
            for (int r = 0; r < G / 8; ++r)
#pragma unroll
                for (int k = 0; k < 8; ++k)
#pragma unroll
                    for (int u = 0; u < U; ++u) adv(o[u], xx[u], yy[u]);
            bool ok = true;
#pragma unroll
            for (int u = 0; u < U; ++u) ok &= below16(__dadd_rn(xx[u], yy[u]));
            if (!ok) {
#pragma unroll
                for (int u = 0; u < U; ++u) { o[u].x = bx[u]; o[u].y = by[u]; }
                flag = 1;
                break;
            }
            i += G;
            if (V == 2) {
                bool hit = false;
#pragma unroll
                for (int u = 0; u < U; ++u) hit |= same(o[u].x, o[u].sx) && same(o[u].y, o[u].sy);
                if (hit) { flag = 2; break; }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (i >= o[u].next_save) { o[u].sx = o[u].x; o[u].sy = o[u].y; o[u].next_save = i + max(8u, (i >> 2) & ~7u); }
            }
        }
    }
    double acc = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) acc += o[u].x + o[u].y + o[u].sx;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + flag + i;
}

template <int V, int G, int U> static void run(const char *name, double ops_per_trip, int sms, double *out, double clock_ghz)
{
    const uint32_t trips = 40000;
    /* resident warps per SM sub-partition: blocks of 128 threads (one warp per sub-partition each) */
    const int warps_list[] = {1, 2, 4, 6, 8, 10, 12, 16};
    printf("%-34s", name);
    for (int w : warps_list) {
        int max_blocks = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, loop_kernel<V, G, U>, 128, 0));
        if (w > max_blocks) { printf("      -   "); continue; }
        int blocks = sms * w;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        loop_kernel<V, G, U><<<blocks, 128>>>(out, trips, 1.0);
        CK(cudaDeviceSynchronize());
        float best = 1e9f;
        for (int r = 0; r < 3; ++r) {
            CK(cudaEventRecord(e0));
            loop_kernel<V, G, U><<<blocks, 128>>>(out, trips, 2.0 + r);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        double lane_trips = (double)blocks * 128 * trips * U;
        double tps = lane_trips / (best * 1e-3);                 /* orbit trips per second, whole GPU */
        double cyc_per_trip = best * 1e-3 * clock_ghz * 1e9 / trips;   /* per warp, per round of U trips */
        printf(" %5.2fT/%3.0f", tps / 1e12, cyc_per_trip);
        (void)ops_per_trip;
    }
    printf("\n");
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    double ghz = khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal; columns: warps per SM sub-partition 1 2 4 6 8 10 12 16; cell = T orbit-trips/s / cycles per warp per trip-round\n",
           p.name, p.multiProcessorCount, ghz);
    double *out;
    CK(cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 16 * 128));
    int sms = p.multiProcessorCount;
    run<0, 8, 1>("tested x8 (6 ops)", 6, sms, out, ghz);
    run<1, 8, 1>("deferred x8 (5.125)", 5.125, sms, out, ghz);
    run<1, 16, 1>("deferred x16 (5.06)", 5.0625, sms, out, ghz);
    run<1, 32, 1>("deferred x32", 5.03, sms, out, ghz);
    run<2, 8, 1>("deferred x8 + recurrence", 5.125, sms, out, ghz);
    run<2, 16, 1>("deferred x16 + recurrence", 5.0625, sms, out, ghz);
    run<2, 32, 1>("deferred x32 + recurrence", 5.03, sms, out, ghz);
    run<0, 8, 2>("tested x8, 2 orbits/lane", 6, sms, out, ghz);
    run<1, 8, 2>("deferred x8, 2 orbits/lane", 5.125, sms, out, ghz);
    run<2, 8, 2>("deferred x8 + rec, 2 orbits/lane", 5.125, sms, out, ghz);
    run<2, 16, 2>("deferred x16 + rec, 2 orbits/lane", 5.0625, sms, out, ghz);
    return 0;
}
