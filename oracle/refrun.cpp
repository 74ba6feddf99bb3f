/*
 * oracle/refrun.cpp -- runs the REFERENCE's own kernels on the GPU.  TEST INFRASTRUCTURE ONLY.
 *
 * Loads a module built by oracle/Makefile from the reference's sources where they lie
 * (oracle/_ref/<name>.src.cubin: src/main/cuda/fractals/<name>.cu + fractalRendererGeneric.cu
 * compiled by nvcc 12.9 for sm_100a; oracle/_ref/<name>.ptx92.cubin: the shipped CUDA-9.2 PTX
 * assembled by ptxas for sm_100a) and launches it exactly like the Java host does:
 *   block 32x32, grid ceil(W/32) x ceil(H/32), NULL stream  (CudaFractalRenderer.java:36-37,242-243,371-378)
 *   two pitched buffers of 16-byte records, cuMemAllocPitch element size 16 (DeviceMemoryDoubleBuffer2D.java:134)
 *   positional parameters in the order of RenderingKernel.java:21-25, KernelMain.java:19-20,
 *   KernelAdvanced.java:26-29, KernelCompose.java:24-33; `long pitch` passed as 8 bytes little-endian
 *   float kernels get the four segment doubles cast to float on the host (KernelMainFloat.java:19-21)
 * The GL textures of the reference become two CUDA arrays (RGBA8, surface load/store) wrapped in
 * surface objects for `compose`.
 *
 * Used by tests/ (-m gpu) as the strongest parity check, by tests/golden/make_golden.py to produce
 * the committed fixtures, and by bench.py to time the reference's kernels on the B200.
 * Links against libcuda (stub at build time); only loadable on a machine with a driver.
 */
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

static char g_err[512] = "";
#define RR_TRY(call)                                                              \
    do {                                                                          \
        CUresult _r = (call);                                                     \
        if (_r != CUDA_SUCCESS) {                                                 \
            const char *_n = nullptr;                                             \
            cuGetErrorName(_r, &_n);                                              \
            snprintf(g_err, sizeof g_err, "%s -> %s", #call, _n ? _n : "?");      \
            return -1;                                                            \
        }                                                                         \
    } while (0)

struct refrun {
    CUdevice dev;
    CUcontext ctx;
    CUmodule mod;
    CUfunction main_f, main_d, adv_f, adv_d, compose;
};

extern "C" const char *refrun_last_error(void) { return g_err; }

extern "C" int refrun_open(const char *module_path, int device, refrun **out)
{
    *out = nullptr;
    RR_TRY(cuInit(0));
    refrun *r = new refrun();
    RR_TRY(cuDeviceGet(&r->dev, device));
    RR_TRY(cuDevicePrimaryCtxRetain(&r->ctx, r->dev));
    RR_TRY(cuCtxPushCurrent(r->ctx));
    CUresult e = cuModuleLoad(&r->mod, module_path);
    if (e == CUDA_SUCCESS) e = cuModuleGetFunction(&r->main_f, r->mod, "fractalRenderMainFloat");
    if (e == CUDA_SUCCESS) e = cuModuleGetFunction(&r->main_d, r->mod, "fractalRenderMainDouble");
    if (e == CUDA_SUCCESS) e = cuModuleGetFunction(&r->adv_f, r->mod, "fractalRenderAdvancedFloat");
    if (e == CUDA_SUCCESS) e = cuModuleGetFunction(&r->adv_d, r->mod, "fractalRenderAdvancedDouble");
    if (e == CUDA_SUCCESS) e = cuModuleGetFunction(&r->compose, r->mod, "compose");
    CUcontext c;
    cuCtxPopCurrent(&c);
    if (e != CUDA_SUCCESS) {
        const char *n = nullptr;
        cuGetErrorName(e, &n);
        snprintf(g_err, sizeof g_err, "loading %s -> %s", module_path, n ? n : "?");
        delete r;
        return -1;
    }
    *out = r;
    return 0;
}

extern "C" int refrun_close(refrun *r)
{
    if (!r) return 0;
    cuCtxPushCurrent(r->ctx);
    cuModuleUnload(r->mod);
    CUcontext c;
    cuCtxPopCurrent(&c);
    cuDevicePrimaryCtxRelease(r->dev);
    delete r;
    return 0;
}

extern "C" int refrun_write_constant(refrun *r, const char *name, const void *data, size_t bytes)
{
    RR_TRY(cuCtxPushCurrent(r->ctx));
    CUdeviceptr p;
    size_t sz;
    CUresult e = cuModuleGetGlobal(&p, &sz, r->mod, name);
    if (e == CUDA_SUCCESS && sz >= bytes) e = cuMemcpyHtoD(p, data, bytes);
    CUcontext c;
    cuCtxPopCurrent(&c);
    if (e != CUDA_SUCCESS) { snprintf(g_err, sizeof g_err, "constant %s: error %d", name, (int)e); return -1; }
    return 0;
}

struct ctx_scope {
    explicit ctx_scope(CUcontext c) { cuCtxPushCurrent(c); }
    ~ctx_scope() { CUcontext c; cuCtxPopCurrent(&c); }
};

static int launch_timed(CUfunction fn, uint32_t W, uint32_t H, void **params, int reps, float *ms)
{
    unsigned gx = (W + 31) / 32, gy = (H + 31) / 32;
    CUevent e0, e1;
    RR_TRY(cuEventCreate(&e0, CU_EVENT_DEFAULT));
    RR_TRY(cuEventCreate(&e1, CU_EVENT_DEFAULT));
    float best = 1e30f;
    for (int i = 0; i < (reps < 1 ? 1 : reps); ++i) {
        RR_TRY(cuEventRecord(e0, 0));
        RR_TRY(cuLaunchKernel(fn, gx, gy, 1, 32, 32, 1, 0, 0, params, nullptr));
        RR_TRY(cuEventRecord(e1, 0));
        RR_TRY(cuCtxSynchronize());
        float t = 0;
        cuEventElapsedTime(&t, e0, e1);
        if (t < best) best = t;
    }
    cuEventDestroy(e0);
    cuEventDestroy(e1);
    if (ms) *ms = best;
    return 0;
}

/* fractalRenderMain{Float,Double}.  out_records: host, W*H*16 bytes (pitch removed). */
extern "C" int refrun_main(refrun *r, int is_double, uint32_t W, uint32_t H, const double image[4], uint32_t maxIter,
                           float maxSS, uint32_t flags, void *out_records, int reps, float *ms)
{
    ctx_scope s(r->ctx);
    CUdeviceptr out;
    size_t pitch;
    RR_TRY(cuMemAllocPitch(&out, &pitch, (size_t)W * 16, H, 16));
    RR_TRY(cuMemsetD8(out, 0, pitch * H));
    long long pitch_ll = (long long)pitch;
    uint32_t size[2] = {W, H};
    float imf[4] = {(float)image[0], (float)image[1], (float)image[2], (float)image[3]};
    void *params[7] = {&out, &pitch_ll, size, is_double ? (void *)image : (void *)imf, &maxIter, &maxSS, &flags};
    int rc = launch_timed(is_double ? r->main_d : r->main_f, W, H, params, reps, ms);
    if (rc == 0 && out_records) {
        CUDA_MEMCPY2D c;
        memset(&c, 0, sizeof c);
        c.srcMemoryType = CU_MEMORYTYPE_DEVICE; c.srcDevice = out; c.srcPitch = pitch;
        c.dstMemoryType = CU_MEMORYTYPE_HOST; c.dstHost = out_records; c.dstPitch = (size_t)W * 16;
        c.WidthInBytes = (size_t)W * 16; c.Height = H;
        CUresult e = cuMemcpy2D(&c);
        if (e != CUDA_SUCCESS) { snprintf(g_err, sizeof g_err, "cuMemcpy2D error %d", (int)e); rc = -1; }
    }
    cuMemFree(out);
    return rc;
}

/* fractalRenderAdvanced{Float,Double}.  in_records/out_records: host, W*H*16 bytes. */
extern "C" int refrun_advanced(refrun *r, int is_double, uint32_t W, uint32_t H, const double image[4], uint32_t maxIter,
                               float maxSS, uint32_t flags, const double image_reused[4], const void *in_records,
                               uint32_t focus_x, uint32_t focus_y, void *out_records, int reps, float *ms)
{
    ctx_scope s(r->ctx);
    CUdeviceptr out, in;
    size_t pitch, in_pitch;
    RR_TRY(cuMemAllocPitch(&out, &pitch, (size_t)W * 16, H, 16));
    RR_TRY(cuMemAllocPitch(&in, &in_pitch, (size_t)W * 16, H, 16));
    RR_TRY(cuMemsetD8(out, 0, pitch * H));
    CUDA_MEMCPY2D c;
    memset(&c, 0, sizeof c);
    c.srcMemoryType = CU_MEMORYTYPE_HOST; c.srcHost = in_records; c.srcPitch = (size_t)W * 16;
    c.dstMemoryType = CU_MEMORYTYPE_DEVICE; c.dstDevice = in; c.dstPitch = in_pitch;
    c.WidthInBytes = (size_t)W * 16; c.Height = H;
    RR_TRY(cuMemcpy2D(&c));
    long long pitch_ll = (long long)pitch, in_pitch_ll = (long long)in_pitch;
    uint32_t size[2] = {W, H}, focus[2] = {focus_x, focus_y};
    float imf[4] = {(float)image[0], (float)image[1], (float)image[2], (float)image[3]};
    float oldf[4] = {(float)image_reused[0], (float)image_reused[1], (float)image_reused[2], (float)image_reused[3]};
    void *params[11] = {&out, &pitch_ll, size, is_double ? (void *)image : (void *)imf, &maxIter, &maxSS, &flags,
                        is_double ? (void *)image_reused : (void *)oldf, &in, &in_pitch_ll, focus};
    int rc = launch_timed(is_double ? r->adv_d : r->adv_f, W, H, params, reps, ms);
    if (rc == 0 && out_records) {
        memset(&c, 0, sizeof c);
        c.srcMemoryType = CU_MEMORYTYPE_DEVICE; c.srcDevice = out; c.srcPitch = pitch;
        c.dstMemoryType = CU_MEMORYTYPE_HOST; c.dstHost = out_records; c.dstPitch = (size_t)W * 16;
        c.WidthInBytes = (size_t)W * 16; c.Height = H;
        CUresult e = cuMemcpy2D(&c);
        if (e != CUDA_SUCCESS) { snprintf(g_err, sizeof g_err, "cuMemcpy2D error %d", (int)e); rc = -1; }
    }
    cuMemFree(out);
    cuMemFree(in);
    return rc;
}

static int make_surface(uint32_t w, uint32_t h, CUarray *arr, CUsurfObject *surf)
{
    CUDA_ARRAY3D_DESCRIPTOR d;
    memset(&d, 0, sizeof d);
    d.Width = w; d.Height = h; d.Depth = 0;
    d.Format = CU_AD_FORMAT_UNSIGNED_INT8; d.NumChannels = 4;
    d.Flags = CUDA_ARRAY3D_SURFACE_LDST;
    RR_TRY(cuArray3DCreate(arr, &d));
    CUDA_RESOURCE_DESC rd;
    memset(&rd, 0, sizeof rd);
    rd.resType = CU_RESOURCE_TYPE_ARRAY;
    rd.res.array.hArray = *arr;
    RR_TRY(cuSurfObjectCreate(surf, &rd));
    return 0;
}

/* compose.  records: host W*H*16; palette: R in the low byte; out_rgba: host W*H*4. */
extern "C" int refrun_compose(refrun *r, uint32_t W, uint32_t H, const void *records, const uint32_t *palette,
                              uint32_t palette_len, float maxSS, int visualize, uint32_t *out_rgba, int reps, float *ms)
{
    ctx_scope s(r->ctx);
    unsigned char vis = visualize ? 1 : 0;
    if (refrun_write_constant(r, "VISUALIZE_SAMPLE_COUNT", &vis, 1) != 0) return -1;
    CUdeviceptr in;
    size_t pitch;
    RR_TRY(cuMemAllocPitch(&in, &pitch, (size_t)W * 16, H, 16));
    CUDA_MEMCPY2D c;
    memset(&c, 0, sizeof c);
    c.srcMemoryType = CU_MEMORYTYPE_HOST; c.srcHost = records; c.srcPitch = (size_t)W * 16;
    c.dstMemoryType = CU_MEMORYTYPE_DEVICE; c.dstDevice = in; c.dstPitch = pitch;
    c.WidthInBytes = (size_t)W * 16; c.Height = H;
    RR_TRY(cuMemcpy2D(&c));
    CUarray a_out, a_pal;
    CUsurfObject s_out, s_pal;
    if (make_surface(W, H, &a_out, &s_out)) return -1;
    if (make_surface(palette_len, 1, &a_pal, &s_pal)) return -1;
    memset(&c, 0, sizeof c);
    c.srcMemoryType = CU_MEMORYTYPE_HOST; c.srcHost = palette; c.srcPitch = (size_t)palette_len * 4;
    c.dstMemoryType = CU_MEMORYTYPE_ARRAY; c.dstArray = a_pal;
    c.WidthInBytes = (size_t)palette_len * 4; c.Height = 1;
    RR_TRY(cuMemcpy2D(&c));
    long long pitch_ll = (long long)pitch;
    CUdeviceptr bcg = in;                       /* inputBcg is never read by the kernel */
    void *params[10] = {&in, &pitch_ll, &bcg, &pitch_ll, &s_out, &W, &H, &s_pal, &palette_len, &maxSS};
    int rc = launch_timed(r->compose, W, H, params, reps, ms);
    if (rc == 0 && out_rgba) {
        memset(&c, 0, sizeof c);
        c.srcMemoryType = CU_MEMORYTYPE_ARRAY; c.srcArray = a_out;
        c.dstMemoryType = CU_MEMORYTYPE_HOST; c.dstHost = out_rgba; c.dstPitch = (size_t)W * 4;
        c.WidthInBytes = (size_t)W * 4; c.Height = H;
        CUresult e = cuMemcpy2D(&c);
        if (e != CUDA_SUCCESS) { snprintf(g_err, sizeof g_err, "cuMemcpy2D(array) error %d", (int)e); rc = -1; }
    }
    cuSurfObjectDestroy(s_out);
    cuSurfObjectDestroy(s_pal);
    cuArrayDestroy(a_out);
    cuArrayDestroy(a_pal);
    cuMemFree(in);
    return rc;
}

/*
 * Frame loop of the reference host, for timing (bench.py --impl reference): per frame, exactly the sequence of
 * CudaFractalRenderer.renderQuality :187-206 -- main kernel, cuCtxSynchronize twice (:259-261), two surface objects
 * created, compose launched, cuCtxSynchronize, surfaces destroyed (:279-352) -- minus the GL map/unmap, which has
 * no equivalent without a GL context.  With to_host != 0 the composed CUDA array is then copied to pinned host
 * memory, so the timed region ends where this backend's HOST output mode ends.
 * Returns wall milliseconds for `steps` frames (after `warmup` untimed ones) and the event-timed kernel sums.
 */
#include <chrono>
extern "C" int refrun_frames(refrun *r, int is_double, uint32_t W, uint32_t H, const double image[4], uint32_t maxIter,
                             float maxSS, uint32_t flags, const uint32_t *palette, uint32_t palette_len, int warmup,
                             int steps, int to_host, double *wall_ms, float *main_ms_sum, float *compose_ms_sum,
                             uint32_t *rgba_out)
{
    ctx_scope s(r->ctx);
    unsigned char vis = 0;
    if (refrun_write_constant(r, "VISUALIZE_SAMPLE_COUNT", &vis, 1) != 0) return -1;
    CUdeviceptr buf[2];
    size_t pitch[2];
    for (int i = 0; i < 2; ++i) {
        RR_TRY(cuMemAllocPitch(&buf[i], &pitch[i], (size_t)W * 16, H, 16));
        RR_TRY(cuMemsetD8(buf[i], 0, pitch[i] * H));
    }
    CUarray a_out, a_pal;
    CUDA_ARRAY3D_DESCRIPTOR d;
    memset(&d, 0, sizeof d);
    d.Width = W; d.Height = H; d.Format = CU_AD_FORMAT_UNSIGNED_INT8; d.NumChannels = 4; d.Flags = CUDA_ARRAY3D_SURFACE_LDST;
    RR_TRY(cuArray3DCreate(&a_out, &d));
    d.Width = palette_len; d.Height = 1;
    RR_TRY(cuArray3DCreate(&a_pal, &d));
    CUDA_MEMCPY2D c;
    memset(&c, 0, sizeof c);
    c.srcMemoryType = CU_MEMORYTYPE_HOST; c.srcHost = palette; c.srcPitch = (size_t)palette_len * 4;
    c.dstMemoryType = CU_MEMORYTYPE_ARRAY; c.dstArray = a_pal;
    c.WidthInBytes = (size_t)palette_len * 4; c.Height = 1;
    RR_TRY(cuMemcpy2D(&c));
    uint32_t *pinned = nullptr;
    RR_TRY(cuMemHostAlloc((void **)&pinned, (size_t)W * H * 4, 0));
    CUevent e[4];
    for (int i = 0; i < 4; ++i) RR_TRY(cuEventCreate(&e[i], CU_EVENT_DEFAULT));

    long long pitch_ll = (long long)pitch[0], pitch2_ll = (long long)pitch[1];
    uint32_t size[2] = {W, H};
    float imf[4] = {(float)image[0], (float)image[1], (float)image[2], (float)image[3]};
    void *pm[7] = {&buf[0], &pitch_ll, size, is_double ? (void *)image : (void *)imf, &maxIter, &maxSS, &flags};
    unsigned gx = (W + 31) / 32, gy = (H + 31) / 32;
    float msum = 0, csum = 0;
    std::chrono::steady_clock::time_point t0;
    for (int it = 0; it < warmup + steps; ++it) {
        if (it == warmup) { RR_TRY(cuCtxSynchronize()); t0 = std::chrono::steady_clock::now(); msum = csum = 0; }
        RR_TRY(cuEventRecord(e[0], 0));
        RR_TRY(cuLaunchKernel(is_double ? r->main_d : r->main_f, gx, gy, 1, 32, 32, 1, 0, 0, pm, nullptr));
        RR_TRY(cuEventRecord(e[1], 0));
        RR_TRY(cuCtxSynchronize());
        RR_TRY(cuCtxSynchronize());
        CUsurfObject s_out, s_pal;
        CUDA_RESOURCE_DESC rd;
        memset(&rd, 0, sizeof rd);
        rd.resType = CU_RESOURCE_TYPE_ARRAY;
        rd.res.array.hArray = a_out;
        RR_TRY(cuSurfObjectCreate(&s_out, &rd));
        rd.res.array.hArray = a_pal;
        RR_TRY(cuSurfObjectCreate(&s_pal, &rd));
        void *pc[10] = {&buf[0], &pitch_ll, &buf[1], &pitch2_ll, &s_out, &W, &H, &s_pal, &palette_len, &maxSS};
        RR_TRY(cuEventRecord(e[2], 0));
        RR_TRY(cuLaunchKernel(r->compose, gx, gy, 1, 32, 32, 1, 0, 0, pc, nullptr));
        RR_TRY(cuEventRecord(e[3], 0));
        RR_TRY(cuCtxSynchronize());
        cuSurfObjectDestroy(s_out);
        cuSurfObjectDestroy(s_pal);
        if (to_host) {
            memset(&c, 0, sizeof c);
            c.srcMemoryType = CU_MEMORYTYPE_ARRAY; c.srcArray = a_out;
            c.dstMemoryType = CU_MEMORYTYPE_HOST; c.dstHost = pinned; c.dstPitch = (size_t)W * 4;
            c.WidthInBytes = (size_t)W * 4; c.Height = H;
            RR_TRY(cuMemcpy2D(&c));
        }
        float t = 0;
        cuEventElapsedTime(&t, e[0], e[1]); msum += t;
        cuEventElapsedTime(&t, e[2], e[3]); csum += t;
    }
    RR_TRY(cuCtxSynchronize());
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (wall_ms) *wall_ms = ms;
    if (main_ms_sum) *main_ms_sum = msum;
    if (compose_ms_sum) *compose_ms_sum = csum;
    if (rgba_out && to_host) memcpy(rgba_out, pinned, (size_t)W * H * 4);
    for (int i = 0; i < 4; ++i) cuEventDestroy(e[i]);
    cuMemFreeHost(pinned);
    cuArrayDestroy(a_out);
    cuArrayDestroy(a_pal);
    cuMemFree(buf[0]);
    cuMemFree(buf[1]);
    return 0;
}

/*
 * Zoom sequence of the reference host (bench.py --impl reference --workload c3): frame 0 is a quality frame, then
 * `steps` fast frames, each = CudaFractalRenderer.renderFast :159-184: advanced kernel (input = primary buffer,
 * output = secondary, imageReused = previous frame's segment), cuCtxSynchronize twice, buffer swap, surfaces
 * created, compose, cuCtxSynchronize, surfaces destroyed.  segments: (1 + warmup + steps) x 4 doubles, computed by
 * the caller with RenderingController.zoomAt.  is_double_per_frame: the precision the reference's own rule picks.
 */
extern "C" int refrun_zoom(refrun *r, uint32_t W, uint32_t H, const double *segments, const int *is_double_per_frame,
                           uint32_t maxIter, float maxSS, uint32_t flags, uint32_t focus_x, uint32_t focus_y,
                           const uint32_t *palette, uint32_t palette_len, int warmup, int steps, int to_host,
                           double *wall_ms, float *adv_ms_sum, float *compose_ms_sum, uint32_t *rgba_out)
{
    ctx_scope s(r->ctx);
    unsigned char vis = 0;
    if (refrun_write_constant(r, "VISUALIZE_SAMPLE_COUNT", &vis, 1) != 0) return -1;
    CUdeviceptr buf[2];
    size_t pitch[2];
    for (int i = 0; i < 2; ++i) {
        RR_TRY(cuMemAllocPitch(&buf[i], &pitch[i], (size_t)W * 16, H, 16));
        RR_TRY(cuMemsetD8(buf[i], 0, pitch[i] * H));
    }
    CUarray a_out, a_pal;
    CUDA_ARRAY3D_DESCRIPTOR d;
    memset(&d, 0, sizeof d);
    d.Width = W; d.Height = H; d.Format = CU_AD_FORMAT_UNSIGNED_INT8; d.NumChannels = 4; d.Flags = CUDA_ARRAY3D_SURFACE_LDST;
    RR_TRY(cuArray3DCreate(&a_out, &d));
    d.Width = palette_len; d.Height = 1;
    RR_TRY(cuArray3DCreate(&a_pal, &d));
    CUDA_MEMCPY2D c;
    memset(&c, 0, sizeof c);
    c.srcMemoryType = CU_MEMORYTYPE_HOST; c.srcHost = palette; c.srcPitch = (size_t)palette_len * 4;
    c.dstMemoryType = CU_MEMORYTYPE_ARRAY; c.dstArray = a_pal;
    c.WidthInBytes = (size_t)palette_len * 4; c.Height = 1;
    RR_TRY(cuMemcpy2D(&c));
    uint32_t *pinned = nullptr;
    RR_TRY(cuMemHostAlloc((void **)&pinned, (size_t)W * H * 4, 0));
    CUevent e[4];
    for (int i = 0; i < 4; ++i) RR_TRY(cuEventCreate(&e[i], CU_EVENT_DEFAULT));
    unsigned gx = (W + 31) / 32, gy = (H + 31) / 32;
    uint32_t size[2] = {W, H}, focus[2] = {focus_x, focus_y};
    float asum = 0, csum = 0;
    std::chrono::steady_clock::time_point t0;
    int prim = 0;
    for (int f = 0; f < 1 + warmup + steps; ++f) {
        if (f == 1 + warmup) { RR_TRY(cuCtxSynchronize()); t0 = std::chrono::steady_clock::now(); asum = csum = 0; }
        const double *img = segments + 4 * f;
        int dbl = is_double_per_frame[f];
        float imf[4] = {(float)img[0], (float)img[1], (float)img[2], (float)img[3]};
        long long p_out, p_in;
        RR_TRY(cuEventRecord(e[0], 0));
        if (f == 0) {
            prim = 0;
            p_out = (long long)pitch[0];
            float ss0 = maxSS < 1.f ? 1.f : maxSS;
            void *pm[7] = {&buf[0], &p_out, size, dbl ? (void *)img : (void *)imf, &maxIter, &ss0, &flags};
            RR_TRY(cuLaunchKernel(dbl ? r->main_d : r->main_f, gx, gy, 1, 32, 32, 1, 0, 0, pm, nullptr));
        } else {
            const double *old = segments + 4 * (f - 1);
            float oldf[4] = {(float)old[0], (float)old[1], (float)old[2], (float)old[3]};
            int sec = 1 - prim;
            p_out = (long long)pitch[sec]; p_in = (long long)pitch[prim];
            void *pa[11] = {&buf[sec], &p_out, size, dbl ? (void *)img : (void *)imf, &maxIter, &maxSS, &flags,
                            dbl ? (void *)old : (void *)oldf, &buf[prim], &p_in, focus};
            RR_TRY(cuLaunchKernel(dbl ? r->adv_d : r->adv_f, gx, gy, 1, 32, 32, 1, 0, 0, pa, nullptr));
            prim = sec;                                       /* switch2DBuffers */
        }
        RR_TRY(cuEventRecord(e[1], 0));
        RR_TRY(cuCtxSynchronize());
        RR_TRY(cuCtxSynchronize());
        CUsurfObject s_out, s_pal;
        CUDA_RESOURCE_DESC rd;
        memset(&rd, 0, sizeof rd);
        rd.resType = CU_RESOURCE_TYPE_ARRAY;
        rd.res.array.hArray = a_out;
        RR_TRY(cuSurfObjectCreate(&s_out, &rd));
        rd.res.array.hArray = a_pal;
        RR_TRY(cuSurfObjectCreate(&s_pal, &rd));
        long long pp = (long long)pitch[prim], ps = (long long)pitch[1 - prim];
        void *pc[10] = {&buf[prim], &pp, &buf[1 - prim], &ps, &s_out, &W, &H, &s_pal, &palette_len, &maxSS};
        RR_TRY(cuEventRecord(e[2], 0));
        RR_TRY(cuLaunchKernel(r->compose, gx, gy, 1, 32, 32, 1, 0, 0, pc, nullptr));
        RR_TRY(cuEventRecord(e[3], 0));
        RR_TRY(cuCtxSynchronize());
        cuSurfObjectDestroy(s_out);
        cuSurfObjectDestroy(s_pal);
        if (to_host) {
            memset(&c, 0, sizeof c);
            c.srcMemoryType = CU_MEMORYTYPE_ARRAY; c.srcArray = a_out;
            c.dstMemoryType = CU_MEMORYTYPE_HOST; c.dstHost = pinned; c.dstPitch = (size_t)W * 4;
            c.WidthInBytes = (size_t)W * 4; c.Height = H;
            RR_TRY(cuMemcpy2D(&c));
        }
        float t = 0;
        cuEventElapsedTime(&t, e[0], e[1]); asum += t;
        cuEventElapsedTime(&t, e[2], e[3]); csum += t;
    }
    RR_TRY(cuCtxSynchronize());
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (wall_ms) *wall_ms = ms;
    if (adv_ms_sum) *adv_ms_sum = asum;
    if (compose_ms_sum) *compose_ms_sum = csum;
    if (rgba_out && to_host) memcpy(rgba_out, pinned, (size_t)W * H * 4);
    for (int i = 0; i < 4; ++i) cuEventDestroy(e[i]);
    cuMemFreeHost(pinned);
    cuArrayDestroy(a_out);
    cuArrayDestroy(a_pal);
    cuMemFree(buf[0]);
    cuMemFree(buf[1]);
    return 0;
}
