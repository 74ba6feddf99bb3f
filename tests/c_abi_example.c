/* A host written directly against include/chaos_ultra.h in C: the same calls a JNI shim makes.  Renders one quality
 * frame and one fast frame of the mandelbrot module and prints an XOR checksum of each RGBA frame; tests/test_c_host.py
 * compares them with the frames the Python mirror gets.  usage: c_abi_example <libchaos_ultra.so> <kernels_dir> */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "chaos_ultra.h"

#define LOAD(name) __typeof__(&name) p_##name = (__typeof__(&name))dlsym(lib, #name); if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 2; }
#define CHECK(call) do { chaos_status st_ = (call); if (st_ != CHAOS_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, (int)st_, p_chaos_last_error()); return 1; } } while (0)

static uint32_t checksum(const uint32_t *p, size_t n) { uint32_t x = 0; for (size_t i = 0; i < n; ++i) x ^= p[i] * (uint32_t)(2654435761u + i); return x; }

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    void *lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    LOAD(chaos_provider_create) LOAD(chaos_provider_destroy) LOAD(chaos_open) LOAD(chaos_initialize) LOAD(chaos_render_quality)
    LOAD(chaos_render_fast) LOAD(chaos_output_rgba) LOAD(chaos_supply_defaults) LOAD(chaos_get_stats) LOAD(chaos_last_error) LOAD(chaos_close)
    const uint32_t W = 320, H = 180;
    uint32_t palette[512];
    for (uint32_t i = 0; i < 512; ++i) palette[i] = 0xff000000u | (i & 0xff) | ((i * 3 & 0xff) << 8) | ((255 - (i >> 1)) << 16);
    chaos_provider *prov = NULL;
    chaos_renderer *r = NULL;
    CHECK(p_chaos_provider_create(argv[2], 0, &prov));
    CHECK(p_chaos_open(prov, "mandelbrot", 0, &r));
    CHECK(p_chaos_initialize(r, W, H, palette, 512, CHAOS_OUTPUT_HOST));
    chaos_defaults d;
    memset(&d, 0, sizeof d);
    d.struct_size = sizeof d;
    CHECK(p_chaos_supply_defaults(r, &d));
    chaos_params m;
    memset(&m, 0, sizeof m);
    m.struct_size = sizeof m;
    m.max_iterations = d.max_iterations;                        /* 1600 */
    m.max_super_sampling = d.max_super_sampling;                /* 5 */
    double relw = 1.0 / H * W;                                  /* Model.setPlaneSegmentFromCenter */
    m.segment[0] = d.center_x - relw * d.zoom / 2; m.segment[1] = d.center_y - d.zoom / 2;
    m.segment[2] = d.center_x + relw * d.zoom / 2; m.segment[3] = d.center_y + d.zoom / 2;
    m.use_adaptive_super_sampling = 1; m.use_foveated_rendering = 1; m.use_sample_reuse = 1;
    m.mouse_focus[0] = W / 2; m.mouse_focus[1] = H / 2;
    CHECK(p_chaos_render_quality(r, &m));
    chaos_stats s;
    s.struct_size = sizeof s;
    CHECK(p_chaos_get_stats(r, &s));
    printf("quality %08x %llu %d\n", checksum(p_chaos_output_rgba(r), (size_t)W * H), (unsigned long long)s.pixel_iterations, m.float_precision);
    /* one zoom step about the centre, RenderingController.zoomAt with ZOOM_COEFF = 0.977f */
    double zc = (double)0.977f, sw = m.segment[2] - m.segment[0], sh = m.segment[3] - m.segment[1];
    double cx = m.segment[0] + sw * 0.5, cy = m.segment[1] + sh * 0.5;
    m.segment[0] = cx - sw * 0.5 * zc; m.segment[1] = cy - sh * 0.5 * zc; m.segment[2] = cx + sw * 0.5 * zc; m.segment[3] = cy + sh * 0.5 * zc;
    m.is_zooming = 1; m.is_zooming_in = 1;
    CHECK(p_chaos_render_fast(r, &m));
    CHECK(p_chaos_get_stats(r, &s));
    printf("fast %08x %llu %d\n", checksum(p_chaos_output_rgba(r), (size_t)W * H), (unsigned long long)s.pixel_iterations, m.float_precision);
    CHECK(p_chaos_close(r));
    CHECK(p_chaos_provider_destroy(prov));
    return 0;
}
