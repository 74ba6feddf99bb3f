/*
 * oracle/chaos_oracle.h -- interface of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 * See chaos_oracle.c for what is restated and from which reference lines.
 */
#ifndef CHAOS_ORACLE_H
#define CHAOS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* helpers.cuh:106-130 -- 16 bytes, align 4; bytes 9..11 are padding the kernels never write */
typedef struct {
    float value;
    float weight;
    uint8_t isReused;
    uint8_t pad[3];
    float weightOfNewSamples;
} ora_pixel_info;

enum { ORA_FRACTAL_MANDELBROT = 0, ORA_FRACTAL_JULIA = 1, ORA_FRACTAL_TEST = 2 };

typedef struct {
    int kind;
    double julia_c[2];   /* julia.cu:3 */
    int amplifier;       /* test.cu:6 */
} ora_fractal;

/* fractalRendererGeneric.cu:157-161 */
enum {
    ORA_FLAG_ADAPTIVE_SS = 1u << 0,
    ORA_FLAG_FOVEATION = 1u << 2,
    ORA_FLAG_SAMPLE_REUSE = 1u << 3,
    ORA_FLAG_IS_ZOOMING = 1u << 4,
    ORA_FLAG_ZOOMING_IN = 1u << 5
};

/* which build of the reference the pixel->plane mapping follows */
enum { ORA_VARIANT_NVCC129 = 0, ORA_VARIANT_SHIPPED_PTX = 1 };

typedef struct { uint32_t x, y; } ora_lane;

typedef struct {
    const ora_fractal *fractal;
    int real_is_double;          /* 0: *Float kernels, 1: *Double kernels */
    int variant;                 /* ORA_VARIANT_* */
    uint32_t width, height;
    double image[4];             /* lb.x, lb.y, rt.x, rt.y */
    double image_reused[4];      /* advanced only */
    uint32_t focus_x, focus_y;   /* advanced only */
    uint32_t maxIter;
    float maxSuperSampling;
    uint32_t flags;
    /* optional statistics (may be NULL except trips) */
    uint64_t *trips;             /* += while-loop trip count of every evaluated sample */
    uint64_t *sample_hist;       /* [64]: += samples evaluated in round i */
    uint64_t *ub_reads;          /* += reads of samples[k>=10] (undefined behaviour in the reference) */
} ora_frame;

int ora_render_main(const ora_frame *f, void *out, size_t out_pitch);
int ora_render_main_rows(const ora_frame *f, void *out, size_t out_pitch, uint32_t tr0, uint32_t tr1, uint32_t stride);
int ora_render_advanced(const ora_frame *f, void *out, size_t out_pitch, const void *in, size_t in_pitch);
int ora_compose(const ora_frame *f, const void *in, size_t in_pitch, uint32_t *out_rgba,
                const uint32_t *palette, uint32_t palette_len, int visualize_sample_count);

void ora_default_palette(uint32_t *p1536);
int ora_choose_precision(const double image[4], uint32_t W, uint32_t H);
void ora_segment_from_center(double cx, double cy, double zoom, uint32_t W, uint32_t H, double image[4]);
void ora_zoom_at(double image[4], uint32_t W, uint32_t H, int where_x, int where_y, int into);
uint64_t ora_scalar_rows(const ora_frame *f, uint32_t y0, uint32_t y1, uint32_t *et_out);

#ifdef __cplusplus
}
#endif
#endif
