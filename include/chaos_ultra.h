/*
 * chaos_ultra.h -- C ABI of the B200-native render backend for chaos-ultra.
 *
 * This is the drop-in boundary: the functions below are exactly what a JNI / Panama binding
 * of the reference's renderer plugin interface would bind (INTEGRATION.md shows the Java
 * side).  Plain pointers and sizes only; no CUDA, torch or C++ types.
 *
 * Reference interfaces replaced (paths relative to
 * /root/reference/src/main/java/cz/cuni/mff/cgg/teichmaa/chaosultra/):
 *   rendering/FractalRendererProvider.java:11-27      -> chaos_provider_*, chaos_open
 *   rendering/FractalRenderer.java:14-77              -> chaos_initialize ... chaos_close
 *   cudarenderer/CudaFractalRendererProvider.java:14-91  (registry, one active renderer)
 *   cudarenderer/CudaFractalRenderer.java:32-430      (frame logic, state machine)
 *   cudarenderer/FractalRenderingModule.java:31-280   (module file by name, constants by name)
 *   rendering/model/RenderingModel.java:6-20          -> chaos_params
 *
 * Threading: like the reference (one AWT/GL/CUDA thread, CudaHelpers.java:47-49) all calls
 * on one provider must come from one thread at a time.  No GL context is required.
 * Errors: every function returns a chaos_status; chaos_last_error() returns the message of
 * the last failing call on the calling thread.  The library never aborts the process.
 */
#ifndef CHAOS_ULTRA_H
#define CHAOS_ULTRA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHAOS_ABI_VERSION 1
#if defined(__GNUC__)
#define CHAOS_API __attribute__((visibility("default")))
#else
#define CHAOS_API
#endif
#define CHAOS_MAX_SUPER_SAMPLING 64 /* FractalRenderer.java:15 */

typedef enum chaos_status {
    CHAOS_OK = 0,
    CHAOS_ERR_ILLEGAL_STATE = 1,    /* java.lang.IllegalStateException in the reference */
    CHAOS_ERR_ILLEGAL_ARGUMENT = 2, /* java.lang.IllegalArgumentException */
    CHAOS_ERR_CUDA_INIT = 3,        /* CudaInitializationException */
    CHAOS_ERR_CUDA = 4,             /* jcuda.CudaException at launch/copy time */
    CHAOS_ERR_RENDERER = 5          /* FractalRendererException (e.g. grid too large) */
} chaos_status;

/* util/FloatPrecision.java:3-10 */
typedef enum chaos_precision {
    CHAOS_PRECISION_SINGLE = 0,
    CHAOS_PRECISION_DOUBLE = 1,
    CHAOS_PRECISION_TOO_BIG = 2
} chaos_precision;

/* rendering/FractalRendererState.java */
typedef enum chaos_state { CHAOS_STATE_NOT_INITIALIZED = 0, CHAOS_STATE_READY_TO_RENDER = 1 } chaos_state;

/* where the composed RGBA8 frame is written */
typedef enum chaos_output_mode {
    CHAOS_OUTPUT_HOST = 0,  /* compose stores straight into library-owned pinned, mapped host memory */
    CHAOS_OUTPUT_DEVICE = 1 /* compose stores to device memory; chaos_download_rgba() copies on demand */
} chaos_output_mode;

/* The getters of RenderingModel (the interfaces under rendering/model/) that the renderer reads, plus the two
 * values it writes back (sampleReuseCacheDirty, floatingPointPrecision). */
typedef struct chaos_params {
    uint32_t struct_size;       /* = sizeof(chaos_params) */
    int32_t max_iterations;     /* IterationLimitModel; must be >= 1 (RenderingKernel.java:69) */
    double segment[4];          /* PlaneSegment: leftBottom.x, leftBottom.y, rightTop.x, rightTop.y; finite */
    float max_super_sampling;   /* SuperSamplingModel, [0, 64] (Model.java:166-168) */
    uint8_t use_adaptive_super_sampling;
    uint8_t visualise_sample_count;
    uint8_t use_foveated_rendering;
    uint8_t use_sample_reuse;
    uint8_t is_zooming;
    uint8_t is_zooming_in;
    uint8_t sample_reuse_cache_dirty; /* in/out: cleared by a quality render (CudaFractalRenderer.java:205) */
    uint8_t force_precision;    /* 0 = reference rule (CudaFractalRenderer.java:409-419); 1 = single; 2 = double */
    int32_t mouse_focus[2];     /* FoveatedRenderingModel.getMouseFocus */
    int32_t float_precision;    /* out: chaos_precision chosen for this frame */
} chaos_params;

/* DefaultFractalModel values a module supplies (supplyDefaultValues of the classes under cudarenderer/modules/) */
typedef struct chaos_defaults {
    uint32_t struct_size;
    uint8_t has_segment;        /* setPlaneSegmentFromCenter was called */
    uint8_t has_max_iterations;
    uint8_t has_max_super_sampling;
    uint8_t reserved;
    double center_x, center_y, zoom;
    int32_t max_iterations;
    float max_super_sampling;
    char custom_params[256];    /* setFractalCustomParams text ("" if none) */
} chaos_defaults;

/* Device-side timings and exact work counters of the most recent render call. */
typedef struct chaos_stats {
    uint32_t struct_size;
    uint32_t kernel_launches;        /* kernels of this library launched by the last render call */
    float render_ms;                 /* iteration / reuse kernel, CUDA events on the render stream */
    float compose_ms;                /* compose kernel */
    float reuse_ms;                  /* fast frames: the reprojection pass alone (part of render_ms), else 0 */
    float frame_ms;                  /* first render kernel's start to compose's end (what one frame costs on the device) */
    uint64_t pixel_iterations;       /* sum of escape-loop trip counts over every evaluated sample, as the reference's
                                      * loop counts them */
    uint64_t samples;                /* number of evaluated samples (orbits) */
    uint64_t launches_total;         /* kernels launched since the renderer was opened */
    uint64_t foreign_orbits;         /* several GPUs, one-sample frames: orbits of this rank's tiles that other ranks iterated
                                      * (cross-GPU tile stealing); their work is in THOSE ranks' pixel_iterations */
    uint64_t skipped_iterations;     /* the part of pixel_iterations that was PROVEN instead of executed: an orbit whose
                                      * state recurs bit for bit never escapes, so its trip count is maxIterations
                                      * (same records as the reference; CHAOS_SHORTCUTS=0 executes every trip) */
} chaos_stats;

typedef struct chaos_provider chaos_provider;
typedef struct chaos_renderer chaos_renderer;

/* ---- provider: CudaFractalRendererProvider.java:14-91, FractalRenderingModule.java:38-56 ---- */
/* kernels_dir plays the role of <user.dir>/<-DcudaKernelsDir>; modules are <dir>/<file>.cubin.
 * device is the CUDA ordinal (the reference is hard-wired to 0, CudaHelpers.java:40). */
CHAOS_API chaos_status chaos_provider_create(const char *kernels_dir, int device, chaos_provider **out);
CHAOS_API chaos_status chaos_provider_destroy(chaos_provider *p);
/* getAvailableFractals(): the registered display names, independent of which files exist. */
CHAOS_API chaos_status chaos_list_fractals(chaos_provider *p, const char **names, uint32_t capacity, uint32_t *count);
/* A fractal author's own module: <kernels_dir>/<file_stem>.cubin becomes available under `fractal_name` -- what adding a
 * Module*.java and registering it does in the reference (CudaFractalRendererProvider.java:19-31).  The module is built from
 * a file written for the reference's contract by `python chaos-ultra_b200/build.py compat <file.cu>` (csrc/compat/) or from a
 * `struct Fractal` file; it has no defaults, and its constants are written with chaos_write_constant. */
CHAOS_API chaos_status chaos_register_module(chaos_provider *p, const char *fractal_name, const char *file_stem);
/* getRenderer(name, forceReload) :47-66 -- at most one active renderer per provider; same name and
 * !force_reload returns the active one; otherwise the old one is closed (module unloaded) and the
 * module file is read again. */
CHAOS_API chaos_status chaos_open(chaos_provider *p, const char *fractal_name, int force_reload, chaos_renderer **out);
/* The provider's active renderer, NULL if there is none.  An unknown name leaves the active renderer untouched; a
 * chaos_open that fails later (module file missing or corrupt) has already closed it, as the reference does
 * (CudaFractalRendererProvider.java:52) -- the old handle is then dead and this returns NULL. */
CHAOS_API chaos_renderer *chaos_active_renderer(const chaos_provider *p);

/* ---- renderer: FractalRenderer.java:14-77 / CudaFractalRenderer.java ---- */
/* initializeRendering(GLParams) :85-101.  palette_rgba: R in bits 0-7 (ImageHelpers.java:138-158); copied. */
CHAOS_API chaos_status chaos_initialize(chaos_renderer *r, uint32_t width, uint32_t height,
                              const uint32_t *palette_rgba, uint32_t palette_len, chaos_output_mode mode);
CHAOS_API chaos_status chaos_free_resources(chaos_renderer *r);          /* freeRenderingResources :103-112 */
CHAOS_API chaos_status chaos_render_quality(chaos_renderer *r, chaos_params *model); /* renderQuality :187-206 */
CHAOS_API chaos_status chaos_render_fast(chaos_renderer *r, chaos_params *model);    /* renderFast :159-184 */
CHAOS_API chaos_status chaos_debug(chaos_renderer *r);                   /* launchDebugKernel :147-154 */
CHAOS_API chaos_status chaos_set_custom_params(chaos_renderer *r, const char *text); /* setFractalCustomParams :388-391 */
/* FractalRenderingModule.writeToConstantMemory :163-227: size-checked write to a named __constant__ */
CHAOS_API chaos_status chaos_write_constant(chaos_renderer *r, const char *symbol, const void *data, size_t bytes);
CHAOS_API chaos_status chaos_supply_defaults(chaos_renderer *r, chaos_defaults *out); /* supplyDefaultValues */
CHAOS_API chaos_status chaos_close(chaos_renderer *r);                   /* close :381-386 */

CHAOS_API chaos_state chaos_get_state(const chaos_renderer *r);
CHAOS_API uint32_t chaos_get_width(const chaos_renderer *r);
CHAOS_API uint32_t chaos_get_height(const chaos_renderer *r);
CHAOS_API const char *chaos_fractal_name(const chaos_renderer *r);

/* The composed frame: width*height RGBA8, row 0 = top of the plane segment (Model.java:23-32).
 * HOST mode: pointer into pinned memory, valid until chaos_free_resources.  DEVICE mode: NULL. */
CHAOS_API const uint32_t *chaos_output_rgba(const chaos_renderer *r);
/* device address (as integer) of the RGBA frame in DEVICE mode, 0 otherwise; for NVLink gathers */
CHAOS_API uint64_t chaos_output_rgba_device(const chaos_renderer *r);
/* copy the composed frame (either mode) into caller memory */
CHAOS_API chaos_status chaos_download_rgba(chaos_renderer *r, uint32_t *dst, size_t dst_bytes);
/* copy the current primary pixel_info_t buffer (16 B/pixel, pitch removed) into caller memory;
 * the debugging counterpart of copy2DFromDevToHost (CudaFractalRenderer.java:119-134) */
CHAOS_API chaos_status chaos_download_records(chaos_renderer *r, void *dst, size_t dst_bytes);
CHAOS_API chaos_status chaos_get_stats(const chaos_renderer *r, chaos_stats *out);
/* Diagnostics (no counterpart in the reference): copy the renderer's device-side scheduler counters as they are right
 * now into dst -- callable from a second thread while a render call is running (tools/pool_watch.py looks at a frame
 * that does not end with it).  Returns the size of one counter block (there is one per strand), negative on error. */
CHAOS_API int chaos_debug_peek_counters(chaos_renderer *r, void *dst, size_t bytes);

/* Multi-GPU: this renderer renders and composes only the row bands b with b % part_count ==
 * part_index, bands being band_rows pixel rows high (a multiple of 4).  part_count 1 = whole frame.
 * While a partition is set chaos_render_fast renders a quality frame of the own bands: the previous frame's
 * records of the other ranks' bands are not in this renderer's memory, so there is nothing valid to reproject. */
CHAOS_API chaos_status chaos_set_partition(chaos_renderer *r, uint32_t part_index, uint32_t part_count, uint32_t band_rows);

/* Multi-GPU, DEVICE mode: compose writes its bands into `device_ptr` instead of the renderer's own frame -- e.g. rank 0's
 * frame mapped into this process with CUDA IPC, so that the composed bands cross NVLink as the compose kernel's own
 * stores and no separate gather step is needed.  width*height*4 bytes must be writable there.  0 = the own frame again.
 * chaos_free_resources resets it. */
CHAOS_API chaos_status chaos_set_output_target(chaos_renderer *r, uint64_t device_ptr);

/* ---- multi-GPU plumbing: one process per GPU (the reference has none of this: device 0 only, CudaHelpers.java:36-41) ----
 * The frame of an N-GPU render is assembled by the compose kernels themselves: every rank composes its bands straight
 * into ONE frame, either rank 0's device frame (chaos_ipc_export_frame / chaos_ipc_open_frame: the bands cross NVLink
 * as 128-bit stores) or a frame in host shared memory that every rank's process has mapped (chaos_set_host_target:
 * every GPU writes its bands over its own PCIe link).  chaos_set_frame_barrier makes every render call end when all
 * ranks' bands have landed. */
typedef struct chaos_ipc_handle { unsigned char bytes[64]; } chaos_ipc_handle;
/* DEVICE mode, the rank that owns the frame: a handle other processes can open */
CHAOS_API chaos_status chaos_ipc_export_frame(chaos_renderer *r, chaos_ipc_handle *out);
/* DEVICE mode, the other ranks: map that frame into this process and make it the compose target
 * (chaos_set_output_target with the mapped address); chaos_free_resources / chaos_close unmap it */
CHAOS_API chaos_status chaos_ipc_open_frame(chaos_renderer *r, const chaos_ipc_handle *frame);
/* DEVICE mode: compose writes into caller-owned HOST memory instead (width*height*4 bytes, page-aligned, e.g. a POSIX
 * shared-memory segment all ranks have mapped).  The library pins and maps it (cuMemHostRegister) for as long as it is
 * the target; NULL, chaos_free_resources and chaos_close release it. */
CHAOS_API chaos_status chaos_set_host_target(chaos_renderer *r, void *host_frame, size_t bytes);
/* Multi-GPU fast frames (zoom sequences).  The frame is cut into one slab of rows per rank: chaos_set_partition(rank, world,
 * band_rows) with band_rows * world >= height.  Every rank keeps the records of its slab and exports its two record buffers;
 * every rank opens every other rank's pair.  chaos_render_fast then reprojects for real: a tap into another slab is a peer
 * load from the owner's primary buffer over NVLink.  Needs a frame barrier (a rank must not start frame f + 1 before every
 * rank has finished frame f) and the same sequence of render calls on all ranks.  Without the peers' buffers a fast frame
 * of a partitioned renderer renders its bands afresh (see chaos_set_partition).
 * The same mappings give one-sample quality frames (the deep-zoom configuration) DYNAMIC REDISTRIBUTION: a rank whose own tiles
 * are handed out claims tiles from the other ranks' cursors (system-scope atomics over NVLink), iterates them and stores the
 * records into the owner's buffer; the owner's call returns when every pixel of its tiles has been finished by somebody. */
CHAOS_API chaos_status chaos_ipc_export_records(chaos_renderer *r, chaos_ipc_handle out[3]);   /* two record buffers + scheduler counters */
CHAOS_API chaos_status chaos_ipc_open_records(chaos_renderer *r, uint32_t peer_rank, const chaos_ipc_handle in[3]);
/* `shm_block`: 64 zero-initialised bytes of host memory shared by the `world` processes of the job (NULL = no barrier).
 * Every later render call announces its frame there when its own kernels are done and returns when all ranks have
 * announced theirs -- on return the target frame holds every rank's bands.  Waits are bounded (10 s -> CHAOS_ERR_CUDA). */
CHAOS_API chaos_status chaos_set_frame_barrier(chaos_renderer *r, void *shm_block, uint32_t world);

/* ---- frame driver: rendering/RenderingModeFSM.java:9-155 + the automatic-quality controller of
 * rendering/GLRenderer.java:113-245 + RenderingController.zoomAt (:130-150), see csrc/chaos_driver.cpp ---- */
typedef enum chaos_rendering_mode {      /* RenderingModeFSM.RenderingMode */
    CHAOS_MODE_WAITING = 0, CHAOS_MODE_ZOOMING_AUTO = 1, CHAOS_MODE_ZOOMING_ONCE = 2, CHAOS_MODE_MOVING = 3, CHAOS_MODE_PROGRESSIVE_RENDERING = 4
} chaos_rendering_mode;
typedef enum chaos_driver_clock {
    CHAOS_CLOCK_DEVICE = 0,    /* the controller sees the frame's device time (CUDA events, float ms) */
    CHAOS_CLOCK_WALL_INT = 1   /* host clock truncated to int ms, as the reference (GLRenderer.java:129,145) */
} chaos_driver_clock;
typedef struct chaos_driver chaos_driver;
/* a frame rendered by somebody else (tests; a host that wraps the render calls): fill *frame_ms or leave it negative */
typedef chaos_status (*chaos_render_fn)(void *user, chaos_params *model, float *frame_ms);
typedef struct chaos_driver_state {
    uint32_t struct_size;
    chaos_rendering_mode mode, last_mode;
    int32_t progressive_rendering_level;
    uint8_t zooming, moving, zooming_in, last_kind;   /* last_kind: 0 nothing rendered yet, 1 fast, 2 quality */
    float last_frame_render_time_ms;
    uint64_t frames;
    chaos_params model;                                /* the model as the driver has left it (segment, maxSuperSampling, flags) */
} chaos_driver_state;
CHAOS_API chaos_status chaos_driver_create(chaos_renderer *r, const chaos_params *model, chaos_driver **out);
CHAOS_API chaos_status chaos_driver_create_custom(chaos_render_fn fast, chaos_render_fn quality, void *user, const chaos_params *model,
                                                  uint32_t canvas_width, uint32_t canvas_height, chaos_driver **out);
CHAOS_API chaos_status chaos_driver_destroy(chaos_driver *d);
CHAOS_API chaos_status chaos_driver_mouse(chaos_driver *d, int x, int y);                        /* lastMousePosition, mouseFocus */
CHAOS_API chaos_status chaos_driver_start_zooming(chaos_driver *d, int inside, int moving_too);  /* startZooming / startZoomingAndMoving */
CHAOS_API chaos_status chaos_driver_zoom_once(chaos_driver *d, int inside);                      /* doZoomingManualOnce */
CHAOS_API chaos_status chaos_driver_stop_zooming(chaos_driver *d);
CHAOS_API chaos_status chaos_driver_start_moving(chaos_driver *d);
CHAOS_API chaos_status chaos_driver_stop_moving(chaos_driver *d);
CHAOS_API chaos_status chaos_driver_start_progressive_rendering(chaos_driver *d, int reset_first);
CHAOS_API chaos_status chaos_driver_step(chaos_driver *d);                                       /* RenderingModeFSM.step() alone */
CHAOS_API chaos_status chaos_driver_set_automatic_quality(chaos_driver *d, int on);
CHAOS_API chaos_status chaos_driver_set_clock(chaos_driver *d, chaos_driver_clock clock);
CHAOS_API chaos_status chaos_driver_set_model(chaos_driver *d, const chaos_params *model);
CHAOS_API chaos_status chaos_driver_display(chaos_driver *d, int *rendered);                     /* GLRenderer.display(): one tick */
CHAOS_API chaos_status chaos_driver_get_state(const chaos_driver *d, chaos_driver_state *out);
CHAOS_API chaos_status chaos_driver_run_zoom_session(chaos_driver *d, int x, int y, int inside, uint32_t ticks, uint32_t *frames_rendered);
CHAOS_API const char *chaos_driver_last_error(void);

CHAOS_API const char *chaos_last_error(void);
CHAOS_API uint32_t chaos_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CHAOS_ULTRA_H */
