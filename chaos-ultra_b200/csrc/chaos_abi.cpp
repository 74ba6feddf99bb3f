/*
 * chaos_abi.cpp -- host half of the render backend behind include/chaos_ultra.h.
 *
 * Re-implements, on the CUDA driver API and without any Java/JCuda/GL dependency, the host
 * logic of the reference backend (paths under
 * /root/reference/src/main/java/cz/cuni/mff/cgg/teichmaa/chaosultra/cudarenderer/):
 *   CudaFractalRendererProvider.java:14-91   registry of modules, one active renderer
 *   FractalRenderingModule.java:31-280       module = one file found by name; constants by name
 *   modules/Module*.java                     per-fractal defaults and custom-parameter parsing
 *   CudaFractalRenderer.java:32-430          state machine, quality/fast frame logic, precision rule
 *   DeviceMemoryDoubleBuffer2D.java:17-152   two pitched record buffers, swap, dirty flag
 *   RenderingKernel.java:68-72,124-146       argument validation, float/double limit tests
 * What is new: modules are sm_100a cubins launched as persistent grids sized from the SM count,
 * the composed frame goes to pinned host memory (or stays on the device) instead of a GL
 * texture, every frame is timed with CUDA events, and a renderer can own a row-band partition
 * of the frame for multi-GPU rendering.
 *
 * There is deliberately no CPU fallback: without a driver, a device or the module file every
 * entry point fails with a message.
 */
#include <dlfcn.h>
#include <sched.h>
#include <time.h>
#include <errno.h>
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/chaos_ultra.h"
#include "chaos_device.h"
#include "cuda_driver.h"
#include "mini_json.h"

/* ------------------------------------------------------------------------------------------
 * errors
 * ---------------------------------------------------------------------------------------- */
static thread_local char g_last_error[1024] = "";

static chaos_status fail(chaos_status st, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof g_last_error, fmt, ap);
    va_end(ap);
    return st;
}

extern "C" const char *chaos_last_error(void) { return g_last_error; }
extern "C" uint32_t chaos_abi_version(void) { return CHAOS_ABI_VERSION; }

/* ------------------------------------------------------------------------------------------
 * driver binding
 * ---------------------------------------------------------------------------------------- */
#define CHAOS_STR2(x) #x
#define CHAOS_STR(x) CHAOS_STR2(x)

const chaos_cuda_driver *chaos_cuda_driver_get(const char **err)
{
    static chaos_cuda_driver drv;
    static bool tried = false, ok = false;
    static std::string error;
    if (!tried) {
        tried = true;
        drv.handle = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!drv.handle) drv.handle = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
        if (!drv.handle) {
            error = std::string("Error while loading the Cuda native library. Do you have CUDA installed? (") + dlerror() + ")";
        } else {
            ok = true;
            /* cuda.h maps e.g. cuMemAlloc -> cuMemAlloc_v2; stringify after expansion */
#define X(name)                                                                    \
    drv.p_##name = (decltype(&name))dlsym(drv.handle, CHAOS_STR(name));            \
    if (!drv.p_##name) { ok = false; error = std::string("libcuda lacks ") + CHAOS_STR(name); }
            CHAOS_CU_FUNCS(X)
#undef X
        }
    }
    if (!ok) {
        if (err) *err = error.c_str();
        return nullptr;
    }
    return &drv;
}

static const chaos_cuda_driver *D = nullptr;

static const char *cu_err_name(CUresult r)
{
    const char *s = nullptr;
    if (D && D->p_cuGetErrorName(r, &s) == CUDA_SUCCESS && s) return s;
    return "CUDA_ERROR_UNKNOWN";
}

#define CU_TRY(call, st)                                                                     \
    do {                                                                                     \
        CUresult _r = D->p_##call;                                                           \
        if (_r != CUDA_SUCCESS) return fail(st, "%s failed: %s", #call, cu_err_name(_r));    \
    } while (0)

/* ------------------------------------------------------------------------------------------
 * module registry (CudaFractalRendererProvider.java:19-31 + modules/*.java)
 * ---------------------------------------------------------------------------------------- */
struct chaos_renderer;

struct module_desc {
    const char *fractal_name;                      /* display name, the key of getRenderer() */
    const char *file_stem;                         /* <kernels_dir>/<file_stem>.cubin */
    chaos_status (*on_initialize)(chaos_renderer *);                  /* Module*.initialize() extras */
    chaos_status (*set_custom_params)(chaos_renderer *, const char *);
    void (*supply_defaults)(chaos_defaults *);
};

static chaos_status write_constant(chaos_renderer *r, const char *symbol, const void *data, size_t bytes, const char *what);

/* FractalRenderingModule.parseParamsAsDoubles :233-245: split on [,;], Double.parseDouble (trims) */
static bool parse_doubles(const char *text, std::vector<double> &out)
{
    std::string s(text ? text : "");
    size_t pos = 0;
    while (pos <= s.size()) {
        size_t e = s.find_first_of(",;", pos);
        if (e == std::string::npos) e = s.size();
        std::string tok = s.substr(pos, e - pos);
        size_t a = tok.find_first_not_of(" \t\r\n"), b = tok.find_last_not_of(" \t\r\n");
        if (a == std::string::npos) return false;      /* NumberFormatException: empty String */
        tok = tok.substr(a, b - a + 1);
        char *end = nullptr;
        double v = strtod(tok.c_str(), &end);
        if (end == tok.c_str() || *end != '\0') return false;
        out.push_back(v);
        pos = e + 1;
        if (e == s.size()) break;
    }
    return !out.empty();
}

static void defaults_base(chaos_defaults *d) { d->custom_params[0] = '\0'; }

/* modules/ModuleMandelbrot.java:9-25 */
static void defaults_mandelbrot(chaos_defaults *d)
{
    defaults_base(d);
    d->has_segment = 1; d->center_x = -0.5; d->center_y = 0; d->zoom = 2;
    d->has_max_iterations = 1; d->max_iterations = 1600;
    d->has_max_super_sampling = 1; d->max_super_sampling = 5;
}
static chaos_status custom_none(chaos_renderer *, const char *) { return CHAOS_OK; }

/* modules/ModuleJulia.java:9-49 */
static chaos_status julia_set_c(chaos_renderer *r, double x, double y)
{
    double c[2] = {x, y};
    return write_constant(r, "julia_c", c, sizeof c, "PointDouble");
}
static chaos_status julia_on_initialize(chaos_renderer *r) { return julia_set_c(r, 0, 0); }
static chaos_status julia_custom(chaos_renderer *r, const char *text)
{
    std::vector<double> v;
    if (!parse_doubles(text, v)) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "NumberFormatException: For input string: \"%s\"", text ? text : "");
    if (v.size() < 2) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "ArrayIndexOutOfBoundsException: julia expects \"x;y\" but got \"%s\"", text);
    return julia_set_c(r, v[0], v[1]);
}
static void defaults_julia(chaos_defaults *d)
{
    defaults_base(d);
    d->has_max_iterations = 1; d->max_iterations = 900;
    snprintf(d->custom_params, sizeof d->custom_params, "-0.4;0.6");
}

/* modules/ModuleTest.java:9-23 */
static chaos_status test_custom(chaos_renderer *r, const char *text)
{
    /* Integer.parseInt: optional sign, then digits only (no surrounding white space), value within int */
    const char *t = text ? text : "";
    char *end = nullptr;
    errno = 0;
    long long v = strtoll(t, &end, 10);
    const bool shape_ok = (*t == '-' || *t == '+') ? (t[1] >= '0' && t[1] <= '9') : (*t >= '0' && *t <= '9');
    if (!shape_ok || *end != '\0' || errno == ERANGE || v < INT_MIN || v > INT_MAX)
        return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "NumberFormatException: For input string: \"%s\"", t);
    int iv = (int)v;
    return write_constant(r, "amplifier", &iv, sizeof iv, "double");
}
static void defaults_test(chaos_defaults *d)
{
    defaults_base(d);
    snprintf(d->custom_params, sizeof d->custom_params, "10");
}
/* modules/ModuleNewtonWired.java:7-20 */
static void defaults_newton_wired(chaos_defaults *d)
{
    defaults_base(d);
    d->has_max_iterations = 1; d->max_iterations = 200;
}

/* modules/ModuleNewtonGeneric.java:34-74: {"coefficients":[a3,a2,a1,a0], "roots":[[re,im] x 3]}; the coefficient
 * order is reversed between the user's text and the device array (:46-49) */
static const char *k_newton_default_params =
    "{ \"coefficients\" : [1, 0, 0, -1], \"roots\" : [ [1,0], [-0.5,0.86602540378] , [-0.5,-0.86602540378] ] }";
static chaos_status newton_parse(const char *text, mj_value &json)
{
    mj_parser ps(text);
    if (!ps.parse(json) || json.kind != mj_value::OBJ)
        return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "JsonSyntaxException: cannot parse the fractal parameters: %s", text ? text : "");
    return CHAOS_OK;
}
static chaos_status newton_generic_custom(chaos_renderer *r, const char *text)
{
    mj_value json;
    chaos_status st = newton_parse(text, json);
    if (st != CHAOS_OK) return st;
    const mj_value *co = json.get("coefficients"), *ro = json.get("roots");
    if (!co || co->kind != mj_value::ARR || !ro || ro->kind != mj_value::ARR)
        return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "NullPointerException: the parameters need \"coefficients\" and \"roots\" arrays");
    for (const mj_value &root : ro->arr)
        if (root.kind != mj_value::ARR || root.arr.size() != 2)
            return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "Found a root that is not represented as [real, imag].");
    /* Gson's getAsDouble throws on anything that is not a number (strings, null, booleans, nested arrays) */
    for (const mj_value &root : ro->arr)
        for (const mj_value &part : root.arr)
            if (part.kind != mj_value::NUM) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "NumberFormatException: a root component is not a number");
    for (const mj_value &c : co->arr)
        if (c.kind != mj_value::NUM) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "NumberFormatException: a coefficient is not a number");
    if (co->arr.size() != 4) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "expecting 4 coefficients");
    if (ro->arr.size() != 3) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "expecting 3 roots");
    double roots[6], coefs[4];
    for (int i = 0; i < 3; ++i) { roots[2 * i] = ro->arr[i].arr[0].num; roots[2 * i + 1] = ro->arr[i].arr[1].num; }
    for (int i = 0; i < 4; ++i) coefs[3 - i] = co->arr[i].num;
    st = write_constant(r, "roots", roots, sizeof roots, "double[] of size 6");
    if (st != CHAOS_OK) return st;
    return write_constant(r, "coefficients", coefs, sizeof coefs, "double[] of size 4");
}
static void defaults_newton_generic(chaos_defaults *d)
{
    defaults_base(d);
    d->has_max_iterations = 1; d->max_iterations = 200;
    snprintf(d->custom_params, sizeof d->custom_params, "%s", k_newton_default_params);
}

/* modules/ModuleNewtonIterations.java:9-33: the generic parameters plus an optional "colorMagnifier" */
static chaos_status newton_iterations_custom(chaos_renderer *r, const char *text)
{
    chaos_status st = newton_generic_custom(r, text);
    if (st != CHAOS_OK) return st;
    mj_value json;
    st = newton_parse(text, json);
    if (st != CHAOS_OK) return st;
    if (const mj_value *cm = json.get("colorMagnifier")) {
        if (cm->kind != mj_value::NUM || !(cm->num >= (double)INT_MIN && cm->num <= (double)INT_MAX))
            return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "NumberFormatException: colorMagnifier is not an int");
        int v = (int)cm->num;
        return write_constant(r, "colorMagnifier", &v, sizeof v, "double");
    }
    return CHAOS_OK;
}
static void defaults_newton_iterations(chaos_defaults *d)
{
    defaults_newton_generic(d);
    /* "{\"colorMagnifier\": 11," + the generic text without its opening brace (:28-31) */
    snprintf(d->custom_params, sizeof d->custom_params, "{\"colorMagnifier\": 11,%s", k_newton_default_params + 1);
}

/* modules/ModuleGoci.java:9-28 */
static void defaults_goc(chaos_defaults *d)
{
    defaults_base(d);
    d->has_max_iterations = 1; d->max_iterations = 900;
    d->has_segment = 1; d->center_x = 1.1; d->center_y = -0.2; d->zoom = 0.20000000000000004;
}

/* the 7 names the reference registers (CudaFractalRendererProvider.java:21-27), display name -> module file */
static const module_desc g_modules[] = {
    {"julia", "julia", julia_on_initialize, julia_custom, defaults_julia},
    {"mandelbrot", "mandelbrot", nullptr, custom_none, defaults_mandelbrot},
    {"newton wired", "newton_wired", nullptr, custom_none, defaults_newton_wired},
    {"newton generic", "newton_generic", nullptr, newton_generic_custom, defaults_newton_generic},
    {"newton colored by iterations", "newton_iterations", nullptr, newton_iterations_custom, defaults_newton_iterations},
    {"test", "test", nullptr, test_custom, defaults_test},
    {"goc", "goc", nullptr, custom_none, defaults_goc},
};
static const uint32_t g_n_modules = sizeof g_modules / sizeof g_modules[0];

/* ------------------------------------------------------------------------------------------
 * objects
 * ---------------------------------------------------------------------------------------- */
struct registered_module {      /* chaos_register_module: owns the strings its module_desc points at */
    std::string name, stem;
    module_desc desc;
};

struct chaos_provider {
    std::vector<registered_module *> registered;
    std::string kernels_dir;
    int device = 0;
    CUdevice cu_device = 0;
    CUcontext ctx = nullptr;
    int sm_count = 0;
    chaos_renderer *active = nullptr;
};

struct ctx_guard {
    bool pushed = false;
    explicit ctx_guard(chaos_provider *p) { if (p && p->ctx && D->p_cuCtxPushCurrent(p->ctx) == CUDA_SUCCESS) pushed = true; }
    ~ctx_guard() { if (pushed) { CUcontext c; D->p_cuCtxPopCurrent(&c); } }
};

#define CHAOS_MAX_STRANDS 8
#define CHAOS_DEFAULT_STRANDS 2

struct record_buffer {
    CUdeviceptr ptr = 0;
    size_t pitch = 0;
};

struct chaos_renderer {
    chaos_provider *provider = nullptr;
    const module_desc *desc = nullptr;
    /* module (FractalRenderingModule) */
    CUmodule module = nullptr;
    CUfunction k_main_f = nullptr, k_main_d = nullptr, k_adv_f = nullptr, k_adv_d = nullptr;
    CUfunction k_pass_a[2] = {nullptr, nullptr}, k_pass_b[2] = {nullptr, nullptr}, k_pass_c[2] = {nullptr, nullptr};   /* [0] float, [1] double */
    int blocks_pass_a[2] = {0, 0}, blocks_pass_b[2] = {0, 0}, blocks_pass_c[2] = {0, 0};
    CUfunction k_main_f_sync = nullptr, k_main_d_sync = nullptr;   /* engine 0 (differential check) */
    /* engine 2 (render_streams.cuh): probe -> long -> finish, [0] float, [1] double */
    CUfunction k_probe[2] = {nullptr, nullptr}, k_long[2] = {nullptr, nullptr}, k_finish[2] = {nullptr, nullptr};
    int blocks_probe[2] = {0, 0}, blocks_long[2] = {0, 0}, blocks_finish[2] = {0, 0};
    CUdeviceptr long_list = 0, finish_list = 0;    /* allocated by the first engine-2 frame of a frame size */
    size_t list_capacity = 0;                      /* entries */
    uint32_t probe_trips = 64;
    uint32_t list_shrink = 1;                      /* CHAOS_LIST_SHRINK=n: the lists take 1/n of their entries (tests: the overflow paths) */
    uint32_t long_occ[3] = {2, 4, 8};             /* chaos_render_args::occ_orbits_per_lane (CHAOS_LONG_OCC=a,b,c; 0,0,0 = always the full grid) */
    uint32_t hot_first = 3;                        /* orbits expected to be long are started first: bit 0 pass C (by sample 0's cost), bit 1 the
                                                    * other passes (survivors of tiles next to the boundary); CHAOS_HOT_FIRST=0: list order */
    CUfunction k_classify = nullptr, k_order = nullptr;            /* between the two passes of engine 1 */
    CUfunction k_replay = nullptr;                                 /* pass D */
    CUfunction k_export_all = nullptr;                             /* between order and pass B: few tiles left -> all exported */
    chaos_export exp_buf = {0, nullptr, nullptr, nullptr, nullptr, nullptr};
    CUdeviceptr late_tiles = 0;    /* one bit per vote tile of the frame, see chaos_render_args::late_tiles */   /* pass B -> pass C -> pass D (device memory) */
    uint32_t export_enabled = 1;
    int export_all_below = -1;     /* see chaos_render_args::export_all_below; -1 = two tiles per slot of the pass B launch */
    CUfunction k_reuse_f = nullptr, k_reuse_d = nullptr;           /* pass R of a fast frame */
    int blocks_reuse_f = 0, blocks_reuse_d = 0;
    CUdeviceptr tile_key = 0, tile_order = 0;
    CUdeviceptr warp_trace = 0;         /* diagnostics, CHAOS_WARP_TRACE=<file>: per-warp timeline of pass B */
    /* diagnostics, CHAOS_TIMELINE=<file>: an event pair around every launch of a frame, written as "kernel stream start_ms end_ms"
     * (start = when the stream's previous work was through); changes nothing but the frame's host-side cost */
    struct timeline_entry { const char *name; int stream; CUevent a, b; };
    std::vector<timeline_entry> timeline;
    size_t timeline_used = 0;
    const char *timeline_path = nullptr;
    std::vector<std::pair<CUfunction, const char *>> fn_names;
    uint32_t sync_below_iters = 2048;   /* see render_quality_locked */
    int blocks_main_f_sync = 0, blocks_main_d_sync = 0;
    uint32_t refill_smem = 0;
    CUfunction k_compose = nullptr, k_undersampled = nullptr, k_debug = nullptr;
    int blocks_main_f = 0, blocks_main_d = 0, blocks_adv_f = 0, blocks_adv_d = 0;
    /* renderer state (CudaFractalRenderer) */
    chaos_state state = CHAOS_STATE_NOT_INITIALIZED;
    uint32_t width = 0, height = 0;
    chaos_output_mode mode = CHAOS_OUTPUT_HOST;
    record_buffer buf[2];          /* DeviceMemoryDoubleBuffer2D: [0] primary, [1] secondary */
    CUdeviceptr alloc[2] = {0, 0}; /* the two buffers in allocation order; buf[0].ptr == alloc[primary_alloc] */
    uint32_t primary_alloc = 0;
    /* multi-GPU fast frames: the other ranks' two record buffers mapped into this process (chaos_ipc_open_records) */
    CUdeviceptr peer_records[CHAOS_MAX_PEERS][2] = {};
    CUdeviceptr peer_counters[CHAOS_MAX_PEERS] = {};   /* their scheduler counters: the tile cursor other ranks steal from */
    CUfunction k_wait_foreign = nullptr;
    uint32_t frame_seq = 0;                            /* quality frames rendered: what a cursor's frame_seq must say before it is stolen from */
    uint32_t steal = 1;                                /* CHAOS_STEAL=0: every rank renders exactly its own tiles */
    bool buffers_switched = false;
    bool primary_dirty = true;
    bool have_last = false;        /* lastRendering != null */
    chaos_params last;             /* lastRendering = model.copy() */
    CUdeviceptr palette = 0;
    uint32_t palette_len = 0;
    CUdeviceptr rgba_dev = 0;      /* DEVICE mode frame, or device alias of rgba_host */
    CUdeviceptr rgba_target = 0;   /* chaos_set_output_target: where compose writes instead (0 = rgba_dev) */
    uint32_t *rgba_host = nullptr; /* HOST mode: pinned + mapped */
    /* multi-GPU: where the target came from (so that it can be given back), and the frame barrier in host shared memory */
    CUdeviceptr ipc_frame = 0;     /* chaos_ipc_open_frame: a peer's device frame mapped into this process */
    void *host_target = nullptr;   /* chaos_set_host_target: caller's host memory, registered by us */
    volatile unsigned long long *barrier = nullptr;
    uint32_t barrier_world = 0;
    unsigned long long barrier_seq = 0;
    CUdeviceptr counters = 0;
    chaos_counters *counters_host = nullptr; /* pinned staging for the read-back */
    CUstream stream = nullptr;
    CUevent ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   /* render start/end, compose start/end, pass boundary,
                                                                    * pass B done, early compose start/end (stream2) */
    CUstream stream2 = nullptr;    /* the frame-wide compose of a multi-pass render runs here, next to passes C and D */
    /* Strands: a multi-pass frame is cut into `strands` interleaved sets of row bands (the multi-GPU partition, one level
     * down), each with its own stream, counters, tile order and export arrays.  Their pass chains A -> B -> C -> D are
     * independent, so the CTAs of one strand's next pass fill the SMs that the other strand's draining pass leaves idle:
     * the tail of every pass but the last is hidden.  Strand 0 runs on `stream`. */
    uint32_t strands = CHAOS_DEFAULT_STRANDS;
    uint32_t strand_min_tiles = 50000;   /* vote tiles a strand must have (CHAOS_STRAND_MIN_TILES) */
    /* threads per CTA of the persistent pass kernels (warps are independent there): a CTA gives its SM share back only
     * when its last warp is done, so smaller CTAs let the next pass in sooner */
    uint32_t pass_threads = 256;
    /* resident warps per SM of the escape-loop kernels (0 = what fits).  Fewer warps per scheduler make every orbit
     * advance faster (a trip is a chain of three dependent FP64 instructions) at some cost in pipe utilisation */
    int loop_warps_per_sm = 0;
    /* A frame that runs as several strands has several long kernels in flight, each a persistent grid sized for the whole
     * machine.  Asking for a third of an SM's shared memory per CTA (never touched) caps the long-kernel CTAs an SM holds
     * ACROSS those launches at 3: the fourth CTA's worth of registers and issue slots stays with the other kernels of the
     * strands' chains (finish, classify, the next probe ...), which the strand that is ahead needs to get on (c2 2.95 ->
     * 2.85 ms; 2 per SM: 3.03).  CHAOS_LONG_SMEM=bytes, 0 = off.  A single chain gets the uncapped grid. */
    uint32_t long_smem = 72u * 1024u;
    int blocks_long_shared[2] = {0, 0};
    /* A frame that is ONE launch (one sample per pixel, or the tile-synchronous kernel) and goes to host memory used to be
     * rendered, then composed over PCIe: 0.6 ms of a 4K frame's, 5 ms of an 8192^2 frame's end-to-end time with the GPU idle.
     * It is rendered in host_parts interleaved sets of row bands instead, one launch each, and part k is composed into host
     * memory (second stream) while part k + 1 is rendered: only the last part's compose is left over.  CHAOS_HOST_PARTS,
     * 1 = off; a part has at least 60 000 vote tiles. */
    uint32_t host_parts = 4;
    uint32_t late_by_strand = 1;   /* the exported tiles of a strand are composed when ITS pass D is over (CHAOS_LATE_BY_STRAND=0: all of them after the render) */
    /* orbit pool of the independent-orbit passes (chaos_render_args::pool): one per strand, allocated by the first frame */
    CUdeviceptr pool[CHAOS_MAX_STRANDS] = {};
    uint32_t pool_capacity = 0;
    uint32_t pool_min_lanes = 20;
    uint32_t pool_epoch = 0;
    CUstream strand_stream[CHAOS_MAX_STRANDS] = {};
    CUevent strand_ev_b[CHAOS_MAX_STRANDS] = {}, strand_ev_done[CHAOS_MAX_STRANDS] = {};
    uint32_t overlap_compose = 1;
    /* fast frames: pass R colours the pixels it finishes (chaos_render_args::fuse_rgba): 0 never, 1 when the frame goes to
     * pinned host memory (the 33 MB PCIe write then starts with the first kernel instead of the last: 0.88 -> 0.82 ms
     * end to end), 2 always (in device memory it saves 5 of 189 us and the memory passes run further from the HBM
     * roofline: 36 B per pixel in 82 us against 52 B in 87 us) */
    uint32_t fuse_fast = 1;
    int host_compose_blocks = 0;   /* CTAs of that compose: 0 = one per SM, -1 = the usual grid (CHAOS_HOST_COMPOSE_BLOCKS) */
    uint32_t part_index = 0, part_count = 1, band_rows = 64;
    chaos_stats stats;
    /* which kernels iterate: 0 = tile-synchronous with the reference's operation sequence (differential check), 1 = lane-refill
     * scheduler (render_refill.cuh), 2 = orbit streams (render_streams.cuh), 3 = by frame (default): one sample per pixel ->
     * the single refill launch (c4 35.0 ms against 36.1 with streams: nothing to sort, every orbit is long); several samples
     * with an iteration limit under sync_below_iters -> tile-synchronous kernel; several samples otherwise -> streams */
    uint32_t engine = 3;
    /* engine 3, several samples, high limit: streams pay off where many orbits are long and never escape (c2: 86 % of the
     * reference's trips are proven); where every orbit escapes after a few hundred trips (c2ex2: 0.3 %) the restart, the
     * failed group and the finish replay are a large share of each orbit and the lane-refill engine is 8 % faster.  The
     * previous multi-sample frame of this renderer says which kind of view this is (frames of a session resemble each
     * other); the first frame goes through the streams. */
    float proven_fraction = -1.f;                      /* of the last such frame; < 0: none yet */
    float proven_any = -1.f;                           /* the same of the last quality frame of any kind: below streams_above the orbits' recurrence
                                                        * check goes back to one compare per group (CHAOS_SHORTCUT_DENSE_COMPARE off; CHAOS_DENSE_COMPARE=0/1 forces) */
    int dense_compare = -1;
    float streams_above = 0.05f;                       /* CHAOS_STREAMS_ABOVE */
    uint32_t block_iters = 0;      /* 0 = choose from maxIterations */
    uint32_t shortcuts = CHAOS_SHORTCUT_DEFER_TEST | CHAOS_SHORTCUT_RECURRENCE;   /* count-preserving shortcuts of the escape loop */
    uint32_t sched_idle_indep = 10, sched_idle_rounds = 16;   /* see take_scheduling_pass (render_refill.cuh) */
};

static chaos_status check_renderer(const chaos_renderer *r)
{
    if (!r) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "renderer handle is NULL");
    if (!r->module) return fail(CHAOS_ERR_ILLEGAL_STATE, "Module has not been initialized or has been closed.");
    return CHAOS_OK;
}

/* ------------------------------------------------------------------------------------------
 * provider
 * ---------------------------------------------------------------------------------------- */
extern "C" chaos_status chaos_provider_create(const char *kernels_dir, int device, chaos_provider **out)
{
    if (!out) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (!kernels_dir) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "kernels_dir is NULL (the reference falls back to 'cudaKernels', FractalRenderingModule.java:51-55; pass it explicitly)");
    const char *err = nullptr;
    D = chaos_cuda_driver_get(&err);
    if (!D) return fail(CHAOS_ERR_CUDA_INIT, "%s", err ? err : "cannot load libcuda");
    CU_TRY(cuInit(0), CHAOS_ERR_CUDA_INIT);
    int n = 0;
    CU_TRY(cuDeviceGetCount(&n), CHAOS_ERR_CUDA_INIT);
    if (n <= 0) return fail(CHAOS_ERR_CUDA_INIT, "CUDA_ERROR_NO_DEVICE: no CUDA device is visible; this backend has no CPU fallback");
    if (device < 0 || device >= n) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "device %d out of range (0..%d)", device, n - 1);
    chaos_provider *p = new chaos_provider();
    p->kernels_dir = kernels_dir;
    p->device = device;
    CUresult r = D->p_cuDeviceGet(&p->cu_device, device);
    if (r == CUDA_SUCCESS) r = D->p_cuDevicePrimaryCtxRetain(&p->ctx, p->cu_device);
    if (r != CUDA_SUCCESS) {
        delete p;
        return fail(CHAOS_ERR_CUDA_INIT, "cannot create a context on device %d: %s", device, cu_err_name(r));
    }
    D->p_cuDeviceGetAttribute(&p->sm_count, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, p->cu_device);
    int major = 0, minor = 0;
    D->p_cuDeviceGetAttribute(&major, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR, p->cu_device);
    D->p_cuDeviceGetAttribute(&minor, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR, p->cu_device);
    if (major != 10) {
        D->p_cuDevicePrimaryCtxRelease(p->cu_device);
        delete p;
        return fail(CHAOS_ERR_CUDA_INIT, "device %d is sm_%d%d; the modules are built for sm_100a (B200) only", device, major, minor);
    }
    *out = p;
    return CHAOS_OK;
}

extern "C" chaos_status chaos_close(chaos_renderer *r);

extern "C" chaos_status chaos_provider_destroy(chaos_provider *p)
{
    if (!p) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "provider handle is NULL");
    if (p->active) { chaos_close(p->active); }
    for (registered_module *m : p->registered) delete m;
    if (p->ctx) D->p_cuDevicePrimaryCtxRelease(p->cu_device);
    delete p;
    return CHAOS_OK;
}

extern "C" chaos_renderer *chaos_active_renderer(const chaos_provider *p) { return p ? p->active : nullptr; }

extern "C" chaos_status chaos_list_fractals(chaos_provider *p, const char **names, uint32_t capacity, uint32_t *count)
{
    if (!p) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "provider handle is NULL");
    const uint32_t n = g_n_modules + (uint32_t)p->registered.size();
    if (count) *count = n;
    if (names)
        for (uint32_t i = 0; i < n && i < capacity; ++i)
            names[i] = i < g_n_modules ? g_modules[i].fractal_name : p->registered[i - g_n_modules]->desc.fractal_name;
    return CHAOS_OK;
}

static void defaults_none(chaos_defaults *d) { defaults_base(d); }

/* the counterpart of adding a Module*.java and registering it (CudaFractalRendererProvider.java:19-31): a fractal author's
 * module <kernels_dir>/<file_stem>.cubin under a display name of its own; no defaults, constants through chaos_write_constant */
extern "C" chaos_status chaos_register_module(chaos_provider *p, const char *fractal_name, const char *file_stem)
{
    if (!p) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "provider handle is NULL");
    if (!fractal_name || !*fractal_name || !file_stem || !*file_stem) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "a module needs a name and a file");
    if (strchr(file_stem, '/') || strstr(file_stem, "..")) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "the module file must lie in the kernels directory: %s", file_stem);
    for (uint32_t i = 0; i < g_n_modules; ++i)
        if (!strcmp(g_modules[i].fractal_name, fractal_name)) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "Fractal %s is already registered", fractal_name);
    for (registered_module *m : p->registered)
        if (m->name == fractal_name) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "Fractal %s is already registered", fractal_name);
    registered_module *m = new registered_module();
    m->name = fractal_name; m->stem = file_stem;
    m->desc = module_desc{m->name.c_str(), m->stem.c_str(), nullptr, custom_none, defaults_none};
    p->registered.push_back(m);
    return CHAOS_OK;
}

static chaos_status get_function(chaos_renderer *r, const char *name, CUfunction *fn)
{
    CUresult e = D->p_cuModuleGetFunction(fn, r->module, name);
    if (e == CUDA_ERROR_NOT_FOUND)
        return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "Function %s not found in module %s (CudaKernel.java:36-38)", name, r->desc->file_stem);
    if (e != CUDA_SUCCESS) return fail(CHAOS_ERR_CUDA_INIT, "cuModuleGetFunction(%s) failed: %s", name, cu_err_name(e));
    return CHAOS_OK;
}

static void unload_module(chaos_renderer *r)
{
    if (r->module) { D->p_cuModuleUnload(r->module); r->module = nullptr; }
}

static int persistent_blocks(chaos_renderer *r, CUfunction fn, int threads, size_t smem = 0, int max_warps_per_sm = 0)
{
    int per_sm = 0;
    if (D->p_cuOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem) != CUDA_SUCCESS || per_sm < 1) per_sm = 1;
    if (max_warps_per_sm > 0) per_sm = std::max(1, std::min(per_sm, max_warps_per_sm / (threads / 32)));
    return per_sm * r->provider->sm_count;   /* a whole number of CTAs per SM: 148 x resident CTAs */
}

/* FractalRenderingModule.initialize :73-100 */
static chaos_status load_module(chaos_renderer *r)
{
    std::string path = r->provider->kernels_dir + "/" + r->desc->file_stem + ".cubin";
    FILE *f = fopen(path.c_str(), "rb");
    if (!f)
        return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "Invalid module file name: %s\nHave you set the kernels directory (cudaKernelsDir) properly?", path.c_str());
    std::vector<char> image;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    image.resize(sz > 0 ? (size_t)sz : 0);
    size_t got = image.empty() ? 0 : fread(image.data(), 1, image.size(), f);
    fclose(f);
    if (got != image.size() || image.empty()) return fail(CHAOS_ERR_CUDA_INIT, "cannot read module file %s", path.c_str());
    CUresult e = D->p_cuModuleLoadData(&r->module, image.data());
    if (e != CUDA_SUCCESS) { r->module = nullptr; return fail(CHAOS_ERR_CUDA_INIT, "cuModuleLoadData(%s) failed: %s", path.c_str(), cu_err_name(e)); }

    /* all kernels are resolved eagerly; a missing one is an error (FractalRenderingModule.java:91-97) */
    struct { const char *name; CUfunction *fn; } fns[] = {
        {"fractalRenderMainFloat", &r->k_main_f}, {"fractalRenderMainDouble", &r->k_main_d},
        {"chaosPassAFloat", &r->k_pass_a[0]}, {"chaosPassADouble", &r->k_pass_a[1]},
        {"chaosPassBFloat", &r->k_pass_b[0]}, {"chaosPassBDouble", &r->k_pass_b[1]},
        {"chaosPassCFloat", &r->k_pass_c[0]}, {"chaosPassCDouble", &r->k_pass_c[1]},
        {"fractalRenderAdvancedFloat", &r->k_adv_f}, {"fractalRenderAdvancedDouble", &r->k_adv_d},
        {"fractalRenderUnderSampled", &r->k_undersampled}, {"compose", &r->k_compose}, {"debug", &r->k_debug},

        {"fractalRenderMainFloatSync", &r->k_main_f_sync}, {"fractalRenderMainDoubleSync", &r->k_main_d_sync},
        {"chaosClassifyTiles", &r->k_classify}, {"chaosOrderTiles", &r->k_order}, {"chaosReplayExported", &r->k_replay}, {"chaosExportAll", &r->k_export_all}, {"chaosWaitForeign", &r->k_wait_foreign},
        {"chaosReusePassFloat", &r->k_reuse_f}, {"chaosReusePassDouble", &r->k_reuse_d},
        {"chaosProbeFloat", &r->k_probe[0]}, {"chaosProbeDouble", &r->k_probe[1]}, {"chaosLongFloat", &r->k_long[0]},
        {"chaosLongDouble", &r->k_long[1]}, {"chaosFinishFloat", &r->k_finish[0]}, {"chaosFinishDouble", &r->k_finish[1]},
    };
    r->fn_names.clear();
    for (auto &k : fns) {
        chaos_status st = get_function(r, k.name, k.fn);
        if (st != CHAOS_OK) { unload_module(r); return st; }
        r->fn_names.emplace_back(*k.fn, k.name);
    }
    CUdeviceptr abi_ptr = 0;
    size_t abi_size = 0;
    uint32_t abi = 0;
    if (D->p_cuModuleGetGlobal(&abi_ptr, &abi_size, r->module, "CHAOS_MODULE_ABI_VERSION") != CUDA_SUCCESS || abi_size != 4 ||
        D->p_cuMemcpyDtoH(&abi, abi_ptr, 4) != CUDA_SUCCESS || abi != CHAOS_MODULE_ABI) {
        unload_module(r);
        return fail(CHAOS_ERR_CUDA_INIT, "module %s was built for launch contract %u, this library speaks %u; rebuild the module", path.c_str(), abi, CHAOS_MODULE_ABI);
    }
    /* the lane-refill kernels keep their tile slots in dynamic shared memory; the module says how much */
    CUdeviceptr smem_ptr = 0;
    size_t smem_size = 0;
    if (D->p_cuModuleGetGlobal(&smem_ptr, &smem_size, r->module, "CHAOS_REFILL_SMEM") != CUDA_SUCCESS || smem_size != 4 ||
        D->p_cuMemcpyDtoH(&r->refill_smem, smem_ptr, 4) != CUDA_SUCCESS) {
        unload_module(r);
        return fail(CHAOS_ERR_CUDA_INIT, "module %s does not export CHAOS_REFILL_SMEM", path.c_str());
    }
    for (CUfunction fn : {r->k_pass_b[0], r->k_pass_b[1]}) {
        CUresult ea = D->p_cuFuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)r->refill_smem);
        if (ea != CUDA_SUCCESS) { unload_module(r); return fail(CHAOS_ERR_CUDA_INIT, "cannot reserve %u B of shared memory: %s", r->refill_smem, cu_err_name(ea)); }
    }
    r->blocks_main_f = persistent_blocks(r, r->k_main_f, 256, 0, r->loop_warps_per_sm);
    r->blocks_main_d = persistent_blocks(r, r->k_main_d, 256, 0, r->loop_warps_per_sm);
    const int T = (int)r->pass_threads;
    r->refill_smem = r->refill_smem / 8u * (r->pass_threads / 32u);   /* the module states it for 8 warps */
    for (int p = 0; p < 2; ++p) {
        r->blocks_pass_a[p] = persistent_blocks(r, r->k_pass_a[p], T, 0, r->loop_warps_per_sm);
        r->blocks_pass_b[p] = persistent_blocks(r, r->k_pass_b[p], T, r->refill_smem);
        r->blocks_pass_c[p] = persistent_blocks(r, r->k_pass_c[p], T, 0, r->loop_warps_per_sm);
        r->blocks_probe[p] = persistent_blocks(r, r->k_probe[p], 256);
        r->blocks_long[p] = persistent_blocks(r, r->k_long[p], T, 0, r->loop_warps_per_sm);
        r->blocks_long_shared[p] = r->blocks_long[p];
        if (r->long_smem && D->p_cuFuncSetAttribute(r->k_long[p], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)r->long_smem) == CUDA_SUCCESS)
            r->blocks_long_shared[p] = persistent_blocks(r, r->k_long[p], T, r->long_smem, r->loop_warps_per_sm);
        else r->long_smem = 0;
        r->blocks_finish[p] = persistent_blocks(r, r->k_finish[p], 256);
    }
    /* the one-launch kernels run next to the compose of the part before (chaos_renderer::host_parts), which stages its palette in
     * shared memory: an SM that holds their CTAs must be set up with room for it (they use none themselves, and nothing of L1) */
    for (CUfunction fn : {r->k_main_f, r->k_main_d, r->k_main_f_sync, r->k_main_d_sync})
        D->p_cuFuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, 50);
    r->blocks_main_f_sync = persistent_blocks(r, r->k_main_f_sync, 256);
    r->blocks_main_d_sync = persistent_blocks(r, r->k_main_d_sync, 256);
    r->blocks_reuse_f = persistent_blocks(r, r->k_reuse_f, 256);
    r->blocks_reuse_d = persistent_blocks(r, r->k_reuse_d, 256);
    r->blocks_adv_f = persistent_blocks(r, r->k_adv_f, 256);
    r->blocks_adv_d = persistent_blocks(r, r->k_adv_d, 256);
    if (r->desc->on_initialize) {
        chaos_status st = r->desc->on_initialize(r);
        if (st != CHAOS_OK) { unload_module(r); return st; }
    }
    return CHAOS_OK;
}

extern "C" chaos_status chaos_open(chaos_provider *p, const char *fractal_name, int force_reload, chaos_renderer **out)
{
    if (!p) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "provider handle is NULL");
    if (!out) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "out is NULL");
    if (!fractal_name) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "Unknown fractal: null");
    ctx_guard g(p);
    if (p->active && !strcmp(p->active->desc->fractal_name, fractal_name) && !force_reload) { *out = p->active; return CHAOS_OK; }
    *out = nullptr;
    /* an unknown name leaves the active renderer alone (a C caller's old handle stays valid); from here on the old
     * renderer is closed first, as the reference does (:52), and chaos_active_renderer() tells what is left */
    const module_desc *desc = nullptr;
    for (uint32_t i = 0; i < g_n_modules; ++i)
        if (!strcmp(g_modules[i].fractal_name, fractal_name)) desc = &g_modules[i];
    for (registered_module *m : p->registered)
        if (m->name == fractal_name) desc = &m->desc;
    if (!desc) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "Unknown fractal: %s", fractal_name);
    if (p->active) chaos_close(p->active);                         /* closing the previous renderer :52 */
    chaos_renderer *r = new chaos_renderer();
    r->provider = p;
    r->desc = desc;
    memset(&r->stats, 0, sizeof r->stats);
    r->stats.struct_size = sizeof(chaos_stats);
    /* debugging knobs (not part of the reference's interface): engine 0 is the differential check of engine 1 */
    const char *eng = getenv("CHAOS_ENGINE");
    if (eng) r->engine = (uint32_t)std::min(std::max(atoi(eng), 0), 3);
    const char *sab = getenv("CHAOS_STREAMS_ABOVE");
    if (sab) r->streams_above = (float)atof(sab);
    const char *stl = getenv("CHAOS_STEAL");
    if (stl) r->steal = (uint32_t)atoi(stl) ? 1u : 0u;
    const char *lo = getenv("CHAOS_LONG_OCC");
    if (lo) { unsigned x = 0, y = 0, z = 0; if (sscanf(lo, "%u,%u,%u", &x, &y, &z) == 3) { r->long_occ[0] = x; r->long_occ[1] = y; r->long_occ[2] = z; } }
    const char *hf = getenv("CHAOS_HOT_FIRST");
    if (hf) r->hot_first = (uint32_t)atoi(hf) & 3u;
    r->timeline_path = getenv("CHAOS_TIMELINE");
    const char *dcp = getenv("CHAOS_DENSE_COMPARE");
    if (dcp) r->dense_compare = atoi(dcp) ? 1 : 0;
    const char *lsh = getenv("CHAOS_LIST_SHRINK");
    if (lsh) r->list_shrink = (uint32_t)std::max(atoi(lsh), 1);
    const char *prb = getenv("CHAOS_PROBE_TRIPS");
    if (prb) r->probe_trips = (uint32_t)std::max(atoi(prb), 8);
    const char *sb = getenv("CHAOS_SYNC_BELOW");
    if (sb) r->sync_below_iters = (uint32_t)atoi(sb);
    const char *nb = getenv("CHAOS_BLOCK_ITERS");
    if (nb) r->block_iters = ((uint32_t)atoi(nb) + 3u) & ~3u;
    const char *si = getenv("CHAOS_SCHED_IDLE");   /* "indep,rounds" */
    if (si) {
        unsigned x = 0, y = 0;
        if (sscanf(si, "%u,%u", &x, &y) == 2 && x >= 1 && y >= 1) { r->sched_idle_indep = x; r->sched_idle_rounds = y; }
    }
    const char *ff = getenv("CHAOS_FUSE_FAST");   /* 0 = compose colours every pixel of a fast frame */
    if (ff) r->fuse_fast = (uint32_t)std::min(std::max(atoi(ff), 0), 2);
    const char *hb = getenv("CHAOS_HOST_COMPOSE_BLOCKS");
    if (hb) r->host_compose_blocks = atoi(hb);
    const char *oc = getenv("CHAOS_OVERLAP_COMPOSE");   /* 0 = compose only after the last render pass */
    if (oc) r->overlap_compose = (uint32_t)atoi(oc) ? 1u : 0u;
    const char *ea = getenv("CHAOS_EXPORT_ALL_BELOW");
    if (ea) r->export_all_below = atoi(ea);
    const char *ex = getenv("CHAOS_EXPORT");      /* 0 = every tile keeps all its rounds in pass B */
    if (ex) r->export_enabled = (uint32_t)atoi(ex) ? 1u : 0u;
    const char *smt = getenv("CHAOS_STRAND_MIN_TILES");
    if (smt) r->strand_min_tiles = (uint32_t)std::max(atoi(smt), 0);
    const char *sn = getenv("CHAOS_STRANDS");     /* 1 = the passes of a multi-sample frame run one after the other */
    if (sn) r->strands = (uint32_t)std::min(std::max(atoi(sn), 1), CHAOS_MAX_STRANDS);
    const char *pm = getenv("CHAOS_POOL_MIN");    /* 0 = orbits never change warps */
    if (pm) r->pool_min_lanes = (uint32_t)std::min(std::max(atoi(pm), 0), 32);
    const char *pt = getenv("CHAOS_PASS_THREADS");
    if (pt && (atoi(pt) == 32 || atoi(pt) == 64 || atoi(pt) == 128 || atoi(pt) == 256)) r->pass_threads = (uint32_t)atoi(pt);
    const char *lw = getenv("CHAOS_LOOP_WARPS_PER_SM");
    if (lw) r->loop_warps_per_sm = std::max(atoi(lw), 0);
    const char *lbs = getenv("CHAOS_LATE_BY_STRAND");
    if (lbs) r->late_by_strand = atoi(lbs) ? 1u : 0u;
    const char *hps = getenv("CHAOS_HOST_PARTS");
    if (hps) r->host_parts = (uint32_t)std::min(std::max(atoi(hps), 1), CHAOS_MAX_STRANDS);
    const char *lsm = getenv("CHAOS_LONG_SMEM");
    if (lsm) r->long_smem = (uint32_t)std::max(atoi(lsm), 0);
    const char *sc = getenv("CHAOS_SHORTCUTS");   /* 0 = every trip executed and tested, as the reference does */
    if (sc) r->shortcuts = (uint32_t)atoi(sc) & (CHAOS_SHORTCUT_DEFER_TEST | CHAOS_SHORTCUT_RECURRENCE);
    chaos_status st = load_module(r);
    if (st != CHAOS_OK) { delete r; return st; }
    CUresult e = D->p_cuStreamCreate(&r->stream, CU_STREAM_NON_BLOCKING);
    if (e == CUDA_SUCCESS) e = D->p_cuStreamCreate(&r->stream2, CU_STREAM_NON_BLOCKING);
    for (int i = 0; i < 8 && e == CUDA_SUCCESS; ++i) e = D->p_cuEventCreate(&r->ev[i], CU_EVENT_DEFAULT);
    r->strand_stream[0] = r->stream;
    for (uint32_t i = 0; i < CHAOS_MAX_STRANDS && e == CUDA_SUCCESS; ++i) {
        if (i) e = D->p_cuStreamCreate(&r->strand_stream[i], CU_STREAM_NON_BLOCKING);
        if (e == CUDA_SUCCESS) e = D->p_cuEventCreate(&r->strand_ev_b[i], CU_EVENT_DISABLE_TIMING);
        if (e == CUDA_SUCCESS) e = D->p_cuEventCreate(&r->strand_ev_done[i], CU_EVENT_DISABLE_TIMING);
    }
    if (e == CUDA_SUCCESS) e = D->p_cuMemAlloc(&r->counters, sizeof(chaos_counters) * CHAOS_MAX_STRANDS);
    if (e == CUDA_SUCCESS) e = D->p_cuMemHostAlloc((void **)&r->counters_host, sizeof(chaos_counters) * CHAOS_MAX_STRANDS, 0);
    if (e != CUDA_SUCCESS) {
        st = fail(CHAOS_ERR_CUDA_INIT, "cannot create stream/events/counters: %s", cu_err_name(e));
        p->active = r;
        chaos_close(r);
        return st;
    }
    p->active = r;
    *out = r;
    return CHAOS_OK;
}

/* ------------------------------------------------------------------------------------------
 * lifecycle
 * ---------------------------------------------------------------------------------------- */
static void free_export(chaos_renderer *r)
{
    chaos_export &x = r->exp_buf;
    if (x.tile) D->p_cuMemFree((CUdeviceptr)x.tile);
    if (x.first) D->p_cuMemFree((CUdeviceptr)x.first);
    if (x.et) D->p_cuMemFree((CUdeviceptr)x.et);
    if (x.iters) D->p_cuMemFree((CUdeviceptr)x.iters);
    if (x.skipped) D->p_cuMemFree((CUdeviceptr)x.skipped);
    x = chaos_export{0, nullptr, nullptr, nullptr, nullptr, nullptr};
}

/* arrays for the tiles pass B hands to pass C; allocated by the first multi-sample render of a frame size */
static bool ensure_export(chaos_renderer *r, uint32_t n_tiles)
{
    chaos_export &x = r->exp_buf;
    if (x.capacity >= n_tiles) return true;
    D->p_cuStreamSynchronize(r->stream);
    free_export(r);
    CUdeviceptr p[5] = {0, 0, 0, 0, 0};
    const size_t bytes[5] = {(size_t)n_tiles * 4u, (size_t)n_tiles * 4u, (size_t)n_tiles * CHAOS_EXPORT_ROUNDS * 32u * 4u,
                             (size_t)n_tiles * CHAOS_EXPORT_ROUNDS * 8u, (size_t)n_tiles * CHAOS_EXPORT_ROUNDS * 8u};
    for (int i = 0; i < 5; ++i) {
        if (D->p_cuMemAlloc(&p[i], bytes[i]) != CUDA_SUCCESS) {
            for (int j = 0; j < i; ++j) D->p_cuMemFree(p[j]);
            return false;
        }
    }
    x.capacity = n_tiles;
    x.tile = (uint32_t *)p[0]; x.first = (uint32_t *)p[1]; x.et = (uint32_t *)p[2];
    x.iters = (unsigned long long *)p[3]; x.skipped = (unsigned long long *)p[4];
    return true;
}

/* the pool of strand s; 0 if it cannot be had (the passes then run without) */
static CUdeviceptr ensure_pool(chaos_renderer *r, uint32_t s)
{
    if (!r->pool_min_lanes) return 0;
    if (!r->pool_capacity) {
        int most = std::max(r->blocks_main_f, r->blocks_main_d) * 256;
        for (int p = 0; p < 2; ++p)
            most = std::max(most, std::max(std::max(r->blocks_pass_a[p], r->blocks_pass_c[p]), r->blocks_long[p]) * (int)r->pass_threads);
        const uint32_t warps = (uint32_t)most / 32u;   /* of the largest launch */
        r->pool_capacity = CHAOS_POOL_SHARDS * 32u * ((warps + CHAOS_POOL_SHARDS - 1u) / CHAOS_POOL_SHARDS);
    }
    if (!r->pool[s]) {
        const size_t bytes = (size_t)r->pool_capacity * CHAOS_POOL_STRIDE;
        if (D->p_cuMemAlloc(&r->pool[s], bytes) != CUDA_SUCCESS) r->pool[s] = 0;
        else if (D->p_cuMemsetD8Async(r->pool[s], 0, bytes, r->stream) != CUDA_SUCCESS) { D->p_cuMemFree(r->pool[s]); r->pool[s] = 0; }   /* no tag matches */
    }
    return r->pool[s];
}

/* engine 2: the long and finish lists, `entries` each, 32 bytes per entry (CHAOS_LIST_STRIDE) */
static bool ensure_lists(chaos_renderer *r, size_t entries)
{
    if (r->list_capacity >= entries) return true;
    D->p_cuStreamSynchronize(r->stream);
    if (r->long_list) { D->p_cuMemFree(r->long_list); r->long_list = 0; }
    if (r->finish_list) { D->p_cuMemFree(r->finish_list); r->finish_list = 0; }
    r->list_capacity = 0;
    if (D->p_cuMemAlloc(&r->long_list, entries * 32u) != CUDA_SUCCESS) { r->long_list = 0; return false; }
    if (D->p_cuMemAlloc(&r->finish_list, entries * 32u) != CUDA_SUCCESS) { D->p_cuMemFree(r->long_list); r->long_list = 0; r->finish_list = 0; return false; }
    r->list_capacity = entries;
    return true;
}

/* engine 2: one probe -> long -> finish chain on stream q; b.phase says which pass it serves */
static chaos_status launch(chaos_renderer *r, CUfunction fn, int blocks, int threads, unsigned smem, void *arg, CUstream stream);
static chaos_status launch_stream_chain(chaos_renderer *r, chaos_render_args &b, int p, CUstream q, bool shared_machine = false)
{
    chaos_status st = launch(r, r->k_probe[p], r->blocks_probe[p], 256, 0, &b, q);
    if (st == CHAOS_OK) st = shared_machine && r->long_smem ? launch(r, r->k_long[p], r->blocks_long_shared[p], (int)r->pass_threads, r->long_smem, &b, q)
                                                            : launch(r, r->k_long[p], r->blocks_long[p], (int)r->pass_threads, 0, &b, q);
    if (st == CHAOS_OK) st = launch(r, r->k_finish[p], r->blocks_finish[p], 256, 0, &b, q);
    return st;
}

static void release_peer_records(chaos_renderer *r)
{
    for (uint32_t q = 0; q < CHAOS_MAX_PEERS; ++q) {
        for (int i = 0; i < 2; ++i)
            if (r->peer_records[q][i]) { D->p_cuIpcCloseMemHandle(r->peer_records[q][i]); r->peer_records[q][i] = 0; }
        if (r->peer_counters[q]) { D->p_cuIpcCloseMemHandle(r->peer_counters[q]); r->peer_counters[q] = 0; }
    }
}

static void release_targets(chaos_renderer *r)
{
    if (r->ipc_frame) { D->p_cuIpcCloseMemHandle(r->ipc_frame); r->ipc_frame = 0; }
    if (r->host_target) { D->p_cuMemHostUnregister(r->host_target); r->host_target = nullptr; }
    r->rgba_target = 0;
}

static void free_frame_memory(chaos_renderer *r)
{
    release_peer_records(r);
    for (int i = 0; i < 2; ++i) if (r->buf[i].ptr) { D->p_cuMemFree(r->buf[i].ptr); r->buf[i].ptr = 0; r->buf[i].pitch = 0; r->alloc[i] = 0; }
    if (r->palette) { D->p_cuMemFree(r->palette); r->palette = 0; }
    if (r->tile_key) { D->p_cuMemFree(r->tile_key); r->tile_key = 0; }
    if (r->tile_order) { D->p_cuMemFree(r->tile_order); r->tile_order = 0; }
    free_export(r);
    if (r->long_list) { D->p_cuMemFree(r->long_list); r->long_list = 0; }
    if (r->finish_list) { D->p_cuMemFree(r->finish_list); r->finish_list = 0; }
    r->list_capacity = 0;
    if (r->late_tiles) { D->p_cuMemFree(r->late_tiles); r->late_tiles = 0; }
    if (r->warp_trace) { D->p_cuMemFree(r->warp_trace); r->warp_trace = 0; }
    release_targets(r);
    if (r->rgba_host) { D->p_cuMemFreeHost(r->rgba_host); r->rgba_host = nullptr; r->rgba_dev = 0; }
    if (r->rgba_dev) { D->p_cuMemFree(r->rgba_dev); r->rgba_dev = 0; }
}

extern "C" chaos_status chaos_initialize(chaos_renderer *r, uint32_t width, uint32_t height,
                                         const uint32_t *palette_rgba, uint32_t palette_len, chaos_output_mode mode)
{
    chaos_status st = check_renderer(r);
    if (st != CHAOS_OK) return st;
    if (r->state == CHAOS_STATE_READY_TO_RENDER) return fail(CHAOS_ERR_ILLEGAL_STATE, "Already initialized.");
    if (width == 0 || height == 0) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "output size must be positive but is %ux%u", width, height);
    if (!palette_rgba || palette_len == 0) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "palette must hold at least one colour");
    if (palette_len > 12000) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "palette of %u entries does not fit the compose kernel's shared-memory stage (max 12000)", palette_len);
    /* the reference refuses grids above 65535 blocks of 32 px (CudaFractalRenderer.java:248-253) */
    if (width > 65535u * 32u) return fail(CHAOS_ERR_RENDERER, "Unsupported input parameter: width must be smaller than %u", 65535u * 32u);
    if (height > 65535u * 32u) return fail(CHAOS_ERR_RENDERER, "Unsupported input parameter: height must be smaller than %u", 65535u * 32u);
    if (mode != CHAOS_OUTPUT_HOST && mode != CHAOS_OUTPUT_DEVICE) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "unknown output mode %d", (int)mode);
    ctx_guard g(r->provider);
    free_frame_memory(r);
    /* two pitched buffers of 16-byte records (DeviceMemoryDoubleBuffer2D.java:116-146) */
    for (int i = 0; i < 2; ++i) {
        CUresult e = D->p_cuMemAllocPitch(&r->buf[i].ptr, &r->buf[i].pitch, (size_t)width * 16u, height, 16);
        if (e != CUDA_SUCCESS) { free_frame_memory(r); return fail(CHAOS_ERR_CUDA, "cuMemAllocPitch(%ux%u) failed: %s", width, height, cu_err_name(e)); }
        r->alloc[i] = r->buf[i].ptr;
    }
    r->primary_alloc = 0;
    const size_t all_tiles = (size_t)((width + 7u) / 8u) * ((height + 3u) / 4u);
    CUresult e = D->p_cuMemAlloc(&r->tile_key, all_tiles * 4u);
    if (e == CUDA_SUCCESS) e = D->p_cuMemAlloc(&r->tile_order, all_tiles * 4u);
    if (e == CUDA_SUCCESS) e = D->p_cuMemAlloc(&r->late_tiles, ((all_tiles + 31u) / 32u) * 4u);
    if (e == CUDA_SUCCESS) e = D->p_cuMemAlloc(&r->palette, (size_t)palette_len * 4u);
    if (e == CUDA_SUCCESS) e = D->p_cuMemcpyHtoD(r->palette, palette_rgba, (size_t)palette_len * 4u);
    size_t frame_bytes = (size_t)width * height * 4u;
    if (e == CUDA_SUCCESS) {
        if (mode == CHAOS_OUTPUT_HOST) {
            e = D->p_cuMemHostAlloc((void **)&r->rgba_host, frame_bytes, CU_MEMHOSTALLOC_DEVICEMAP);
            if (e == CUDA_SUCCESS) e = D->p_cuMemHostGetDevicePointer(&r->rgba_dev, r->rgba_host, 0);
            if (e == CUDA_SUCCESS) memset(r->rgba_host, 0, frame_bytes);
        } else {
            e = D->p_cuMemAlloc(&r->rgba_dev, frame_bytes);
            if (e == CUDA_SUCCESS) e = D->p_cuMemsetD8Async(r->rgba_dev, 0, frame_bytes, r->stream);
            if (e == CUDA_SUCCESS) e = D->p_cuStreamSynchronize(r->stream);
        }
    }
    if (e != CUDA_SUCCESS) { free_frame_memory(r); return fail(CHAOS_ERR_CUDA, "cannot allocate palette/output: %s", cu_err_name(e)); }
    r->width = width; r->height = height; r->mode = mode; r->palette_len = palette_len;
    r->buffers_switched = false;
    r->primary_dirty = true;                                       /* reallocatePrimary2DBuffer :45-49 */
    r->state = CHAOS_STATE_READY_TO_RENDER;
    return CHAOS_OK;
}

extern "C" chaos_status chaos_free_resources(chaos_renderer *r)
{
    chaos_status st = check_renderer(r);
    if (st != CHAOS_OK) return st;
    if (r->state == CHAOS_STATE_NOT_INITIALIZED) return fail(CHAOS_ERR_ILLEGAL_STATE, "Already free.");
    ctx_guard g(r->provider);
    D->p_cuStreamSynchronize(r->stream);
    free_frame_memory(r);
    r->state = CHAOS_STATE_NOT_INITIALIZED;
    return CHAOS_OK;
}

extern "C" chaos_status chaos_close(chaos_renderer *r)
{
    if (!r) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "renderer handle is NULL");
    ctx_guard g(r->provider);
    if (r->stream) D->p_cuStreamSynchronize(r->stream);
    free_frame_memory(r);
    for (uint32_t i = 0; i < CHAOS_MAX_STRANDS; ++i) if (r->pool[i]) D->p_cuMemFree(r->pool[i]);
    if (r->counters) D->p_cuMemFree(r->counters);
    if (r->counters_host) D->p_cuMemFreeHost(r->counters_host);
    for (int i = 0; i < 8; ++i) if (r->ev[i]) D->p_cuEventDestroy(r->ev[i]);
    for (auto &te : r->timeline) { D->p_cuEventDestroy(te.a); D->p_cuEventDestroy(te.b); }
    r->timeline.clear();
    if (r->stream2) { D->p_cuStreamSynchronize(r->stream2); D->p_cuStreamDestroy(r->stream2); }
    for (uint32_t i = 0; i < CHAOS_MAX_STRANDS; ++i) {
        if (i && r->strand_stream[i]) { D->p_cuStreamSynchronize(r->strand_stream[i]); D->p_cuStreamDestroy(r->strand_stream[i]); }
        if (r->strand_ev_b[i]) D->p_cuEventDestroy(r->strand_ev_b[i]);
        if (r->strand_ev_done[i]) D->p_cuEventDestroy(r->strand_ev_done[i]);
    }
    if (r->stream) D->p_cuStreamDestroy(r->stream);
    unload_module(r);
    if (r->provider && r->provider->active == r) r->provider->active = nullptr;
    delete r;
    return CHAOS_OK;
}

extern "C" chaos_state chaos_get_state(const chaos_renderer *r) { return r ? r->state : CHAOS_STATE_NOT_INITIALIZED; }
extern "C" uint32_t chaos_get_width(const chaos_renderer *r) { return r ? r->width : 0; }
extern "C" uint32_t chaos_get_height(const chaos_renderer *r) { return r ? r->height : 0; }
extern "C" const char *chaos_fractal_name(const chaos_renderer *r) { return r ? r->desc->fractal_name : ""; }
extern "C" const uint32_t *chaos_output_rgba(const chaos_renderer *r) { return r ? r->rgba_host : nullptr; }
extern "C" chaos_status chaos_set_output_target(chaos_renderer *r, uint64_t device_ptr)
{
    chaos_status st = check_renderer(r);
    if (st != CHAOS_OK) return st;
    if (r->state != CHAOS_STATE_READY_TO_RENDER) return fail(CHAOS_ERR_ILLEGAL_STATE, "Renderer has to be initialized first");
    if (r->mode != CHAOS_OUTPUT_DEVICE) return fail(CHAOS_ERR_ILLEGAL_STATE, "an output target needs CHAOS_OUTPUT_DEVICE mode");
    if (device_ptr & 15u) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "the output target must be 16-byte aligned");
    ctx_guard g(r->provider);
    D->p_cuStreamSynchronize(r->stream);
    release_targets(r);
    r->rgba_target = (CUdeviceptr)device_ptr;
    return CHAOS_OK;
}

static chaos_status check_device_target(chaos_renderer *r)
{
    chaos_status st = check_renderer(r);
    if (st != CHAOS_OK) return st;
    if (r->state != CHAOS_STATE_READY_TO_RENDER) return fail(CHAOS_ERR_ILLEGAL_STATE, "Renderer has to be initialized first");
    if (r->mode != CHAOS_OUTPUT_DEVICE) return fail(CHAOS_ERR_ILLEGAL_STATE, "an output target needs CHAOS_OUTPUT_DEVICE mode");
    return CHAOS_OK;
}

extern "C" chaos_status chaos_ipc_export_frame(chaos_renderer *r, chaos_ipc_handle *out)
{
    chaos_status st = check_device_target(r);
    if (st != CHAOS_OK) return st;
    if (!out) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "out is NULL");
    static_assert(sizeof(CUipcMemHandle) == sizeof(chaos_ipc_handle), "IPC handle size");
    ctx_guard g(r->provider);
    CUresult e = D->p_cuIpcGetMemHandle((CUipcMemHandle *)out, r->rgba_dev);
    if (e != CUDA_SUCCESS) return fail(CHAOS_ERR_CUDA, "cuIpcGetMemHandle failed: %s", cu_err_name(e));
    return CHAOS_OK;
}

extern "C" chaos_status chaos_ipc_open_frame(chaos_renderer *r, const chaos_ipc_handle *frame)
{
    chaos_status st = check_device_target(r);
    if (st != CHAOS_OK) return st;
    if (!frame) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "frame handle is NULL");
    ctx_guard g(r->provider);
    D->p_cuStreamSynchronize(r->stream);
    release_targets(r);
    CUipcMemHandle h;
    memcpy(&h, frame, sizeof h);
    CUresult e = D->p_cuIpcOpenMemHandle(&r->ipc_frame, h, CU_IPC_MEM_LAZY_ENABLE_PEER_ACCESS);
    if (e != CUDA_SUCCESS) { r->ipc_frame = 0; return fail(CHAOS_ERR_CUDA, "cuIpcOpenMemHandle failed: %s", cu_err_name(e)); }
    r->rgba_target = r->ipc_frame;
    return CHAOS_OK;
}

extern "C" chaos_status chaos_ipc_export_records(chaos_renderer *r, chaos_ipc_handle out[3])
{
    chaos_status st = check_renderer(r);
    if (st != CHAOS_OK) return st;
    if (r->state != CHAOS_STATE_READY_TO_RENDER) return fail(CHAOS_ERR_ILLEGAL_STATE, "Renderer has to be initialized first");
    if (!out) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "out is NULL");
    ctx_guard g(r->provider);
    for (int i = 0; i < 3; ++i) {
        CUresult e = D->p_cuIpcGetMemHandle((CUipcMemHandle *)&out[i], i < 2 ? r->alloc[i] : r->counters);
        if (e != CUDA_SUCCESS) return fail(CHAOS_ERR_CUDA, "cuIpcGetMemHandle failed: %s", cu_err_name(e));
    }
    return CHAOS_OK;
}

extern "C" chaos_status chaos_ipc_open_records(chaos_renderer *r, uint32_t peer_rank, const chaos_ipc_handle in[3])
{
    chaos_status st = check_renderer(r);
    if (st != CHAOS_OK) return st;
    if (r->state != CHAOS_STATE_READY_TO_RENDER) return fail(CHAOS_ERR_ILLEGAL_STATE, "Renderer has to be initialized first");
    if (!in || peer_rank >= CHAOS_MAX_PEERS) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "peer rank %u out of range (0..%d)", peer_rank, CHAOS_MAX_PEERS - 1);
    ctx_guard g(r->provider);
    D->p_cuStreamSynchronize(r->stream);
    for (int i = 0; i < 3; ++i) {
        CUdeviceptr *slot = i < 2 ? &r->peer_records[peer_rank][i] : &r->peer_counters[peer_rank];
        if (*slot) { D->p_cuIpcCloseMemHandle(*slot); *slot = 0; }
        CUipcMemHandle h;
        memcpy(&h, &in[i], sizeof h);
        CUresult e = D->p_cuIpcOpenMemHandle(slot, h, CU_IPC_MEM_LAZY_ENABLE_PEER_ACCESS);
        if (e != CUDA_SUCCESS) { *slot = 0; return fail(CHAOS_ERR_CUDA, "cuIpcOpenMemHandle(%s of rank %u) failed: %s", i < 2 ? "records" : "counters", peer_rank, cu_err_name(e)); }
    }
    return CHAOS_OK;
}

extern "C" chaos_status chaos_set_host_target(chaos_renderer *r, void *host_frame, size_t bytes)
{
    chaos_status st = check_device_target(r);
    if (st != CHAOS_OK) return st;
    ctx_guard g(r->provider);
    D->p_cuStreamSynchronize(r->stream);
    release_targets(r);
    if (!host_frame) return CHAOS_OK;
    if (bytes < (size_t)r->width * r->height * 4u) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "the host target must hold width * height * 4 bytes but holds %zu", bytes);
    if ((uintptr_t)host_frame & 4095u) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "the host target must be page-aligned");
    CUresult e = D->p_cuMemHostRegister(host_frame, bytes, CU_MEMHOSTREGISTER_DEVICEMAP | CU_MEMHOSTREGISTER_PORTABLE);
    if (e != CUDA_SUCCESS) return fail(CHAOS_ERR_CUDA, "cuMemHostRegister(%zu bytes) failed: %s", bytes, cu_err_name(e));
    CUdeviceptr dp = 0;
    e = D->p_cuMemHostGetDevicePointer(&dp, host_frame, 0);
    if (e != CUDA_SUCCESS) { D->p_cuMemHostUnregister(host_frame); return fail(CHAOS_ERR_CUDA, "cuMemHostGetDevicePointer failed: %s", cu_err_name(e)); }
    r->host_target = host_frame;
    r->rgba_target = dp;
    return CHAOS_OK;
}

extern "C" chaos_status chaos_set_frame_barrier(chaos_renderer *r, void *shm_block, uint32_t world)
{
    chaos_status st = check_renderer(r);
    if (st != CHAOS_OK) return st;
    if (shm_block && world == 0) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "a frame barrier needs world >= 1");
    if ((uintptr_t)shm_block & 7u) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "the barrier block must be 8-byte aligned");
    r->barrier = (volatile unsigned long long *)shm_block;
    r->barrier_world = shm_block ? world : 0u;
    r->barrier_seq = 0;
    return CHAOS_OK;
}

/* end of a render call under a frame barrier: this rank's kernels are done (the stream was synchronised), so its bands
 * are in the shared frame; announce it and wait for the others */
static chaos_status frame_barrier(chaos_renderer *r)
{
    if (!r->barrier) return CHAOS_OK;
    r->barrier_seq += 1;
    __atomic_fetch_add((unsigned long long *)r->barrier, 1ull, __ATOMIC_SEQ_CST);
    const unsigned long long want = r->barrier_seq * r->barrier_world;
    struct timespec t0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (unsigned spins = 0; __atomic_load_n((unsigned long long *)r->barrier, __ATOMIC_SEQ_CST) < want; ++spins) {
        if ((spins & 255u) == 255u) {
            struct timespec t;
            clock_gettime(CLOCK_MONOTONIC, &t);
            if ((t.tv_sec - t0.tv_sec) > 10) return fail(CHAOS_ERR_CUDA, "frame barrier: %llu of %llu announcements after 10 s (a rank died or renders a different number of frames)",
                                                        __atomic_load_n((unsigned long long *)r->barrier, __ATOMIC_SEQ_CST), want);
            sched_yield();
        }
    }
    return CHAOS_OK;
}

extern "C" uint64_t chaos_output_rgba_device(const chaos_renderer *r) { return (r && r->mode == CHAOS_OUTPUT_DEVICE) ? (uint64_t)r->rgba_dev : 0; }

/* ------------------------------------------------------------------------------------------
 * constants by name (FractalRenderingModule.java:163-227)
 * ---------------------------------------------------------------------------------------- */
static chaos_status write_constant(chaos_renderer *r, const char *symbol, const void *data, size_t bytes, const char *what)
{
    CUdeviceptr ptr = 0;
    size_t size = 0;
    CUresult e = D->p_cuModuleGetGlobal(&ptr, &size, r->module, symbol);
    if (e == CUDA_ERROR_NOT_FOUND) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "module %s has no constant named %s", r->desc->file_stem, symbol);
    if (e != CUDA_SUCCESS) return fail(CHAOS_ERR_CUDA, "cuModuleGetGlobal(%s) failed: %s", symbol, cu_err_name(e));
    if (size < bytes) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "Attempt to write %s to device memory allocated to size %zu", what, size);
    e = D->p_cuMemcpyHtoD(ptr, data, bytes);
    if (e != CUDA_SUCCESS) return fail(CHAOS_ERR_CUDA, "write to constant %s failed: %s", symbol, cu_err_name(e));
    return CHAOS_OK;
}

extern "C" chaos_status chaos_write_constant(chaos_renderer *r, const char *symbol, const void *data, size_t bytes)
{
    chaos_status st = check_renderer(r);
    if (st != CHAOS_OK) return st;
    if (!symbol || !data) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "symbol/data is NULL");
    ctx_guard g(r->provider);
    D->p_cuStreamSynchronize(r->stream);
    char what[32];
    snprintf(what, sizeof what, "%zu bytes", bytes);
    return write_constant(r, symbol, data, bytes, what);
}

extern "C" chaos_status chaos_set_custom_params(chaos_renderer *r, const char *text)
{
    chaos_status st = check_renderer(r);
    if (st != CHAOS_OK) return st;
    ctx_guard g(r->provider);
    D->p_cuStreamSynchronize(r->stream);
    return r->desc->set_custom_params(r, text);
}

extern "C" chaos_status chaos_supply_defaults(chaos_renderer *r, chaos_defaults *out)
{
    if (!r) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "renderer handle is NULL");
    if (!out || out->struct_size != sizeof(chaos_defaults)) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "chaos_defaults.struct_size mismatch");
    memset(out, 0, sizeof *out);
    out->struct_size = sizeof(chaos_defaults);
    r->desc->supply_defaults(out);
    return CHAOS_OK;
}

extern "C" chaos_status chaos_set_partition(chaos_renderer *r, uint32_t part_index, uint32_t part_count, uint32_t band_rows)
{
    if (!r) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "renderer handle is NULL");
    if (part_count == 0 || part_index >= part_count) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "partition %u of %u is not valid", part_index, part_count);
    if (band_rows == 0 || (band_rows & 3u)) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "band_rows must be a positive multiple of 4 (vote tiles are 4 rows high) but is %u", band_rows);
    r->part_index = part_index; r->part_count = part_count; r->band_rows = band_rows;
    r->primary_dirty = true;
    return CHAOS_OK;
}

/* ------------------------------------------------------------------------------------------
 * precision rule (CudaFractalRenderer.java:409-419, RenderingKernel.java:124-138)
 * ---------------------------------------------------------------------------------------- */
static double ulp_of_float(float v)          /* Math.ulp(float) */
{
    v = fabsf(v);
    if (v != v || isinf(v)) return (double)v;
    float n = nextafterf(v, INFINITY);
    if (isinf(n)) return (double)(v - nextafterf(v, 0.f));
    return (double)(n - v);
}
static double ulp_of_double(double v)        /* Math.ulp(double) */
{
    v = fabs(v);
    if (v != v || isinf(v)) return v;
    double n = nextafter(v, INFINITY);
    if (isinf(n)) return v - nextafter(v, 0.0);
    return n - v;
}
static chaos_precision choose_precision(const double s[4], uint32_t W, uint32_t H)
{
    double pw = fabs(s[2] - s[0]) / (double)W, ph = fabs(s[3] - s[1]) / (double)H;
    chaos_precision p = CHAOS_PRECISION_SINGLE;
    if (pw < ulp_of_float((float)s[0]) || ph < ulp_of_float((float)s[1])) p = CHAOS_PRECISION_DOUBLE;
    if (pw < ulp_of_double(s[0]) || ph < ulp_of_double(s[1])) p = CHAOS_PRECISION_TOO_BIG;
    return p;
}

/* ------------------------------------------------------------------------------------------
 * frames
 * ---------------------------------------------------------------------------------------- */
static chaos_status validate_model(const chaos_renderer *r, const chaos_params *m)
{
    if (r->state != CHAOS_STATE_READY_TO_RENDER) return fail(CHAOS_ERR_ILLEGAL_STATE, "Renderer has to be initialized first");
    if (!m || m->struct_size != sizeof(chaos_params)) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "chaos_params.struct_size mismatch");
    if (m->max_iterations < 1) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "maxIterations must be a positive number, but is : %d", m->max_iterations);
    static const char *names[4] = {"left_bottom_x", "left_bottom_y", "right_top_x", "right_top_y"};
    for (int i = 0; i < 4; ++i)
        if (!isfinite(m->segment[i])) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "Argument segment %s must be a finite float but is %g.", names[i], m->segment[i]);
    return CHAOS_OK;
}

/* tile rows of the bands b with b % part_count == part_index */
static uint32_t owned_tile_rows(uint32_t tile_rows, uint32_t band_tile_rows, uint32_t part_index, uint32_t part_count)
{
    if (part_count <= 1u) return tile_rows;
    uint32_t owned = 0;
    const uint32_t bands = (tile_rows + band_tile_rows - 1u) / band_tile_rows;
    for (uint32_t b = part_index; b < bands; b += part_count)
        owned += std::min(band_tile_rows, tile_rows - b * band_tile_rows);
    return owned;
}

static void fill_render_args(const chaos_renderer *r, const chaos_params *m, chaos_render_args *a)
{
    memset(a, 0, sizeof *a);
    a->counters = (chaos_counters *)r->counters;
    for (int i = 0; i < 4; ++i) { a->image[i] = m->segment[i]; a->imagef[i] = (float)m->segment[i]; }
    a->width = r->width; a->height = r->height;
    a->max_iter = (uint32_t)m->max_iterations;
    a->max_ss = m->max_super_sampling;
    /* flag word: KernelMain.java:12-13, KernelAdvanced.java:17-20 */
    a->flags = (m->use_adaptive_super_sampling ? CHAOS_FLAG_ADAPTIVE_SS : 0u) | (m->visualise_sample_count ? 2u : 0u) |
               (m->use_foveated_rendering ? CHAOS_FLAG_FOVEATION : 0u) | (m->use_sample_reuse ? CHAOS_FLAG_SAMPLE_REUSE : 0u) |
               (m->is_zooming ? CHAOS_FLAG_IS_ZOOMING : 0u) | (m->is_zooming_in ? CHAOS_FLAG_ZOOMING_IN : 0u);
    a->focus_x = (uint32_t)m->mouse_focus[0]; a->focus_y = (uint32_t)m->mouse_focus[1];
    {   /* pixel radius at which the visual angle reaches the foveal threshold (fractalRendererGeneric.cu:233-239);
         * pass R only uses it to skip the exact evaluation far away from that circle */
        double thr = m->max_super_sampling >= 1.0f ? 5.5 : (double)m->max_super_sampling * 5.5;
        double rpx = tan(thr * 0.017453292519943295) * 60.0 / 0.02652;
        a->focus_d2_thr = (float)(rpx * rpx);
    }
    a->tiles_x = (r->width + 7u) / 8u;
    a->tile_rows = (r->height + 3u) / 4u;
    a->part_index = r->part_index; a->part_count = r->part_count; a->band_tile_rows = r->band_rows / 4u;
    a->n_tiles = owned_tile_rows(a->tile_rows, a->band_tile_rows, r->part_index, r->part_count) * a->tiles_x;
    a->tile_key = (uint32_t *)r->tile_key;
    a->tile_order = (uint32_t *)r->tile_order;
    a->engine = r->engine ? 1u : 0u;
    /* trips between scheduling points: long enough to amortise a scheduling pass, short enough that a lane whose
     * orbit ended does not idle long; orbits are at most max_iter long */
    a->block_iters = r->block_iters ? r->block_iters : 128u;
    a->shortcuts = r->shortcuts;
    if ((r->shortcuts & CHAOS_SHORTCUT_RECURRENCE) &&
        (r->dense_compare >= 0 ? r->dense_compare != 0 : !(r->proven_any >= 0.f && r->proven_any < r->streams_above)))
        a->shortcuts |= CHAOS_SHORTCUT_DENSE_COMPARE;
    a->sched_idle_lanes_indep = r->sched_idle_indep;
    a->sched_idle_lanes_rounds = r->sched_idle_rounds;
    a->sm_count = (uint32_t)r->provider->sm_count;
    for (int i = 0; i < 3; ++i) a->occ_orbits_per_lane[i] = r->long_occ[i];

}

static chaos_status launch(chaos_renderer *r, CUfunction fn, int blocks, int threads, unsigned smem, void *arg, CUstream stream)
{
    void *params[1] = {arg};
    CUstream q = stream ? stream : r->stream;
    chaos_renderer::timeline_entry *te = nullptr;
    if (r->timeline_path) {
        if (r->timeline_used == r->timeline.size()) {
            chaos_renderer::timeline_entry n = {nullptr, 0, nullptr, nullptr};
            if (D->p_cuEventCreate(&n.a, CU_EVENT_DEFAULT) == CUDA_SUCCESS && D->p_cuEventCreate(&n.b, CU_EVENT_DEFAULT) == CUDA_SUCCESS) r->timeline.push_back(n);
        }
        if (r->timeline_used < r->timeline.size()) {
            te = &r->timeline[r->timeline_used++];
            te->name = "?";
            for (auto &k : r->fn_names) if (k.first == fn) te->name = k.second;
            te->stream = q == r->stream ? 0 : q == r->stream2 ? 9 : 1;
            for (int i = 0; i < CHAOS_MAX_STRANDS; ++i) if (i && q == r->strand_stream[i]) te->stream = i;
            D->p_cuEventRecord(te->a, q);
        }
    }
    CUresult e = D->p_cuLaunchKernel(fn, (unsigned)blocks, 1, 1, (unsigned)threads, 1, 1, smem, q, params, nullptr);
    if (te) D->p_cuEventRecord(te->b, q);
    if (e != CUDA_SUCCESS) return fail(CHAOS_ERR_CUDA, "Error just after launching a kernel:%s", cu_err_name(e));
    r->stats.kernel_launches += 1;
    r->stats.launches_total += 1;
    return CHAOS_OK;
}

/* launchDrawingKernel :275-364 without the GL map/unmap and surface objects */
static void fill_compose_args(chaos_renderer *r, const chaos_params *m, chaos_compose_args &c);

/* only_tiles: compose just the tiles flagged there (the ones pass D finished after the frame-wide compose had started) */
static chaos_status launch_compose(chaos_renderer *r, const chaos_params *m, CUstream stream = nullptr, const uint32_t *only_tiles = nullptr,
                                   uint32_t part_index = 0, uint32_t part_count = 0)
{
    chaos_compose_args c;
    fill_compose_args(r, m, c);
    c.only_tiles = only_tiles;
    if (part_count) { c.part_index = part_index; c.part_count = part_count; }      /* (one part of a frame rendered in parts, below) */
    c.tiles_x = (r->width + 7u) / 8u;
    uint64_t quads = (uint64_t)((r->width + 3u) / 4u) * r->height;
    int max_blocks = r->provider->sm_count * 8;
    /* The frame-wide compose that runs next to passes C and D into pinned host memory moves at PCIe speed (33 MB: 0.6 ms)
     * whatever its grid; a small grid leaves the SMs (and their register files: a pass C CTA needs a quarter of one) to
     * the passes it runs next to. */
    if (stream == r->stream2 && stream && ((r->mode == CHAOS_OUTPUT_HOST && !r->rgba_target) || r->host_target) && r->host_compose_blocks >= 0)
        max_blocks = r->host_compose_blocks ? r->host_compose_blocks : r->provider->sm_count;   /* measured: 16 / 37 / 148 / 1184 CTAs -> c2 4.13 / 3.99 / 3.91 / 3.93 ms, c2ex2 7.63 / 7.61 / 7.70 / 7.91 ms end to end */
    int blocks = (int)std::min<uint64_t>((quads + 255u) / 256u, (uint64_t)max_blocks);
    if (blocks < 1) blocks = 1;
    return launch(r, r->k_compose, blocks, 256, r->palette_len * 4u, &c, stream);
}

static void fill_compose_args(chaos_renderer *r, const chaos_params *m, chaos_compose_args &c)
{
    memset(&c, 0, sizeof c);
    c.in = (const chaos_pixel_info *)r->buf[0].ptr;
    c.in_pitch = r->buf[0].pitch;
    c.out_rgba = (uint32_t *)(r->rgba_target ? r->rgba_target : r->rgba_dev);
    c.palette = (const uint32_t *)r->palette;
    c.palette_len = r->palette_len;
    c.width = r->width; c.height = r->height;
    c.max_ss = m->max_super_sampling;
    c.part_index = r->part_index; c.part_count = r->part_count; c.band_rows = r->band_rows;
}

static chaos_status finish_frame(chaos_renderer *r)
{
    CUresult e = D->p_cuMemcpyDtoHAsync(r->counters_host, r->counters, sizeof(chaos_counters) * CHAOS_MAX_STRANDS, r->stream);
    if (e == CUDA_SUCCESS) e = D->p_cuStreamSynchronize(r->stream);
    if (e != CUDA_SUCCESS) return fail(CHAOS_ERR_CUDA, "Error just after launching a kernel:%s", cu_err_name(e));
    D->p_cuEventElapsedTime(&r->stats.render_ms, r->ev[0], r->ev[1]);
    D->p_cuEventElapsedTime(&r->stats.compose_ms, r->ev[2], r->ev[3]);
    D->p_cuEventElapsedTime(&r->stats.frame_ms, r->ev[0], r->ev[3]);
    if (r->timeline_path && r->timeline_used) {       /* (diagnostics) the last frame's launches, overwritten every frame */
        D->p_cuCtxSynchronize();
        if (FILE *f = fopen(r->timeline_path, "w")) {
            for (size_t i = 0; i < r->timeline_used; ++i) {
                float t0 = 0.f, t1 = 0.f;
                D->p_cuEventElapsedTime(&t0, r->ev[0], r->timeline[i].a);
                D->p_cuEventElapsedTime(&t1, r->ev[0], r->timeline[i].b);
                fprintf(f, "%-28s %d %8.4f %8.4f\n", r->timeline[i].name, r->timeline[i].stream, t0, t1);
            }
            fclose(f);
        }
    }
    r->timeline_used = 0;
    r->stats.reuse_ms = 0.f;
    r->stats.pixel_iterations = r->stats.samples = r->stats.skipped_iterations = 0;
    r->stats.foreign_orbits = r->counters_host[0].foreign_done;
    for (uint32_t s = 0; s < CHAOS_MAX_STRANDS; ++s)
        if (r->counters_host[s].abort) {     /* a kernel gave up waiting in the orbit pool (bounded spins): the records are not to be trusted */
            r->primary_dirty = true;
            return fail(CHAOS_ERR_CUDA, "a render kernel timed out waiting for an orbit hand-over (strand %u); the frame is void", s);
        }
    for (uint32_t s = 0; s < CHAOS_MAX_STRANDS; ++s) {      /* every strand of the frame counts into its own block */
        r->stats.pixel_iterations += r->counters_host[s].pixel_iterations;
        r->stats.samples += r->counters_host[s].samples;
        r->stats.skipped_iterations += r->counters_host[s].skipped_iterations;
    }
    if (getenv("CHAOS_LANE_STATS")) {   /* diagnostics of modules built with -DCHAOS_LANE_STATS */
        for (uint32_t s = 0; s < CHAOS_MAX_STRANDS; ++s)
            if (r->counters_host[s].next_tile)
                fprintf(stderr, "strand %u: tiles after sample 1: to pass B %u, exported (by the classifier or pass B) %u\n", s,
                        r->counters_host[s].n_continuing, r->counters_host[s].n_exported);
        for (uint32_t s = 0; s < CHAOS_MAX_STRANDS; ++s) for (int w = 0; w < 2; ++w) {      /* the long kernels' timeline (engine 2) */
            unsigned long long *v = &r->counters_host[s].lane_stats[3][1][w * 4];
            if (!v[3]) continue;
            const double t0 = (double)~v[0];
            fprintf(stderr, "strand %u long kernel %c: first warp saw the list dry after %.3f ms, last warp after %.3f ms, last warp ended after %.3f ms\n",
                    s, w ? 'C' : 'A', v[1] ? ((double)~v[1] - t0) * 1e-6 : -1.0, ((double)v[2] - t0) * 1e-6, ((double)v[3] - t0) * 1e-6);
            v[0] = v[1] = v[2] = v[3] = 0ull;
            unsigned long long *u = &r->counters_host[s].lane_stats[3][0][w * 4];
            if (u[0]) fprintf(stderr, "strand %u long kernel %c: %llu orbits ran all the way unproven: mean %.3f ms, longest %.3f ms, the last one started after %.3f ms\n",
                              s, w ? 'C' : 'A', u[0], (double)u[1] / (double)u[0] * 1e-6, (double)u[2] * 1e-6, (double)u[3] * 1e-6);
            u[0] = u[1] = u[2] = u[3] = 0ull;
            unsigned long long *h = &r->counters_host[s].lane_stats[1][w][0];
            fprintf(stderr, "strand %u long kernel %c: orbits that ran (nearly) all the way, by when they were taken from the list (0.131 ms bins): %llu %llu %llu %llu %llu %llu %llu %llu\n",
                    s, w ? 'C' : 'A', h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
            for (int k = 0; k < 8; ++k) h[k] = 0ull;
            fprintf(stderr, "strand %u long kernel %c over time: warps alive at 0, 65.5, 131 ... us:", s, w ? 'C' : 'A');
            for (int k = 0; k < 32; ++k) if (r->counters_host[s].long_hist[w][0][k]) fprintf(stderr, " %llu", r->counters_host[s].long_hist[w][0][k]);
            fprintf(stderr, "\n");
            {
                unsigned long long *c = &r->counters_host[s].long_hist[w][3][0];
                if (c[1]) fprintf(stderr, "   inside the loop, list not dry: %.1f cycles per trip (%llu calls); dry: %.1f cycles per trip (%llu calls)\n",
                                  (double)c[0] / (double)c[1], c[2], c[5] ? (double)c[4] / (double)c[5] : 0.0, c[6]);
            }
        }
        static const char *pass[4] = {"A", "B", "C", "main"}, *kind[2] = {"tested", "untested"};
        for (int p = 0; p < 4; ++p) for (int t = 0; t < 2; ++t) {
            unsigned long long v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (uint32_t s = 0; s < CHAOS_MAX_STRANDS; ++s) for (int k = 0; k < 8; ++k) v[k] += r->counters_host[s].lane_stats[p][t][k];
            if (!v[CHAOS_LS_CAPACITY]) continue;
            const double c = (double)v[CHAOS_LS_CAPACITY];   /* every lane added the block's length */
            fprintf(stderr, "lane_stats pass %-4s %-8s lane-trips %.4g useful %.3f replay_wait %.3f fin_wait %.3f idle_queue %.3f idle_dry %.3f (mid-block ends %.3f) blocks %llu passes %llu\n",
                    pass[p], kind[t], c, v[CHAOS_LS_USEFUL] / c, v[CHAOS_LS_REPLAY_WAIT] / c, v[CHAOS_LS_FIN_WAIT] / c,
                    v[CHAOS_LS_IDLE_QUEUE] / c, v[CHAOS_LS_IDLE_DRY] / c,
                    1.0 - (double)(v[CHAOS_LS_USEFUL] + v[CHAOS_LS_REPLAY_WAIT] + v[CHAOS_LS_FIN_WAIT] + v[CHAOS_LS_IDLE_QUEUE] + v[CHAOS_LS_IDLE_DRY]) / c,
                    v[CHAOS_LS_BLOCKS], v[CHAOS_LS_PASSES]);
        }
    }
    return frame_barrier(r);
}

static chaos_status set_module_constants(chaos_renderer *r, const chaos_params *m)
{
    /* setModuleConstants :227-233: rewritten only when the flag changes */
    if (!r->have_last || (m->visualise_sample_count != 0) != (r->last.visualise_sample_count != 0)) {
        int v = m->visualise_sample_count ? 1 : 0;   /* host copies symbol-size bytes of a little-endian int */
        D->p_cuStreamSynchronize(r->stream);
        return write_constant(r, "VISUALIZE_SAMPLE_COUNT", &v, 1, "boolean");
    }
    return CHAOS_OK;
}

static chaos_precision frame_precision(const chaos_renderer *r, chaos_params *m)
{
    chaos_precision p = choose_precision(m->segment, r->width, r->height);
    m->float_precision = (int32_t)p;                               /* model.setFloatingPointPrecision :418 */
    if (m->force_precision == 1) return CHAOS_PRECISION_SINGLE;
    if (m->force_precision == 2) return CHAOS_PRECISION_DOUBLE;
    return p;
}

static bool peers_open(const chaos_renderer *r);

static chaos_status render_quality_locked(chaos_renderer *r, chaos_params *m)
{
    chaos_status st = set_module_constants(r, m);
    if (st != CHAOS_OK) return st;
    /* the main kernel asserts maxSuperSampling >= 1 on the device (fractalRendererGeneric.cu:174) */
    if (!(m->max_super_sampling >= 1.0f))
        return fail(CHAOS_ERR_RENDERER, "maxSuperSampling must be >= 1 for a quality render but is %g (device assert, fractalRendererGeneric.cu:174)", m->max_super_sampling);
    chaos_precision prec = frame_precision(r, m);
    const bool dbl = prec != CHAOS_PRECISION_SINGLE;               /* tooBig still runs the double kernel :213-217 */
    if (r->buffers_switched) { std::swap(r->buf[0], r->buf[1]); r->buffers_switched = false; r->primary_alloc = 0; }  /* resetBufferOrder */
    chaos_render_args a;
    fill_render_args(r, m, &a);
    a.out = (chaos_pixel_info *)r->buf[0].ptr; a.out_pitch = r->buf[0].pitch;
    r->stats.kernel_launches = 0;
    CUresult e = D->p_cuMemsetD8Async(r->counters, 0, sizeof(chaos_counters) * CHAOS_MAX_STRANDS, r->stream);
    if (e != CUDA_SUCCESS) return fail(CHAOS_ERR_CUDA, "cuMemsetD8Async failed: %s", cu_err_name(e));
    D->p_cuEventRecord(r->ev[0], r->stream);
    bool early_compose = false, profile_frame = false, parts_composed = false, late_composed = false;
    if (a.n_tiles) {
        const int p = dbl ? 1 : 0;
        const uint32_t S0 = (uint32_t)std::min(64.0f, roundf(m->max_super_sampling));
        /* Short orbits (low iteration limit) with several samples: the per-orbit scheduling work of the refill
         * engine costs more than the divergence it removes, so those frames take the tile-synchronous kernel
         * (same arithmetic, same records).  CHAOS_ENGINE=0 forces it, with the reference's 7-operation trip. */
        const bool low_limit = S0 >= 2u && a.max_iter < r->sync_below_iters;
        const bool sync_kernel = r->engine == 0 || ((r->engine == 1 || r->engine == 3) && low_limit);
        a.force_exact = r->engine == 0 ? 1u : 0u;
        const bool streams = r->engine == 2 || (r->engine == 3 && S0 >= 2u && !low_limit &&
                                                !(r->proven_fraction >= 0.f && r->proven_fraction < r->streams_above));
        profile_frame = S0 >= 2u && !low_limit;
        a.probe_trips = r->probe_trips;
        if (streams && !ensure_lists(r, (size_t)((a.tiles_x * (size_t)a.tile_rows)) * (S0 <= 1u ? 32u : 64u)))
            return fail(CHAOS_ERR_CUDA, "cannot allocate the orbit lists of a %ux%u frame", r->width, r->height);
        /* one-launch frames on their way to host memory: in parts, composed as they come (chaos_renderer::host_parts) */
        const bool to_host = (r->mode == CHAOS_OUTPUT_HOST && !r->rgba_target) || r->host_target;
        const bool one_launch = sync_kernel || (S0 <= 1u && !streams);
        uint32_t K = (to_host && one_launch && !(r->steal && peers_open(r))) ? r->host_parts : 1u;
        K = (uint32_t)std::min<uint64_t>(K, std::max<uint64_t>(a.n_tiles / 60000u, 1u));
        {
            const uint32_t bands = (a.tile_rows + a.band_tile_rows - 1u) / a.band_tile_rows;
            while (K > 1u && (uint64_t)bands < (uint64_t)r->part_count * K * 2u) --K;       /* at least two bands per part */
        }
        if (K > 1u) {
            for (uint32_t k = 0; k < K && st == CHAOS_OK; ++k) {
                chaos_render_args b = a;
                b.part_count = r->part_count * K; b.part_index = r->part_index + r->part_count * k;
                b.n_tiles = owned_tile_rows(b.tile_rows, b.band_tile_rows, b.part_index, b.part_count) * b.tiles_x;
                b.counters = (chaos_counters *)r->counters + k;
                if (!b.n_tiles) continue;
                /* (one CTA per SM fewer than fit: the compose of the part before needs a CTA slot next to this launch's persistent grid) */
                const int sms = r->provider->sm_count;
                if (sync_kernel) {
                    const int blocks = dbl ? r->blocks_main_d_sync : r->blocks_main_f_sync;
                    st = launch(r, dbl ? r->k_main_d_sync : r->k_main_f_sync, blocks >= 2 * sms ? blocks - sms : blocks, 256, 0, &b, r->stream);
                } else {
                    b.pool = (unsigned char *)ensure_pool(r, 0);
                    b.pool_capacity = r->pool_capacity; b.pool_min_lanes = r->pool_min_lanes; b.pool_epoch = (r->pool_epoch += 2u) & 0xffffffu;
                    const int blocks = dbl ? r->blocks_main_d : r->blocks_main_f;
                    st = launch(r, dbl ? r->k_main_d : r->k_main_f, blocks >= 2 * sms ? blocks - sms : blocks, 256, 0, &b, r->stream);
                }
                D->p_cuEventRecord(r->strand_ev_b[k], r->stream);
                D->p_cuStreamWaitEvent(r->stream2, r->strand_ev_b[k], 0);
                if (k == 0u) D->p_cuEventRecord(r->ev[6], r->stream2);
                if (st == CHAOS_OK) st = launch_compose(r, m, r->stream2, nullptr, b.part_index, b.part_count);
            }
            D->p_cuEventRecord(r->ev[7], r->stream2);
            parts_composed = st == CHAOS_OK;
        } else if (sync_kernel) {
            st = launch(r, dbl ? r->k_main_d_sync : r->k_main_f_sync, dbl ? r->blocks_main_d_sync : r->blocks_main_f_sync, 256, 0, &a, r->stream);
        } else if (S0 <= 1u && streams) {
            a.long_list = (void *)r->long_list; a.finish_list = (void *)r->finish_list;
            a.list_capacity = (uint32_t)std::min<size_t>((size_t)a.n_tiles * 32u, r->list_capacity) / r->list_shrink;
            a.pool = (unsigned char *)ensure_pool(r, 0);
            a.pool_capacity = r->pool_capacity; a.pool_min_lanes = r->pool_min_lanes; a.pool_epoch = (r->pool_epoch += 2u) & 0xffffffu;
            a.phase = 0u;
            if ((r->hot_first & 2u) && r->list_shrink == 1u) a.hot_capacity = a.list_capacity;
            st = launch_stream_chain(r, a, p, r->stream);
        } else if (S0 <= 1u) {
            a.pool = (unsigned char *)ensure_pool(r, 0);
            a.pool_capacity = r->pool_capacity; a.pool_min_lanes = r->pool_min_lanes; a.pool_epoch = (r->pool_epoch += 2u) & 0xffffffu;
            /* several GPUs: a rank whose own tiles are handed out goes on with the other ranks' (chaos_render_args::steal_world) */
            const bool stealing = r->steal && peers_open(r);
            if (stealing) {
                a.steal_world = r->part_count; a.steal_rank = r->part_index; a.frame_seq = ++r->frame_seq;
                for (uint32_t q = 0; q < r->part_count; ++q) {
                    a.steal_n_tiles[q] = owned_tile_rows(a.tile_rows, a.band_tile_rows, q, r->part_count) * a.tiles_x;
                    a.steal_counters[q] = (chaos_counters *)(q == r->part_index ? r->counters : r->peer_counters[q]);
                    a.steal_out[q] = (chaos_pixel_info *)(q == r->part_index ? r->buf[0].ptr : r->peer_records[q][0]);   /* a quality frame goes to the first buffer */
                }
                const uint32_t bands = (r->height + r->band_rows - 1u) / r->band_rows;
                for (uint32_t bnd = r->part_index; bnd < bands; bnd += r->part_count)
                    a.own_pixels += (unsigned long long)std::min(r->band_rows, r->height - bnd * r->band_rows) * r->width;
            }
            st = launch(r, dbl ? r->k_main_d : r->k_main_f, dbl ? r->blocks_main_d : r->blocks_main_f, 256, 0, &a, r->stream);
            if (stealing && st == CHAOS_OK) {
                void *params[1] = {&a};
                CUresult le = D->p_cuLaunchKernel(r->k_wait_foreign, 1, 1, 1, 32, 1, 1, 0, r->stream, params, nullptr);
                if (le != CUDA_SUCCESS) st = fail(CHAOS_ERR_CUDA, "Error just after launching a kernel:%s", cu_err_name(le));
            }
        } else {
            /* pass A: sample 0 of every pixel; classify + order: expected-longest tiles first; pass B: the other rounds,
             * except those of tiles set to use their whole budget -> pass C (independent orbits) + pass D (their decisions).
             * The frame runs as G strands (see chaos_renderer::strands): the same chain over interleaved sets of row bands,
             * one stream each, so one strand's next pass fills the SMs another strand's draining pass leaves idle. */
            const char *trace_path = getenv("CHAOS_WARP_TRACE");
            const uint32_t bands = (a.tile_rows + a.band_tile_rows - 1u) / a.band_tile_rows;
            uint32_t G = trace_path ? 1u : r->strands;
            while (G > 1u && (uint64_t)bands < (uint64_t)r->part_count * G * 2u) --G;   /* at least two bands per strand */
            /* strands overlap the pass chains' tails at the price of twice the launches; a rank of a 4- or 8-GPU frame has so
             * few tiles that its kernels are short and the launches are what it waits for (one rank of 8: 0.91 against 0.94 ms) */
            while (G > 1u && (uint64_t)a.n_tiles < (uint64_t)r->strand_min_tiles * G) --G;
            const bool exporting = r->export_enabled && S0 >= 3u && S0 <= CHAOS_EXPORT_ROUNDS && ensure_export(r, a.n_tiles);
            if (exporting && r->overlap_compose) {
                a.late_tiles = (uint32_t *)r->late_tiles;
                const size_t frame_tiles = (size_t)a.tiles_x * a.tile_rows;
                if (D->p_cuMemsetD32Async(r->late_tiles, 0u, (frame_tiles + 31u) / 32u, r->stream) != CUDA_SUCCESS)
                    return fail(CHAOS_ERR_CUDA, "cuMemsetD32Async failed");
            }
            for (uint32_t s = 0; s < G; ++s) ensure_pool(r, s);      /* (a new pool is cleared on r->stream) */
            /* (chaos_renderer::long_smem) not when the frame is composed into host memory next to passes C and D: that compose moves at
             * PCIe speed from its first CTA on, and an SM set up for the long kernels' shared memory takes it in late (c2 end to end 3.19 -> 3.51 ms) */
            const bool capped = G > 1u && !((r->mode == CHAOS_OUTPUT_HOST && !r->rgba_target) || r->host_target);
            if (G > 1u) D->p_cuEventRecord(r->ev[4], r->stream);     /* counters, bitmap and pools are clear */
            uint32_t tile_base = 0;
            a.pool_epoch = (r->pool_epoch += 2u) & 0xffffffu;
            for (uint32_t s = 0; s < G && st == CHAOS_OK; ++s) {
                CUstream q = r->strand_stream[s];
                chaos_render_args b = a;
                b.pool = (unsigned char *)ensure_pool(r, s);
                b.pool_capacity = r->pool_capacity; b.pool_min_lanes = r->pool_min_lanes;
                if (G > 1u) {
                    b.part_index = r->part_index + r->part_count * s;
                    b.part_count = r->part_count * G;
                    b.n_tiles = owned_tile_rows(a.tile_rows, a.band_tile_rows, b.part_index, b.part_count) * a.tiles_x;
                    b.counters = (chaos_counters *)r->counters + s;
                    b.tile_key = (uint32_t *)r->tile_key + tile_base;
                    b.tile_order = (uint32_t *)r->tile_order + tile_base;
                    if (s) D->p_cuStreamWaitEvent(q, r->ev[4], 0);
                }
                if (exporting) {
                    b.exp = r->exp_buf;
                    b.exp.capacity = b.n_tiles;
                    b.exp.tile += tile_base; b.exp.first += tile_base;
                    b.exp.et += (size_t)tile_base * CHAOS_EXPORT_ROUNDS * 32u;
                    b.exp.iters += (size_t)tile_base * CHAOS_EXPORT_ROUNDS; b.exp.skipped += (size_t)tile_base * CHAOS_EXPORT_ROUNDS;
                }
                if (streams) {           /* this strand's slice of the orbit lists: two orbits per pixel of its tiles */
                    b.long_list = (void *)(r->long_list + (size_t)tile_base * 64u * 32u);
                    b.finish_list = (void *)(r->finish_list + (size_t)tile_base * 64u * 32u);
                    b.list_capacity = b.n_tiles * 64u / r->list_shrink;
                }
                tile_base += b.n_tiles;
                if (b.n_tiles) {
                    const int small_grid = (int)std::min<uint64_t>((b.n_tiles + 255u) / 256u, (uint64_t)r->provider->sm_count * 4u);
                    const int tile_grid = (int)std::min<uint64_t>((b.n_tiles + 7u) / 8u, (uint64_t)r->provider->sm_count * 8u);   /* one warp per tile */
                    b.phase = 1u;
                    if (streams && (r->hot_first & 2u) && r->list_shrink == 1u) b.hot_capacity = b.list_capacity;   /* the two ends cannot meet: the list holds every orbit of pass A */
                    st = streams ? launch_stream_chain(r, b, p, q, capped) : launch(r, r->k_pass_a[p], r->blocks_pass_a[p], (int)r->pass_threads, 0, &b, q);
                    if (st == CHAOS_OK) st = launch(r, r->k_classify, tile_grid, 256, 0, &b, q);
                    if (st == CHAOS_OK) st = launch(r, r->k_order, small_grid, 256, 0, &b, q);
                    b.phase = 2u;
                    b.export_all_below = r->export_all_below >= 0 ? (uint32_t)r->export_all_below
                                                                  : (uint32_t)r->blocks_pass_b[p] * (r->pass_threads / 32u) * 4u * 2u;
                    const size_t trace_bytes = (size_t)r->blocks_pass_b[p] * (r->pass_threads / 32u) * 8u * sizeof(unsigned long long);
                    if (trace_path) {
                        if (!r->warp_trace && D->p_cuMemAlloc(&r->warp_trace, trace_bytes) != CUDA_SUCCESS) r->warp_trace = 0;
                        if (r->warp_trace) { D->p_cuMemsetD8Async(r->warp_trace, 0, trace_bytes, q); b.warp_trace = (unsigned long long *)r->warp_trace; }
                    }
                    if (exporting && st == CHAOS_OK) {
                        b.export_all_done = 1u;
                        st = launch(r, r->k_export_all, small_grid, 256, 0, &b, q);
                    }
                    if (st == CHAOS_OK) st = launch(r, r->k_pass_b[p], r->blocks_pass_b[p], (int)r->pass_threads, r->refill_smem, &b, q);
                    if (trace_path && r->warp_trace && st == CHAOS_OK) {
                        std::vector<unsigned long long> host(trace_bytes / sizeof(unsigned long long));
                        D->p_cuStreamSynchronize(q);
                        if (D->p_cuMemcpyDtoH(host.data(), r->warp_trace, trace_bytes) == CUDA_SUCCESS) {
                            FILE *f = fopen(trace_path, "wb");
                            if (f) { fwrite(host.data(), 1, trace_bytes, f); fclose(f); }
                        }
                    }
                    b.warp_trace = nullptr;
                }
                D->p_cuEventRecord(r->strand_ev_b[s], q);
                if (exporting && b.n_tiles && st == CHAOS_OK) {
                    b.phase = 3u;
                    b.hot_capacity = 0u;
                    if (streams && (r->hot_first & 1u)) { b.hot_capacity = b.list_capacity / 4u; b.hot_trips = std::max(b.max_iter / 4u, 64u); }
                    st = streams ? launch_stream_chain(r, b, p, q, capped) : launch(r, r->k_pass_c[p], r->blocks_pass_c[p], (int)r->pass_threads, 0, &b, q);
                    const int replay_grid = (int)std::min<uint64_t>((b.n_tiles + 7u) / 8u, (uint64_t)r->provider->sm_count * 8u);
                    if (st == CHAOS_OK) st = launch(r, r->k_replay, replay_grid, 256, 0, &b, q);
                }
                if (s || (G > 1u && a.late_tiles)) D->p_cuEventRecord(r->strand_ev_done[s], q);
            }
            if (a.late_tiles && st == CHAOS_OK) {
                /* every tile pass B did not export is final: stream the frame out once every strand's pass B is over,
                 * next to passes C and D */
                for (uint32_t s = 0; s < G; ++s) D->p_cuStreamWaitEvent(r->stream2, r->strand_ev_b[s], 0);
                D->p_cuEventRecord(r->ev[6], r->stream2);
                st = launch_compose(r, m, r->stream2);
                /* ... and the exported tiles of a strand once ITS pass D is over, behind that compose on the same stream (it
                 * coloured them from records that were not final): the strand that is ahead does not wait for the other */
                if (G > 1u && r->late_by_strand)
                    for (uint32_t s = 0; s < G && st == CHAOS_OK; ++s) {
                        D->p_cuStreamWaitEvent(r->stream2, r->strand_ev_done[s], 0);
                        st = launch_compose(r, m, r->stream2, a.late_tiles, r->part_index + r->part_count * s, r->part_count * G);
                    }
                late_composed = G > 1u && r->late_by_strand && st == CHAOS_OK;
                D->p_cuEventRecord(r->ev[7], r->stream2);
                early_compose = st == CHAOS_OK;
            }
            for (uint32_t s = 1; s < G; ++s) D->p_cuStreamWaitEvent(r->stream, r->strand_ev_done[s], 0);   /* join */
        }
        if (st != CHAOS_OK) return st;
    }
    D->p_cuEventRecord(r->ev[1], r->stream);
    if (early_compose || parts_composed) D->p_cuStreamWaitEvent(r->stream, r->ev[7], 0);
    D->p_cuEventRecord(r->ev[2], r->stream);
    if (!parts_composed && !late_composed) st = early_compose ? launch_compose(r, m, nullptr, a.late_tiles) : launch_compose(r, m);
    if (st != CHAOS_OK) return st;
    D->p_cuEventRecord(r->ev[3], r->stream);
    st = finish_frame(r);
    if (st != CHAOS_OK) return st;
    if (early_compose || parts_composed) {
        float early = 0.f;
        D->p_cuEventElapsedTime(&early, r->ev[6], r->ev[7]);
        r->stats.compose_ms += early;
    }
    if (profile_frame && r->stats.pixel_iterations)
        r->proven_fraction = (float)((double)r->stats.skipped_iterations / (double)r->stats.pixel_iterations);
    if (r->stats.pixel_iterations && (r->shortcuts & CHAOS_SHORTCUT_RECURRENCE))
        r->proven_any = (float)((double)r->stats.skipped_iterations / (double)r->stats.pixel_iterations);
    r->last = *m; r->have_last = true;                             /* lastRendering = model.copy() */
    r->primary_dirty = false;
    m->sample_reuse_cache_dirty = 0;
    return CHAOS_OK;
}

extern "C" chaos_status chaos_render_quality(chaos_renderer *r, chaos_params *m)
{
    chaos_status st = check_renderer(r);
    if (st != CHAOS_OK) return st;
    st = validate_model(r, m);
    if (st != CHAOS_OK) return st;
    ctx_guard g(r->provider);
    return render_quality_locked(r, m);
}

/* every other rank's record buffers and counters are mapped, and the ranks run in step */
static bool peers_open(const chaos_renderer *r)
{
    if (r->part_count <= 1u || r->part_count > CHAOS_MAX_PEERS || !r->barrier) return false;
    for (uint32_t q = 0; q < r->part_count; ++q)
        if (q != r->part_index && !(r->peer_records[q][0] && r->peer_records[q][1] && r->peer_counters[q])) return false;
    return true;
}

/* multi-GPU fast frames are possible when the partition is one slab per rank, every other rank's record buffers are mapped
 * and the ranks run in step (frame barrier) */
static bool peers_ready(const chaos_renderer *r)
{
    if (r->part_count > CHAOS_MAX_PEERS || !r->barrier) return false;
    if ((uint64_t)r->band_rows * r->part_count < r->height) return false;
    for (uint32_t q = 0; q < r->part_count; ++q)
        if (q != r->part_index && !(r->peer_records[q][0] && r->peer_records[q][1])) return false;
    return true;
}

extern "C" chaos_status chaos_render_fast(chaos_renderer *r, chaos_params *m)
{
    chaos_status st = check_renderer(r);
    if (st != CHAOS_OK) return st;
    st = validate_model(r, m);
    if (st != CHAOS_OK) return st;
    ctx_guard g(r->provider);
    st = set_module_constants(r, m);
    if (st != CHAOS_OK) return st;
    /* nothing to reuse -> create it (:164-168) */
    if (m->sample_reuse_cache_dirty || r->primary_dirty || !r->have_last) return render_quality_locked(r, m);
    /* A rank that renders only its row bands has only its own rows of the previous frame, and the reprojection reads the
     * previous frame around every pixel: without the other ranks' rows there is nothing valid to reuse outside the own
     * bands, so the frame is rendered afresh (same fallback as a dirty cache). */
    const bool slabs = r->part_count > 1u && peers_ready(r);
    if (r->part_count > 1u && !slabs) return render_quality_locked(r, m);

    chaos_precision prec = frame_precision(r, m);
    const bool dbl = prec != CHAOS_PRECISION_SINGLE;
    chaos_render_args a;
    fill_render_args(r, m, &a);
    for (int i = 0; i < 4; ++i) { a.image_reused[i] = r->last.segment[i]; a.image_reusedf[i] = (float)r->last.segment[i]; }
    bool split = false, fused = false;
    a.in = (const chaos_pixel_info *)r->buf[0].ptr; a.in_pitch = r->buf[0].pitch;   /* input = primary */
    if (slabs) {     /* every rank's primary buffer is its alloc[primary_alloc]: the ranks make the same calls in the same order */
        a.slab_rows = r->band_rows;
        for (uint32_t q = 0; q < r->part_count; ++q)
            a.in_peer[q] = (const chaos_pixel_info *)(q == r->part_index ? r->buf[0].ptr : r->peer_records[q][r->primary_alloc]);
    }
    a.out = (chaos_pixel_info *)r->buf[1].ptr; a.out_pitch = r->buf[1].pitch;       /* output = secondary */
    r->stats.kernel_launches = 0;
    CUresult e = D->p_cuMemsetD8Async(r->counters, 0, sizeof(chaos_counters) * CHAOS_MAX_STRANDS, r->stream);
    if (e != CUDA_SUCCESS) return fail(CHAOS_ERR_CUDA, "cuMemsetD8Async failed: %s", cu_err_name(e));
    D->p_cuEventRecord(r->ev[0], r->stream);
    if (a.n_tiles) {
        /* pass R: reprojection of every tile that needs no new sample (memory-bound, static schedule);
         * pass S: the tiles pass R listed (foveal disc, pixels without history) with dynamic scheduling */
        CUfunction k = dbl ? r->k_adv_d : r->k_adv_f;
        const int blocks = dbl ? r->blocks_adv_d : r->blocks_adv_f;
        a.phase = 1u;
        CUfunction k_reuse = dbl ? r->k_reuse_d : r->k_reuse_f;
        int blocks_reuse = dbl ? r->blocks_reuse_d : r->blocks_reuse_f;
        unsigned smem_reuse = 0;
        if (r->fuse_fast == 2u || (r->fuse_fast == 1u && ((r->mode == CHAOS_OUTPUT_HOST && !r->rgba_target) || r->host_target))) {
            const size_t frame_tiles = (size_t)a.tiles_x * a.tile_rows;
            if (D->p_cuMemsetD32Async(r->late_tiles, 0u, (frame_tiles + 31u) / 32u, r->stream) != CUDA_SUCCESS)
                return fail(CHAOS_ERR_CUDA, "cuMemsetD32Async failed");
            a.late_tiles = (uint32_t *)r->late_tiles;
            a.fuse_rgba = (uint32_t *)(r->rgba_target ? r->rgba_target : r->rgba_dev);
            a.fuse_palette = (const uint32_t *)r->palette;
            a.fuse_palette_len = r->palette_len;
            smem_reuse = r->palette_len * 4u;
            blocks_reuse = persistent_blocks(r, k_reuse, 256, smem_reuse);
            fused = true;
        }
        st = launch(r, k_reuse, blocks_reuse, 256, smem_reuse, &a, r->stream);
        D->p_cuEventRecord(r->ev[4], r->stream);
        a.phase = 2u;
        if (st == CHAOS_OK) st = launch(r, k, blocks, 256, 0, &a, r->stream);
        if (st != CHAOS_OK) return st;
        split = true;
    }
    D->p_cuEventRecord(r->ev[1], r->stream);
    std::swap(r->buf[0], r->buf[1]);                               /* switch2DBuffers :180 */
    r->buffers_switched = !r->buffers_switched;
    r->primary_alloc ^= 1u;
    /* (Starting the frame-wide compose right after the reuse pass, next to the sampling pass, was measured: no gain --
     * with host output a fast frame is the 33 MB PCIe write, 0.63 ms of 0.84, and the sampling pass is 0.1 ms.) */
    D->p_cuEventRecord(r->ev[2], r->stream);
    st = fused ? launch_compose(r, m, nullptr, (const uint32_t *)r->late_tiles) : launch_compose(r, m);   /* fused: only pass S' tiles are left */
    if (st != CHAOS_OK) return st;
    D->p_cuEventRecord(r->ev[3], r->stream);
    st = finish_frame(r);
    if (st != CHAOS_OK) return st;
    if (split) D->p_cuEventElapsedTime(&r->stats.reuse_ms, r->ev[0], r->ev[4]);
    r->last = *m; r->have_last = true;
    return CHAOS_OK;
}

extern "C" chaos_status chaos_debug(chaos_renderer *r)
{
    chaos_status st = check_renderer(r);
    if (st != CHAOS_OK) return st;
    ctx_guard g(r->provider);
    CUresult e = D->p_cuLaunchKernel(r->k_debug, 1, 1, 1, 32, 32, 1, 0, r->stream, nullptr, nullptr);  /* grid 1x1, block 32x32 :147-154 */
    if (e == CUDA_SUCCESS) e = D->p_cuStreamSynchronize(r->stream);
    if (e != CUDA_SUCCESS) return fail(CHAOS_ERR_CUDA, "Error just after launching a kernel:%s", cu_err_name(e));
    r->stats.launches_total += 1;
    fflush(stdout);
    return CHAOS_OK;
}

extern "C" chaos_status chaos_download_rgba(chaos_renderer *r, uint32_t *dst, size_t dst_bytes)
{
    chaos_status st = check_renderer(r);
    if (st != CHAOS_OK) return st;
    if (r->state != CHAOS_STATE_READY_TO_RENDER) return fail(CHAOS_ERR_ILLEGAL_STATE, "Renderer has to be initialized first");
    size_t need = (size_t)r->width * r->height * 4u;
    if (!dst || dst_bytes < need) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "Output buffer must be at least width * height * 4 bytes long. Buffer capacity: %zu", dst_bytes);
    ctx_guard g(r->provider);
    if (r->mode == CHAOS_OUTPUT_HOST) { memcpy(dst, r->rgba_host, need); return CHAOS_OK; }
    CUresult e = D->p_cuMemcpyDtoHAsync(dst, r->rgba_dev, need, r->stream);
    if (e == CUDA_SUCCESS) e = D->p_cuStreamSynchronize(r->stream);
    if (e != CUDA_SUCCESS) return fail(CHAOS_ERR_CUDA, "cuMemcpyDtoH failed: %s", cu_err_name(e));
    return CHAOS_OK;
}

extern "C" chaos_status chaos_download_records(chaos_renderer *r, void *dst, size_t dst_bytes)
{
    chaos_status st = check_renderer(r);
    if (st != CHAOS_OK) return st;
    if (r->state != CHAOS_STATE_READY_TO_RENDER) return fail(CHAOS_ERR_ILLEGAL_STATE, "Renderer has to be initialized first");
    size_t row = (size_t)r->width * 16u, need = row * r->height;
    if (!dst || dst_bytes < need) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "Output buffer must be at least width * height * 16 bytes long. Buffer capacity: %zu", dst_bytes);
    ctx_guard g(r->provider);
    D->p_cuStreamSynchronize(r->stream);
    CUDA_MEMCPY2D c;
    memset(&c, 0, sizeof c);
    c.srcMemoryType = CU_MEMORYTYPE_DEVICE; c.srcDevice = r->buf[0].ptr; c.srcPitch = r->buf[0].pitch;
    c.dstMemoryType = CU_MEMORYTYPE_HOST; c.dstHost = dst; c.dstPitch = row;
    c.WidthInBytes = row; c.Height = r->height;
    CUresult e = D->p_cuMemcpy2D(&c);
    if (e != CUDA_SUCCESS) return fail(CHAOS_ERR_CUDA, "cuMemcpy2D failed: %s", cu_err_name(e));
    return CHAOS_OK;
}

/* diagnostics (not in the public header): the device counters as they are right now, also while a frame is running --
 * callable from another thread; tools/pool_watch.py uses it to look at a frame that does not end */
extern "C" int chaos_debug_peek_counters(chaos_renderer *r, void *dst, size_t bytes)
{
    if (!r || !r->counters || !dst) return -1;
    ctx_guard g(r->provider);
    CUstream q = nullptr;
    if (D->p_cuStreamCreate(&q, CU_STREAM_NON_BLOCKING) != CUDA_SUCCESS) return -2;
    const size_t n = std::min(bytes, sizeof(chaos_counters) * CHAOS_MAX_STRANDS);
    CUresult e = D->p_cuMemcpyDtoHAsync(dst, r->counters, n, q);
    if (e == CUDA_SUCCESS) e = D->p_cuStreamSynchronize(q);
    D->p_cuStreamDestroy(q);
    return e == CUDA_SUCCESS ? (int)sizeof(chaos_counters) : -3;
}

extern "C" chaos_status chaos_get_stats(const chaos_renderer *r, chaos_stats *out)
{
    if (!r) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "renderer handle is NULL");
    if (!out || out->struct_size != sizeof(chaos_stats)) return fail(CHAOS_ERR_ILLEGAL_ARGUMENT, "chaos_stats.struct_size mismatch");
    *out = r->stats;
    out->struct_size = sizeof(chaos_stats);
    return CHAOS_OK;
}
