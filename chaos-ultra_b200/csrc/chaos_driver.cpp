/*
 * chaos_driver.cpp -- the frame driver behind the C ABI: the reference's rendering-mode state machine and its automatic
 * quality controller, which sit in the GUI layer there and are driven by AWT events and the JOGL animator:
 *   rendering/RenderingModeFSM.java:9-155          modes Waiting / ZoomingAuto / ZoomingOnce / Moving / ProgressiveRendering
 *   rendering/GLRenderer.java:113-162              display(): zoom step, quality decision, render, onRenderingDone
 *   rendering/GLRenderer.java:200-245              determineRenderingModeQuality(), setParamsToBeRenderedIn()
 *   rendering/RenderingController.java:81-150      mouse press / release -> FSM transitions; zoomAt()
 *   rendering/RenderingController.java:264-269     onRenderingDone(): FSM step
 * A caller (the Java host through JNI/Panama, bench.py, a test) presses and releases "the mouse" and calls
 * chaos_driver_display() once per animator tick; the driver moves the plane segment, retargets maxSuperSampling so that
 * a frame takes 15 ms while zooming or moving and 30, 60, ... ms while refining, and renders through chaos_render_fast /
 * chaos_render_quality -- or through two callbacks, so that the controller can be exercised without a GPU.
 *
 * One deliberate difference: the controller's clock.  The reference measures a frame with System.currentTimeMillis()
 * truncated to int and divides by it (:239-241); on a B200 nearly every frame takes less than a millisecond, the int is
 * 0, the quotient +Inf and the sample budget jumps to 64 at once.  CHAOS_CLOCK_DEVICE (the default) feeds the
 * controller the frame's device time in float milliseconds (CUDA events around the frame's kernels, chaos_stats::
 * frame_ms), same float arithmetic otherwise; CHAOS_CLOCK_WALL_INT reproduces the reference literally, zero included.
 */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#include "../../include/chaos_ultra.h"

namespace {

thread_local char g_driver_error[256] = "";

enum { MAX_PROGRESSIVE_RENDERING_LEVEL = 6 };            /* RenderingModeFSM.java:19 */
const int kShortestFrameRenderTime = 15, kMaxFrameRenderTime = 1000;   /* GLRenderer.java:190,194 (ms) */

struct mode_fsm {
    chaos_rendering_mode current = CHAOS_MODE_WAITING, last = CHAOS_MODE_WAITING;
    int pr_lvl = 0;
    bool zooming_and_moving = false, zooming_direction = false;

    void change(chaos_rendering_mode to) { last = current; current = to; }
    void reset_state() { change(CHAOS_MODE_WAITING); zooming_and_moving = false; }                       /* :41-45 */
    void step()                                                                                          /* :47-65 */
    {
        chaos_rendering_mode next = current;
        if ((current == CHAOS_MODE_WAITING && (last == CHAOS_MODE_ZOOMING_AUTO || last == CHAOS_MODE_MOVING)) || current == CHAOS_MODE_ZOOMING_ONCE) {
            next = CHAOS_MODE_PROGRESSIVE_RENDERING;
            pr_lvl = -1;
        } else if (current == CHAOS_MODE_PROGRESSIVE_RENDERING && pr_lvl >= MAX_PROGRESSIVE_RENDERING_LEVEL) {
            next = CHAOS_MODE_WAITING;
        }
        change(next);
        if (current == CHAOS_MODE_PROGRESSIVE_RENDERING) pr_lvl = pr_lvl + 1 < MAX_PROGRESSIVE_RENDERING_LEVEL ? pr_lvl + 1 : MAX_PROGRESSIVE_RENDERING_LEVEL;
    }
    void start_zooming(bool inside, bool moving_too) { change(CHAOS_MODE_ZOOMING_AUTO); zooming_direction = inside; zooming_and_moving = moving_too; }   /* :74-86 */
    void zoom_once(bool inside) { change(CHAOS_MODE_ZOOMING_ONCE); zooming_direction = inside; zooming_and_moving = false; }                          /* :67-72 */
    void stop_zooming() { change(zooming_and_moving ? CHAOS_MODE_MOVING : CHAOS_MODE_WAITING); zooming_and_moving = false; }                           /* :88-95 */
    void start_moving() { change(CHAOS_MODE_MOVING); }                                                                                              /* :106-109 */
    void stop_moving() { last = current; if (!zooming_and_moving) current = CHAOS_MODE_WAITING; zooming_and_moving = false; }                       /* :115-121 */
    void start_progressive() { change(CHAOS_MODE_PROGRESSIVE_RENDERING); pr_lvl = 0; zooming_and_moving = false; }                                  /* :123-128 */
    bool is_zooming() const { return current == CHAOS_MODE_ZOOMING_AUTO || zooming_and_moving || current == CHAOS_MODE_ZOOMING_ONCE; }
    bool is_moving() const { return current == CHAOS_MODE_MOVING || zooming_and_moving; }
    bool is_progressive() const { return current == CHAOS_MODE_PROGRESSIVE_RENDERING; }
    bool is_waiting() const { return current == CHAOS_MODE_WAITING; }
    bool different_than_last() const { return current != last; }
};

double wall_ms()
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec * 1e3 + (double)t.tv_nsec * 1e-6;
}

}  // namespace

struct chaos_driver {
    chaos_renderer *renderer = nullptr;
    chaos_render_fn fast = nullptr, quality = nullptr;
    void *user = nullptr;
    chaos_params model;
    uint32_t canvas_w = 0, canvas_h = 0;
    int mouse[2] = {0, 0};
    mode_fsm fsm;
    bool automatic_quality = true;
    chaos_driver_clock clock = CHAOS_CLOCK_DEVICE;
    float last_frame_render_time = (float)kShortestFrameRenderTime;   /* ms; an int in the reference (:195) */
    uint64_t frames = 0;
    int last_kind = 0;
};

static chaos_status dfail(chaos_status st, const char *msg)
{
    snprintf(g_driver_error, sizeof g_driver_error, "%s", msg);
    return st;
}
extern "C" const char *chaos_driver_last_error(void) { return g_driver_error; }

/* the renderer's own entry points as callbacks; the frame's device time comes from its statistics */
static chaos_status render_with_renderer(chaos_driver *d, bool quality_frame, float *frame_ms)
{
    chaos_status st = quality_frame ? chaos_render_quality(d->renderer, &d->model) : chaos_render_fast(d->renderer, &d->model);
    if (st != CHAOS_OK) return dfail(st, chaos_last_error());
    chaos_stats s;
    s.struct_size = sizeof s;
    if (chaos_get_stats(d->renderer, &s) == CHAOS_OK) *frame_ms = s.frame_ms;
    return CHAOS_OK;
}

static chaos_status driver_new(const chaos_params *model, uint32_t w, uint32_t h, chaos_driver **out)
{
    if (!out) return dfail(CHAOS_ERR_ILLEGAL_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (!model || model->struct_size != sizeof(chaos_params)) return dfail(CHAOS_ERR_ILLEGAL_ARGUMENT, "chaos_params.struct_size mismatch");
    if (w == 0 || h == 0) return dfail(CHAOS_ERR_ILLEGAL_ARGUMENT, "the canvas must not be empty");
    chaos_driver *d = new chaos_driver();
    d->model = *model;
    d->canvas_w = w; d->canvas_h = h;
    *out = d;
    return CHAOS_OK;
}

extern "C" chaos_status chaos_driver_create(chaos_renderer *r, const chaos_params *model, chaos_driver **out)
{
    if (!r) return dfail(CHAOS_ERR_ILLEGAL_ARGUMENT, "renderer handle is NULL");
    if (chaos_get_state(r) != CHAOS_STATE_READY_TO_RENDER) return dfail(CHAOS_ERR_ILLEGAL_STATE, "Renderer has to be initialized first");
    chaos_status st = driver_new(model, chaos_get_width(r), chaos_get_height(r), out);
    if (st == CHAOS_OK) (*out)->renderer = r;
    return st;
}

extern "C" chaos_status chaos_driver_create_custom(chaos_render_fn fast, chaos_render_fn quality, void *user, const chaos_params *model,
                                                   uint32_t canvas_width, uint32_t canvas_height, chaos_driver **out)
{
    if (!fast || !quality) return dfail(CHAOS_ERR_ILLEGAL_ARGUMENT, "both render callbacks are needed");
    chaos_status st = driver_new(model, canvas_width, canvas_height, out);
    if (st == CHAOS_OK) { (*out)->fast = fast; (*out)->quality = quality; (*out)->user = user; }
    return st;
}

extern "C" chaos_status chaos_driver_destroy(chaos_driver *d)
{
    if (!d) return dfail(CHAOS_ERR_ILLEGAL_ARGUMENT, "driver handle is NULL");
    delete d;
    return CHAOS_OK;
}

#define DRIVER_OR_FAIL(d) if (!(d)) return dfail(CHAOS_ERR_ILLEGAL_ARGUMENT, "driver handle is NULL")

/* mousePressed / mouseReleased (RenderingController.java:81-123) */
extern "C" chaos_status chaos_driver_mouse(chaos_driver *d, int x, int y) { DRIVER_OR_FAIL(d); d->mouse[0] = x; d->mouse[1] = y; return CHAOS_OK; }
extern "C" chaos_status chaos_driver_start_zooming(chaos_driver *d, int inside, int moving_too) { DRIVER_OR_FAIL(d); d->fsm.start_zooming(inside != 0, moving_too != 0); return CHAOS_OK; }
extern "C" chaos_status chaos_driver_zoom_once(chaos_driver *d, int inside) { DRIVER_OR_FAIL(d); d->fsm.zoom_once(inside != 0); return CHAOS_OK; }
extern "C" chaos_status chaos_driver_stop_zooming(chaos_driver *d) { DRIVER_OR_FAIL(d); d->fsm.stop_zooming(); return CHAOS_OK; }
extern "C" chaos_status chaos_driver_start_moving(chaos_driver *d) { DRIVER_OR_FAIL(d); d->fsm.start_moving(); return CHAOS_OK; }
extern "C" chaos_status chaos_driver_stop_moving(chaos_driver *d) { DRIVER_OR_FAIL(d); d->fsm.stop_moving(); return CHAOS_OK; }
/* the timer fired 100 ms after a release (:103-106), and startProgressiveRenderingAsync (:257-261) */
extern "C" chaos_status chaos_driver_start_progressive_rendering(chaos_driver *d, int reset_first)
{
    DRIVER_OR_FAIL(d);
    if (reset_first) d->fsm.reset_state();
    d->fsm.start_progressive();
    return CHAOS_OK;
}
extern "C" chaos_status chaos_driver_step(chaos_driver *d) { DRIVER_OR_FAIL(d); d->fsm.step(); return CHAOS_OK; }
extern "C" chaos_status chaos_driver_set_automatic_quality(chaos_driver *d, int on) { DRIVER_OR_FAIL(d); d->automatic_quality = on != 0; return CHAOS_OK; }
extern "C" chaos_status chaos_driver_set_clock(chaos_driver *d, chaos_driver_clock clock)
{
    DRIVER_OR_FAIL(d);
    if (clock != CHAOS_CLOCK_DEVICE && clock != CHAOS_CLOCK_WALL_INT) return dfail(CHAOS_ERR_ILLEGAL_ARGUMENT, "unknown clock");
    d->clock = clock;
    return CHAOS_OK;
}
extern "C" chaos_status chaos_driver_set_model(chaos_driver *d, const chaos_params *model)
{
    DRIVER_OR_FAIL(d);
    if (!model || model->struct_size != sizeof(chaos_params)) return dfail(CHAOS_ERR_ILLEGAL_ARGUMENT, "chaos_params.struct_size mismatch");
    d->model = *model;
    return CHAOS_OK;
}

/* RenderingController.zoomAt (:130-150): doubles throughout; ZOOM_COEFF is the float 0.977 widened */
static void zoom_at(chaos_driver *d, const int where[2], bool into)
{
    double *s = d->model.segment;
    const double segment_width = s[2] - s[0], segment_height = s[3] - s[1];
    const double rel_top = where[1] / (double)d->canvas_h, rel_btm = 1 - rel_top;
    const double rel_left = where[0] / (double)d->canvas_w, rel_rght = 1 - rel_left;
    const double center_x = s[0] + segment_width * rel_left, center_y = s[1] + segment_height * rel_btm;
    const double zoom_coeff_in = (double)0.977f;
    const double zoom_coeff = into ? zoom_coeff_in : 2.0 - zoom_coeff_in;    /* `2f - ZOOM_COEFF` is a double subtraction in Java */
    const double lbx = center_x - segment_width * rel_left * zoom_coeff, lby = center_y - segment_height * rel_btm * zoom_coeff;
    const double rtx = center_x + segment_width * rel_rght * zoom_coeff, rty = center_y + segment_height * rel_top * zoom_coeff;
    s[0] = lbx; s[1] = lby; s[2] = rtx; s[3] = rty;
}

static void set_max_super_sampling(chaos_driver *d, float v)     /* Model.setMaxSuperSampling (Model.java:166-168) clamps */
{
    d->model.max_super_sampling = v < 0.f ? 0.f : (v > (float)CHAOS_MAX_SUPER_SAMPLING ? (float)CHAOS_MAX_SUPER_SAMPLING : v);
}

/* setParamsToBeRenderedIn (:239-245): float arithmetic; a frame time of 0 gives +Inf, which the min() turns into 64 */
static void set_params_to_be_rendered_in(chaos_driver *d, int ms)
{
    float new_ss = d->model.max_super_sampling * (float)ms / d->last_frame_render_time;
    new_ss = fminf(new_ss, (float)CHAOS_MAX_SUPER_SAMPLING);
    if (new_ss != new_ss) new_ss = (float)CHAOS_MAX_SUPER_SAMPLING;      /* 0 * Inf: Math.min would pass the NaN on to the clamp, which keeps 64 */
    set_max_super_sampling(d, new_ss);
}

/* determineRenderingModeQuality (:200-237); false = this frame is not rendered (progressive refinement is over) */
static bool determine_quality(chaos_driver *d)
{
    if (!d->automatic_quality) return true;
    mode_fsm &f = d->fsm;
    if (f.different_than_last()) {          /* "RESET SS" */
        set_max_super_sampling(d, 1.f);
        return true;
    }
    const float prev = d->model.max_super_sampling;
    if (f.is_zooming() || f.is_moving()) {
        set_params_to_be_rendered_in(d, kShortestFrameRenderTime);
    } else if (f.is_progressive()) {
        int desired = (kShortestFrameRenderTime * 2) << f.pr_lvl;         /* exponentially growing frame time */
        const float twice_last = d->last_frame_render_time * 2.f;
        float desired_f = twice_last > (float)desired ? twice_last : (float)desired;
        if (d->clock == CHAOS_CLOCK_WALL_INT) desired_f = floorf(desired_f);
        if (desired_f > (float)kMaxFrameRenderTime || d->model.max_super_sampling >= (float)CHAOS_MAX_SUPER_SAMPLING) {
            if (f.pr_lvl != 0) {            /* level 0 happens upon a parameter change: do not stop then */
                f.reset_state();
                set_max_super_sampling(d, prev);
                return false;
            }
        } else {
            set_params_to_be_rendered_in(d, (int)desired_f);
        }
    }
    return true;
}

/* GLRenderer.display + cudaRender + RenderingController.onRenderingDone: one animator tick.
 * *rendered: 0 nothing (waiting, or the refinement ended), 1 a fast frame, 2 a quality frame */
extern "C" chaos_status chaos_driver_display(chaos_driver *d, int *rendered)
{
    DRIVER_OR_FAIL(d);
    if (rendered) *rendered = 0;
    const double start = wall_ms();
    mode_fsm &f = d->fsm;
    if (f.is_zooming()) zoom_at(d, d->mouse, f.zooming_direction);
    if (f.is_waiting()) return CHAOS_OK;
    if (!determine_quality(d)) return CHAOS_OK;
    d->model.is_zooming = f.is_zooming() ? 1 : 0;
    if (f.is_zooming()) d->model.is_zooming_in = f.zooming_direction ? 1 : 0;
    d->model.mouse_focus[0] = d->mouse[0]; d->model.mouse_focus[1] = d->mouse[1];
    const bool quality_frame = f.is_progressive();
    float device_ms = -1.f;
    chaos_status st;
    if (d->renderer) st = render_with_renderer(d, quality_frame, &device_ms);
    else {
        st = (quality_frame ? d->quality : d->fast)(d->user, &d->model, &device_ms);
        if (st != CHAOS_OK) dfail(st, "the render callback failed");
    }
    if (st != CHAOS_OK) return st;
    if (d->clock == CHAOS_CLOCK_WALL_INT || device_ms < 0.f) {
        const double ms = device_ms >= 0.f && !d->renderer ? (double)device_ms : wall_ms() - start;   /* a callback's time is its clock */
        d->last_frame_render_time = d->clock == CHAOS_CLOCK_WALL_INT ? (float)(int)ms : (float)ms;
    } else {
        d->last_frame_render_time = device_ms;
    }
    d->frames += 1;
    d->last_kind = quality_frame ? 2 : 1;
    if (rendered) *rendered = d->last_kind;
    f.step();                                /* onRenderingDone */
    return CHAOS_OK;
}

extern "C" chaos_status chaos_driver_get_state(const chaos_driver *d, chaos_driver_state *out)
{
    DRIVER_OR_FAIL(d);
    if (!out || out->struct_size != sizeof(chaos_driver_state)) return dfail(CHAOS_ERR_ILLEGAL_ARGUMENT, "chaos_driver_state.struct_size mismatch");
    out->mode = d->fsm.current; out->last_mode = d->fsm.last;
    out->progressive_rendering_level = d->fsm.pr_lvl;
    out->zooming = d->fsm.is_zooming() ? 1 : 0; out->moving = d->fsm.is_moving() ? 1 : 0;
    out->zooming_in = d->fsm.zooming_direction ? 1 : 0; out->last_kind = (uint8_t)d->last_kind;
    out->last_frame_render_time_ms = d->last_frame_render_time;
    out->frames = d->frames;
    out->model = d->model;
    return CHAOS_OK;
}

/* the reference's zoom session: the button is pressed at (x, y) for `ticks` animator ticks, released, and after the
 * timer the picture is refined progressively until the machine is back in Waiting */
extern "C" chaos_status chaos_driver_run_zoom_session(chaos_driver *d, int x, int y, int inside, uint32_t ticks, uint32_t *frames_rendered)
{
    DRIVER_OR_FAIL(d);
    uint32_t n = 0;
    d->mouse[0] = x; d->mouse[1] = y;
    d->fsm.start_zooming(inside != 0, false);
    for (uint32_t t = 0; t < ticks; ++t) {
        int k = 0;
        chaos_status st = chaos_driver_display(d, &k);
        if (st != CHAOS_OK) return st;
        n += k ? 1u : 0u;
    }
    d->fsm.stop_zooming();
    d->fsm.start_progressive();
    while (!d->fsm.is_waiting()) {
        int k = 0;
        chaos_status st = chaos_driver_display(d, &k);
        if (st != CHAOS_OK) return st;
        if (!k) break;
        n += 1u;
    }
    if (frames_rendered) *frames_rendered = n;
    return CHAOS_OK;
}
