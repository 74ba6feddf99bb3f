"""Shared parity cases: the same seeded/deterministic inputs are used by

* tests/golden/make_golden.py  -- runs the REFERENCE kernels (oracle/_ref) on a B200 and writes fixtures,
* tests/test_oracle_golden.py  -- CPU: oracle == fixtures,
* tests/test_parity_gpu.py     -- GPU: this backend (through the C ABI) == oracle == fixtures == live reference.

Viewports follow SURVEY.md 8(d): Model.setPlaneSegmentFromCenter (rendering/Model.java:247-256) applied to the
reference's own defaults and premade views (modules/ModuleMandelbrot.java:22, gui/PresenterFX.java:245-302).
Sizes are small enough for the scalar oracle to finish in seconds; ragged sizes (not multiples of the
8x4 vote tile) and 1-pixel frames are the edge cases of this domain.
"""
from __future__ import annotations

A, FOV, REUSE, ZOOMING, ZOOM_IN = 1, 4, 8, 16, 32


def seg(cx, cy, zoom, W, H):
    relW = 1.0 / float(H) * W
    return [cx - relW * zoom / 2, cy - 1.0 * zoom / 2, cx + relW * zoom / 2, cy + 1.0 * zoom / 2]


def _c(name, fractal, W, H, center, zoom, maxIter, maxSS, flags, double, **kw):
    d = dict(name=name, fractal=fractal, W=W, H=H, image=seg(center[0], center[1], zoom, W, H), maxIter=maxIter,
             maxSS=float(maxSS), flags=flags, double=double, julia_c=(0.0, 0.0), amplifier=10)
    d.update(kw)
    return d


MAIN_CASES = [
    # config c1 shape (1 sample, FP64, full set) at oracle-friendly size
    _c("m_full_1s_f64", "mandelbrot", 128, 96, (-0.5, 0.0), 2.0, 500, 1, 0, True),
    _c("m_full_1s_f32", "mandelbrot", 128, 96, (-0.5, 0.0), 2.0, 500, 1, 0, False),
    # config c2 shape: adaptive supersampling, ragged size
    _c("m_full_a8_f64", "mandelbrot", 203, 117, (-0.5, 0.0), 2.0, 300, 8, A, True),
    _c("m_full_a8_f32", "mandelbrot", 203, 117, (-0.5, 0.0), 2.0, 300, 8, A, False),
    # "M ex 2" (PresenterFX.java:255-262): the reference's own rule picks FP64 here at 4K
    _c("m_ex2_a5_f64", "mandelbrot", 160, 100, (-0.235125, 0.827215), 4.0e-5, 1600, 5, A, True),
    # "M ex 1" centre, mid zoom, fractional sample budget, FP32
    _c("m_ex1_a3p6_f32", "mandelbrot", 152, 88, (-0.748, 0.1), 0.0014, 800, 3.6, A, False),
    # non-adaptive: only the i == S/2 clause can shorten the loop
    _c("m_full_n4_f64", "mandelbrot", 96, 64, (-0.5, 0.0), 2.0, 200, 4, 0, True),
    _c("m_full_n7_f32", "mandelbrot", 96, 64, (-0.5, 0.0), 2.0, 200, 7, 0, False),
    # largest budget whose decision block never reads past samples[10] (S/2 <= 9)
    _c("m_edge_a19_f64", "mandelbrot", 64, 48, (-0.75, 0.1), 0.05, 400, 19, A, True),
    # "M ex 5" deep zoom (PresenterFX.java:285-292), FP64 only
    _c("m_ex5_1s_f64", "mandelbrot", 96, 96, (-0.551042868375875, 0.62714332109057), 8.00592947491907e-9, 3000, 1, 0, True),
    _c("m_ex5_a4_f64", "mandelbrot", 64, 64, (-0.551042868375875, 0.62714332109057), 8.00592947491907e-9, 3000, 4, A, True),
    # julia module (ModuleJulia.java:40-43 defaults; "Jul ex 1" PresenterFX.java:295-302)
    _c("j_def_a2_f64", "julia", 160, 96, (0.0, 0.0), 4.0, 900, 2, A, True, julia_c=(-0.4, 0.6)),
    _c("j_def_a2_f32", "julia", 160, 96, (0.0, 0.0), 4.0, 900, 2, A, False, julia_c=(-0.4, 0.6)),
    _c("j_ex1_a6_f64", "julia", 120, 72, (0.8327291525311472, -0.10212349314316674), 0.017509763680325807, 900, 6, A, True,
       julia_c=(-0.8, 0.156)),
    # test module: closed form, checks the pixel->plane mapping alone
    _c("t_amp10_f64", "test", 100, 60, (0.0, 0.0), 4.0, 10, 1, 0, True, amplifier=10),
    _c("t_amp3_a5_f32", "test", 100, 60, (0.3, -0.2), 2.5, 10, 5, A, False, amplifier=3),
    # degenerate sizes
    _c("m_1x1_f64", "mandelbrot", 1, 1, (-0.5, 0.0), 2.0, 100, 4, A, True),
    _c("m_7x3_f32", "mandelbrot", 7, 3, (-0.5, 0.0), 2.0, 100, 4, A, False),
    _c("m_9x5_f64", "mandelbrot", 9, 5, (-0.5, 0.0), 2.0, 100, 8, A, True),
    _c("m_maxiter1_f64", "mandelbrot", 40, 24, (-0.5, 0.0), 2.0, 1, 3, A, True),
]

# Modules without a CPU oracle in this round (thrust::complex arithmetic): checked against the reference's own
# kernels run live and against fixtures produced by them.  `params` is the custom-parameter text the GUI would send
# (PresenterFX.java:335-364, ModuleNewtonGeneric.java:69-73).
N_DEFAULT = '{ "coefficients" : [1, 0, 0, -1], "roots" : [ [1,0], [-0.5,0.86602540378] , [-0.5,-0.86602540378] ] }'
N3 = ('{ "coefficients" : [1, 0, -2, 2], "roots" : [ [-1.7692923542386314,0], [0.884646177119315707620204,0.589742805022205501647280] , '
      '[0.884646177119315707,-0.589742805022205501] ] }')
N_ITER = '{"colorMagnifier": 11,' + N_DEFAULT[1:]
EXTRA_MAIN_CASES = [
    _c("nw_a2_f32", "newton_wired", 160, 96, (0.0, 0.0), 4.0, 200, 2, A, False),
    _c("nw_a4_f64", "newton_wired", 120, 72, (0.1, -0.2), 1.5, 100, 4, A, True),
    _c("nw_n1_deep_f64", "newton_wired", 96, 64, (1.8252568181102808e-4, -1.0538321727858829e-4), 6.94260652234243e-7, 100, 1, 0, True),
    _c("ng_def_a2_f32", "newton_generic", 160, 96, (0.0, 0.0), 4.0, 200, 2, A, False, params=N_DEFAULT),
    _c("ng_n3_a3_f64", "newton_generic", 120, 72, (0.0, 0.0), 4.0, 100, 3, A, True, params=N3),
    _c("ni_def_a2_f32", "newton_iterations", 160, 96, (0.0, 0.0), 4.0, 200, 2, A, False, params=N_ITER),
    _c("ni_n3_a5_f64", "newton_iterations", 120, 72, (0.3, 0.1), 2.0, 70, 5, A, True, params='{"colorMagnifier": 7,' + N3[1:]),
    _c("goc_a2_f32", "goc", 96, 64, (1.1, -0.2), 0.20000000000000004, 40, 2, A, False),
    _c("goc_1s_f64", "goc", 96, 64, (1.1, -0.2), 0.20000000000000004, 30, 1, 0, True),
]
DISPLAY_NAME = {"mandelbrot": "mandelbrot", "julia": "julia", "test": "test", "newton_wired": "newton wired",
                "newton_generic": "newton generic", "newton_iterations": "newton colored by iterations", "goc": "goc"}


def newton_constants(params_text):
    """what ModuleNewtonGeneric/Iterations write to the device for a parameter text: (roots[6], coefficients[4], magnifier|None)"""
    import json
    j = json.loads(params_text)
    roots = [float(v) for r in j["roots"] for v in r]
    coefs = [float(v) for v in reversed(j["coefficients"])]
    return roots, coefs, (int(j["colorMagnifier"]) if "colorMagnifier" in j else None)


# Advanced (fast-frame) cases: frame 0 is a quality render of `image0`; frame 1 is the advanced kernel on the
# segment after `zooms` applications of zoomAt(focus, into) (RenderingController.java:130-150).
def _a(name, fractal, W, H, center, zoom, maxIter, maxSS, flags, double, focus, into=True, zooms=1, **kw):
    d = _c(name, fractal, W, H, center, zoom, maxIter, maxSS, flags, double, **kw)
    d.update(focus=focus, into=into, zooms=zooms, maxSS0=max(1.0, float(maxSS)))  # frame 0 is a quality render: maxSS >= 1
    return d


ADV_CASES = [
    _a("adv_reuse_only_f32", "mandelbrot", 160, 96, (-0.748, 0.1), 2.0, 400, 2, A | REUSE, False, (80, 48)),
    _a("adv_zoomin_fov_f32", "mandelbrot", 160, 96, (-0.748, 0.1), 2.0, 400, 2, A | FOV | REUSE | ZOOMING | ZOOM_IN, False, (80, 48)),
    _a("adv_zoomin_fov_f64", "mandelbrot", 160, 96, (-0.748, 0.1), 2.0, 400, 2, A | FOV | REUSE | ZOOMING | ZOOM_IN, True, (80, 48)),
    _a("adv_zoomin_fov_offc_f32", "mandelbrot", 203, 117, (-0.748, 0.1), 1.0, 300, 4, A | FOV | REUSE | ZOOMING | ZOOM_IN, False, (31, 90)),
    _a("adv_zoomout_f64", "mandelbrot", 160, 96, (-0.748, 0.1), 0.5, 400, 3, A | FOV | REUSE | ZOOMING, True, (100, 30), into=False),
    _a("adv_fov_noreuse_f32", "mandelbrot", 160, 96, (-0.748, 0.1), 2.0, 300, 6, A | FOV | ZOOMING | ZOOM_IN, False, (80, 48)),
    _a("adv_lowss_f32", "mandelbrot", 160, 96, (-0.748, 0.1), 2.0, 300, 0.6, A | FOV | REUSE | ZOOMING | ZOOM_IN, False, (80, 48)),
    _a("adv_deep_f64", "mandelbrot", 128, 80, (-0.235125, 0.827215), 4.0e-5, 800, 2, A | FOV | REUSE | ZOOMING | ZOOM_IN, True, (64, 40), zooms=3),
    _a("adv_julia_f32", "julia", 160, 96, (0.0, 0.0), 4.0, 300, 2, A | FOV | REUSE | ZOOMING | ZOOM_IN, False, (80, 48), julia_c=(-0.4, 0.6)),
    _a("adv_test_f32", "test", 96, 64, (0.1, -0.1), 3.0, 10, 3, A | FOV | REUSE | ZOOMING | ZOOM_IN, False, (40, 30), amplifier=7),
    _a("adv_newton_f32", "newton_generic", 120, 72, (0.0, 0.0), 3.0, 100, 3, A | FOV | REUSE | ZOOMING | ZOOM_IN, False, (60, 36), params=N3),
    _a("adv_newton_iter_f64", "newton_iterations", 96, 64, (0.2, 0.1), 2.0, 80, 2, A | FOV | REUSE | ZOOMING | ZOOM_IN, True, (30, 40), params=N_ITER),
    _a("adv_ragged_f64", "mandelbrot", 61, 35, (-0.5, 0.0), 2.0, 200, 3, A | FOV | REUSE | ZOOMING | ZOOM_IN, True, (30, 17)),
]


def zoom_at(image, W, H, where, into):
    """RenderingController.zoomAt in plain Python doubles (ZOOM_COEFF = (double)0.977f)."""
    import struct
    zc32 = struct.unpack("f", struct.pack("f", 0.977))[0]
    two_minus = struct.unpack("f", struct.pack("f", 2.0 - zc32))[0]
    lbx, lby, rtx, rty = image
    sw, sh = rtx - lbx, rty - lby
    relTop = where[1] / float(H)
    relBtm = 1 - relTop
    relLeft = where[0] / float(W)
    relRght = 1 - relLeft
    cx = lbx + sw * relLeft
    cy = lby + sh * relBtm
    zc = zc32 if into else two_minus
    return [cx - sw * relLeft * zc, cy - sh * relBtm * zc, cx + sw * relRght * zc, cy + sh * relTop * zc]


def adv_segments(case):
    img0 = list(case["image"])
    img1 = img0
    for _ in range(case["zooms"]):
        img1 = zoom_at(img1, case["W"], case["H"], case["focus"], case["into"])
    return img0, img1
