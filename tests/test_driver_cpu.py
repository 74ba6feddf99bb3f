"""CPU: the frame driver (rendering-mode FSM + automatic quality) against the reference's rules, with a stub renderer
whose frame time is a known function of the sample budget (the role FractalRendererNullObjectVerbose plays there)."""
import importlib
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
cu = importlib.import_module("chaos-ultra_b200")
drv = importlib.import_module("chaos-ultra_b200.driver")


class StubRenderer:
    """a frame takes 1 ms + ms_per_sample x maxSuperSampling: the controller is driven by what the render call reports"""

    def __init__(self, ms_per_sample):
        self.k, self.calls = ms_per_sample, []

    def renderFast(self, m):
        self.calls.append(("fast", m.maxSuperSampling, m.zooming, m.zoomingIn, list(m.planeSegment)))
        return 1 + self.k * m.maxSuperSampling

    def renderQuality(self, m):
        self.calls.append(("quality", m.maxSuperSampling, m.zooming, m.zoomingIn, list(m.planeSegment)))
        return 1 + self.k * m.maxSuperSampling


def _driver(ms_per_sample=3.0, **kw):
    m = cu.RenderingModel(canvasWidth=320, canvasHeight=180)
    m.resetRenderingValuesToDefault()
    r = StubRenderer(ms_per_sample)
    return drv.FrameDriver(r, m, **kw), r, m


def test_fsm_transitions_match_the_reference():
    f, _, _ = _driver()                                      # rendering/RenderingModeFSM.java:9-155, driven through the C ABI
    assert f.isWaiting() and not f.isZooming()
    f.startZooming(True)
    assert f.isZooming() and f.state().zooming_in and f.isDifferentThanLast()
    f.step()
    assert f.state().mode == drv.ZOOMING_AUTO and not f.isDifferentThanLast()
    f.stopZooming()
    assert f.isWaiting() and f.state().last_mode == drv.ZOOMING_AUTO
    f.step()                                   # Waiting after ZoomingAuto -> progressive rendering, level 0
    assert f.isProgressiveRendering() and f.getProgressiveRenderingLevel() == 0
    for lvl in range(1, 7):
        f.step()
        assert f.getProgressiveRenderingLevel() == lvl
    f.step()                                   # level 6 reached -> Waiting
    assert f.isWaiting()
    f.doZoomingManualOnce(False)
    assert f.isZooming() and not f.state().zooming_in
    f.step()
    assert f.isProgressiveRendering()
    f.startZoomingAndMoving(True)
    assert f.isZooming() and f.isMoving()
    f.stopZooming()
    assert f.state().mode == drv.MOVING and not f.isZooming()
    f.stopMoving()
    assert f.isWaiting()
    assert f.display() == 0                    # Waiting: a tick renders nothing (GLRenderer.java:139-141)


def test_automatic_quality_converges_to_the_frame_time_target():
    d, r, m = _driver(3.0)
    n = d.run_zoom_session((160, 90), True, frames=12)
    fast = [c for c in r.calls if c[0] == "fast"]
    assert len(fast) == 12 and all(c[2] and c[3] for c in fast)
    assert fast[0][1] == 1.0                                  # mode changed -> "RESET SS" (GLRenderer.java:208-212)
    # closed loop: SS * 15 / lastFrameTime; with t = 1 + 3 SS the fixed point is SS = 14/3
    assert abs(fast[-1][1] - 14.0 / 3.0) < 0.35
    assert 13 <= d.log[11][3] <= 16
    # every zooming frame moved the segment by ZOOM_COEFF about the mouse position (RenderingController.java:130-150)
    h0 = fast[0][4][3] - fast[0][4][1]
    h1 = fast[1][4][3] - fast[1][4][1]
    assert abs(h1 / h0 - float(__import__("numpy").float32(0.977))) < 1e-12
    # after release: progressive refinement with quality frames whose budget grows until a frame would exceed 1 s or SS hits 64
    quality = [c for c in r.calls if c[0] == "quality"]
    assert quality and not any(c[2] for c in quality)
    ss = [c[1] for c in quality]
    assert ss[0] == 1.0 and all(b >= a for a, b in zip(ss[1:], ss[2:])) and max(ss) <= 64.0
    assert d.isWaiting() and n == len(r.calls)
    assert m.maxSuperSampling == d.state().model.max_super_sampling        # the Python model follows the native one


def test_the_reference_clock_jumps_to_the_full_budget_on_a_fast_device():
    """GLRenderer.java:145,239-241: an int millisecond count of 0 makes the quotient +Inf and min(., 64) = 64 at once; the
    device clock (float ms) converges instead.  Both are the native controller; only its clock differs."""
    d, r, _ = _driver(0.01, clock=drv.CLOCK_WALL_INT)         # every frame takes well under a millisecond -> (int) 0 ... 1
    d.mouse(160, 90); d.startZooming(True)
    for _ in range(4):
        d.display()
    assert [c[1] for c in r.calls][:3] == [1.0, 15.0, 64.0]   # RESET, then 1 * 15 / (int)1.01, then 15 * 15 / (int)1.15 -> capped
    d2, r2, _ = _driver(0.01)
    d2.mouse(160, 90); d2.startZooming(True)
    for _ in range(6):
        d2.display()
    assert r2.calls[1][1] == 15.0 / 1.01 or abs(r2.calls[1][1] - 15.0 / 1.01) < 1e-4
    assert all(c[1] <= 64.0 for c in r2.calls)


def test_automatic_quality_off_keeps_the_budget():
    d, r, m = _driver(1.0, automatic_quality=False)
    m.setMaxSuperSampling(7)
    d._ok(d._l.chaos_driver_set_model(d._h, __import__("ctypes").byref(m._to_c())))
    d.run_zoom_session((10, 10), False, frames=3)
    assert all(c[1] == 7.0 for c in r.calls)
    assert [c[0] for c in r.calls[:3]] == ["fast"] * 3 and not any(c[3] for c in r.calls[:3])   # zooming out
    # zooming out multiplies the segment by 2 - (double) 0.977f, a double subtraction in Java
    h0, h1 = r.calls[0][4][3] - r.calls[0][4][1], r.calls[1][4][3] - r.calls[1][4][1]
    assert abs(h1 / h0 - (2.0 - float(__import__("numpy").float32(0.977)))) < 1e-12


def test_native_session_equals_the_tick_by_tick_one():
    a, ra, _ = _driver(2.0)
    b, rb, _ = _driver(2.0)
    na = a.run_zoom_session((100, 50), True, frames=9)
    nb = b.run_zoom_session_native((100, 50), True, frames=9)
    assert na == nb and ra.calls == rb.calls


def test_png_export_round_trip(tmp_path):
    import numpy as np
    io = importlib.import_module("chaos-ultra_b200.imageio")
    pal = cu.createDefaultColorPalette()
    frame = pal[(np.arange(48 * 64).reshape(48, 64) * 7) % pal.size].astype(np.uint32)
    p = tmp_path / "frame.png"
    io.save_png(p, frame)
    data = p.read_bytes()
    assert (io.decode_png_rgba8(data) == frame).all()
    try:
        from PIL import Image
    except ImportError:
        return
    img = np.asarray(Image.open(p).convert("RGBA")).astype(np.uint32)     # an independent decoder agrees on the channel order
    back = img[..., 0] | (img[..., 1] << 8) | (img[..., 2] << 16) | (img[..., 3] << 24)
    assert (back == frame).all()
    # reference palettes are PNGs read the other way round (ImageHelpers.loadColorPaletteFromFile)
    io.save_png(tmp_path / "pal.png", pal.reshape(1, -1))
    assert (cu.loadColorPaletteFromFile(tmp_path / "pal.png") == pal).all()
