"""Frame driver: ctypes face of the native driver in libchaos_ultra.so (csrc/chaos_driver.cpp).

SURVEY.md 8(f)2.  The reference's rendering-mode state machine (rendering/RenderingModeFSM.java:9-155) and automatic
quality controller (rendering/GLRenderer.java:113-245) live in its GUI layer; here they sit behind the C ABI
(``chaos_driver_*`` in include/chaos_ultra.h), so the Java host, a benchmark or a test can replay a zoom session with the
reference's closed-loop behaviour -- maxSuperSampling retargeted every frame so that a frame takes 15 ms while zooming or
moving and 30, 60, ... ms while refining progressively.  This module only marshals: no logic of the controller is in Python.
"""
from __future__ import annotations

import ctypes as C
import importlib
from typing import Callable, List, Optional, Sequence

_pkg = importlib.import_module(__name__.rsplit(".", 1)[0])
_Params, _check_status = _pkg._Params, _pkg._STATUS_TO_EXC

WAITING, ZOOMING_AUTO, ZOOMING_ONCE, MOVING, PROGRESSIVE = 0, 1, 2, 3, 4
MODE_NAMES = ["Waiting", "ZoomingAuto", "ZoomingOnce", "Moving", "ProgressiveRendering"]
CLOCK_DEVICE, CLOCK_WALL_INT = 0, 1
MAX_PROGRESSIVE_RENDERING_LEVEL = 6
_VP = C.c_void_p


class _State(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("mode", C.c_int), ("last_mode", C.c_int), ("progressive_rendering_level", C.c_int32),
                ("zooming", C.c_uint8), ("moving", C.c_uint8), ("zooming_in", C.c_uint8), ("last_kind", C.c_uint8),
                ("last_frame_render_time_ms", C.c_float), ("frames", C.c_uint64), ("model", _Params)]


RENDER_FN = C.CFUNCTYPE(C.c_int, _VP, C.POINTER(_Params), C.POINTER(C.c_float))
_DRIVER_API = {
    "chaos_driver_create": (C.c_int, [_VP, C.POINTER(_Params), C.POINTER(_VP)]),
    "chaos_driver_create_custom": (C.c_int, [RENDER_FN, RENDER_FN, _VP, C.POINTER(_Params), C.c_uint32, C.c_uint32, C.POINTER(_VP)]),
    "chaos_driver_destroy": (C.c_int, [_VP]),
    "chaos_driver_mouse": (C.c_int, [_VP, C.c_int, C.c_int]),
    "chaos_driver_start_zooming": (C.c_int, [_VP, C.c_int, C.c_int]),
    "chaos_driver_zoom_once": (C.c_int, [_VP, C.c_int]),
    "chaos_driver_stop_zooming": (C.c_int, [_VP]),
    "chaos_driver_start_moving": (C.c_int, [_VP]),
    "chaos_driver_stop_moving": (C.c_int, [_VP]),
    "chaos_driver_start_progressive_rendering": (C.c_int, [_VP, C.c_int]),
    "chaos_driver_step": (C.c_int, [_VP]),
    "chaos_driver_set_automatic_quality": (C.c_int, [_VP, C.c_int]),
    "chaos_driver_set_clock": (C.c_int, [_VP, C.c_int]),
    "chaos_driver_set_model": (C.c_int, [_VP, C.POINTER(_Params)]),
    "chaos_driver_display": (C.c_int, [_VP, C.POINTER(C.c_int)]),
    "chaos_driver_get_state": (C.c_int, [_VP, C.POINTER(_State)]),
    "chaos_driver_run_zoom_session": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_uint32)]),
    "chaos_driver_last_error": (C.c_char_p, []),
}


def _lib():
    lib = _pkg.load_library()
    if not getattr(lib, "_chaos_driver_bound", False):
        for name, (res, args) in _DRIVER_API.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        lib._chaos_driver_bound = True
    return lib


class FrameDriver:
    """``renderer``: a CudaFractalRenderer (frames are rendered by chaos_render_fast / chaos_render_quality and timed on the
    device), or any object with renderFast(model) / renderQuality(model) returning the frame's milliseconds (a stub: the
    controller then runs without a GPU).  ``model`` is a RenderingModel; it is kept up to date after every call."""

    def __init__(self, renderer, model, clock: int = CLOCK_DEVICE, automatic_quality: bool = True):
        self._l = _lib()
        self.renderer, self.model = renderer, model
        self.log: List[tuple] = []     # (mode name, kind, maxSuperSampling, frame ms) per rendered frame
        self._h = _VP()
        p = model._to_c()
        if hasattr(renderer, "_h") and hasattr(renderer, "_lib"):
            self._ok(self._l.chaos_driver_create(renderer._h, C.byref(p), C.byref(self._h)))
            self._cbs = None
        else:
            def call(kind):
                def cb(_user, params, frame_ms):
                    try:
                        m = self.model
                        m.planeSegment = list(params.contents.segment)
                        m.maxSuperSampling = float(params.contents.max_super_sampling)
                        m.zooming, m.zoomingIn = bool(params.contents.is_zooming), bool(params.contents.is_zooming_in)
                        m.mouseFocus = tuple(params.contents.mouse_focus)
                        ms = getattr(renderer, kind)(m)
                        frame_ms[0] = float(ms) if ms is not None else -1.0
                        return 0
                    except Exception:
                        return 5
                return RENDER_FN(cb)
            self._cbs = (call("renderFast"), call("renderQuality"))
            self._ok(self._l.chaos_driver_create_custom(self._cbs[0], self._cbs[1], None, C.byref(p), model.canvasWidth, model.canvasHeight,
                                                         C.byref(self._h)))
        self._ok(self._l.chaos_driver_set_clock(self._h, clock))
        self._ok(self._l.chaos_driver_set_automatic_quality(self._h, int(automatic_quality)))

    def _ok(self, st):
        if st != 0:
            msg = (self._l.chaos_driver_last_error() or b"").decode("utf-8", "replace")
            raise _check_status.get(st, _pkg.ChaosError)(msg)

    def state(self) -> _State:
        s = _State()
        s.struct_size = C.sizeof(_State)
        self._ok(self._l.chaos_driver_get_state(self._h, C.byref(s)))
        return s

    def _sync_model(self, s: Optional[_State] = None):
        s = s or self.state()
        m = self.model
        m.planeSegment = list(s.model.segment)
        m.maxSuperSampling = float(s.model.max_super_sampling)
        m.zooming, m.zoomingIn = bool(s.model.is_zooming), bool(s.model.is_zooming_in)
        m.mouseFocus = tuple(s.model.mouse_focus)
        m.sampleReuseCacheDirty = bool(s.model.sample_reuse_cache_dirty)
        m.floatingPointPrecision = int(s.model.float_precision)
        return s

    # RenderingModeFSM / RenderingController's mouse handlers
    def mouse(self, x: int, y: int): self._ok(self._l.chaos_driver_mouse(self._h, int(x), int(y)))
    def startZooming(self, inside: bool): self._ok(self._l.chaos_driver_start_zooming(self._h, int(inside), 0))
    def startZoomingAndMoving(self, inside: bool): self._ok(self._l.chaos_driver_start_zooming(self._h, int(inside), 1))
    def doZoomingManualOnce(self, inside: bool): self._ok(self._l.chaos_driver_zoom_once(self._h, int(inside)))
    def stopZooming(self): self._ok(self._l.chaos_driver_stop_zooming(self._h))
    def startMoving(self): self._ok(self._l.chaos_driver_start_moving(self._h))
    def stopMoving(self): self._ok(self._l.chaos_driver_stop_moving(self._h))
    def startProgressiveRendering(self, reset_first: bool = False): self._ok(self._l.chaos_driver_start_progressive_rendering(self._h, int(reset_first)))
    def step(self): self._ok(self._l.chaos_driver_step(self._h))

    def isWaiting(self) -> bool: return self.state().mode == WAITING
    def isZooming(self) -> bool: return bool(self.state().zooming)
    def isMoving(self) -> bool: return bool(self.state().moving)
    def isProgressiveRendering(self) -> bool: return self.state().mode == PROGRESSIVE
    def getProgressiveRenderingLevel(self) -> int: return int(self.state().progressive_rendering_level)
    def isDifferentThanLast(self) -> bool:
        s = self.state()
        return s.mode != s.last_mode

    def display(self) -> int:
        """one animator tick (GLRenderer.display); 0 = nothing rendered, 1 = a fast frame, 2 = a quality frame"""
        before = self.state().mode
        k = C.c_int(0)
        self._ok(self._l.chaos_driver_display(self._h, C.byref(k)))
        s = self._sync_model()
        if k.value:
            self.log.append((MODE_NAMES[before], "quality" if k.value == 2 else "fast", float(s.model.max_super_sampling), float(s.last_frame_render_time_ms)))
        return k.value

    def run_zoom_session(self, where: Sequence[int], into: bool, frames: int) -> int:
        """the button pressed at `where` for `frames` animator ticks, released, progressive refinement until Waiting; tick by
        tick so that every frame is logged (chaos_driver_run_zoom_session does the same in one native call)"""
        self.mouse(where[0], where[1])
        self.startZooming(into)
        n = 0
        for _ in range(frames):
            n += 1 if self.display() else 0
        self.stopZooming()
        self.startProgressiveRendering()
        while not self.isWaiting():
            if not self.display():
                break
            n += 1
        return n

    def run_zoom_session_native(self, where: Sequence[int], into: bool, frames: int) -> int:
        n = C.c_uint32(0)
        self._ok(self._l.chaos_driver_run_zoom_session(self._h, int(where[0]), int(where[1]), int(into), int(frames), C.byref(n)))
        self._sync_model()
        return int(n.value)

    def close(self):
        if self._h:
            self._l.chaos_driver_destroy(self._h)
            self._h = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
