"""Frame driver: the reference's rendering-mode state machine and automatic-quality controller around a renderer.

SURVEY.md 8(f)2.  In the reference this logic lives in the GUI layer and is driven by AWT events and the JOGL
animator; here it is a plain object so that a zoom session can be replayed (and benchmarked) with the reference's
closed-loop behaviour -- maxSuperSampling retargeted every frame so that a frame takes 15 ms while zooming/moving,
and 30, 60, ... ms while refining progressively -- instead of a fixed sample budget.

Mirrors, method for method:
  RenderingModeFSM        rendering/RenderingModeFSM.java:9-155
  FrameDriver.display     rendering/GLRenderer.java:113-162 (display), :167-185 (cudaRender),
                          :200-237 (determineRenderingModeQuality), :239-245 (setParamsToBeRenderedIn)
  FrameDriver.zoom_at     rendering/RenderingController.java:130-150
  on_rendering_done       rendering/RenderingController.java:264-269
`renderer` is anything with renderFast(model) / renderQuality(model): a CudaFractalRenderer, or a stub in tests.
`clock` returns milliseconds (System.currentTimeMillis in the reference).
"""
from __future__ import annotations

import time
from typing import Callable, List, Optional, Sequence

MAX_SUPER_SAMPLING = 64
WAITING, ZOOMING_AUTO, ZOOMING_ONCE, MOVING, PROGRESSIVE = "Waiting", "ZoomingAuto", "ZoomingOnce", "Moving", "ProgressiveRendering"


class RenderingModeFSM:
    MAX_PROGRESSIVE_RENDERING_LEVEL = 6

    def __init__(self):
        self.current = WAITING
        self.last = WAITING
        self.pr_lvl = 0
        self.zooming_and_moving = False
        self.zooming_direction = False

    def resetState(self):
        self.last, self.current, self.zooming_and_moving = self.current, WAITING, False

    def step(self):
        if (self.current == WAITING and self.last in (ZOOMING_AUTO, MOVING)) or self.current == ZOOMING_ONCE:
            new, self.pr_lvl = PROGRESSIVE, -1
        elif self.current == PROGRESSIVE and self.pr_lvl >= self.MAX_PROGRESSIVE_RENDERING_LEVEL:
            new = WAITING
        else:
            new = self.current
        self.last, self.current = self.current, new
        if self.current == PROGRESSIVE:
            self.pr_lvl = min(self.MAX_PROGRESSIVE_RENDERING_LEVEL, self.pr_lvl + 1)

    def doZoomingManualOnce(self, inside: bool):
        self.last, self.current, self.zooming_direction, self.zooming_and_moving = self.current, ZOOMING_ONCE, inside, False

    def startZooming(self, inside: bool):
        self.last, self.current, self.zooming_direction, self.zooming_and_moving = self.current, ZOOMING_AUTO, inside, False

    def startZoomingAndMoving(self, inside: bool):
        self.startZooming(inside)
        self.zooming_and_moving = True

    def stopZooming(self):
        self.last = self.current
        self.current = MOVING if self.zooming_and_moving else WAITING
        self.zooming_and_moving = False

    def startMoving(self):
        self.last, self.current = self.current, MOVING

    def stopMoving(self):
        self.last = self.current
        if not self.zooming_and_moving:
            self.current = WAITING
        self.zooming_and_moving = False

    def startProgressiveRendering(self):
        self.last, self.current, self.pr_lvl, self.zooming_and_moving = self.current, PROGRESSIVE, 0, False

    def isZooming(self) -> bool:
        return self.current in (ZOOMING_AUTO, ZOOMING_ONCE) or self.zooming_and_moving

    def getZoomingDirection(self) -> bool:
        if not self.isZooming():
            raise RuntimeError("cannot ask for zooming direction when not zooming")
        return self.zooming_direction

    def isMoving(self) -> bool:
        return self.current == MOVING or self.zooming_and_moving

    def isProgressiveRendering(self) -> bool:
        return self.current == PROGRESSIVE

    def getProgressiveRenderingLevel(self) -> int:
        if not self.isProgressiveRendering():
            raise RuntimeError("cannot ask for Progressive rendering level when not Progressive rendering")
        return self.pr_lvl

    def isWaiting(self) -> bool:
        return self.current == WAITING

    def isDifferentThanLast(self) -> bool:
        return self.current != self.last


class FrameDriver:
    SHORTEST_FRAME_RENDER_TIME = 15    # ms, GLRenderer.java:190
    MAX_FRAME_RENDER_TIME = 1000       # ms, GLRenderer.java:194

    def __init__(self, renderer, model, clock: Optional[Callable[[], float]] = None, automatic_quality: bool = True):
        self.renderer = renderer
        self.model = model
        self.state = RenderingModeFSM()
        self.clock = clock or (lambda: time.perf_counter() * 1e3)
        self.automatic_quality = automatic_quality
        self.last_frame_render_time = self.SHORTEST_FRAME_RENDER_TIME
        self.last_mouse_position: Sequence[int] = (0, 0)
        self.log: List[tuple] = []     # (mode, kind, maxSuperSampling, frame_ms) per rendered frame

    # --- RenderingController.zoomAt --------------------------------------------------------------------------
    def zoom_at(self, where: Sequence[int], into: bool):
        self.model.zoomAt(where, into)

    # --- GLRenderer.setParamsToBeRenderedIn -----------------------------------------------------------------
    def _set_params_to_be_rendered_in(self, ms: int):
        new_ss = self.model.maxSuperSampling * ms / float(self.last_frame_render_time)
        self.model.setMaxSuperSampling(min(new_ss, MAX_SUPER_SAMPLING))

    # --- GLRenderer.determineRenderingModeQuality -----------------------------------------------------------
    def _determine_quality(self) -> bool:
        if not self.automatic_quality:
            return True
        st = self.state
        if st.isDifferentThanLast():
            self.model.setMaxSuperSampling(1)
            return True
        prev = self.model.maxSuperSampling
        if st.isZooming() or st.isMoving():
            self._set_params_to_be_rendered_in(self.SHORTEST_FRAME_RENDER_TIME)
        elif st.isProgressiveRendering():
            desired = self.SHORTEST_FRAME_RENDER_TIME * 2 << st.getProgressiveRenderingLevel()
            desired = max(self.last_frame_render_time * 2, desired)
            if desired > self.MAX_FRAME_RENDER_TIME or self.model.maxSuperSampling >= MAX_SUPER_SAMPLING:
                if st.getProgressiveRenderingLevel() != 0:
                    st.resetState()
                    self.model.setMaxSuperSampling(prev)
                    return False
            else:
                self._set_params_to_be_rendered_in(desired)
        return True

    # --- GLRenderer.display + cudaRender + RenderingController.onRenderingDone ---------------------------------
    def display(self) -> bool:
        """one frame; returns True if something was rendered"""
        start = self.clock()
        st = self.state
        if st.isZooming():
            self.zoom_at(self.last_mouse_position, st.getZoomingDirection())
        if st.isWaiting():
            return False
        if not self._determine_quality():
            return False
        self.model.zooming = st.isZooming()
        if st.isZooming():
            self.model.zoomingIn = st.getZoomingDirection()
        self.model.mouseFocus = tuple(self.last_mouse_position)
        mode = st.current
        if st.isProgressiveRendering():
            self.renderer.renderQuality(self.model)
            kind = "quality"
        else:
            self.renderer.renderFast(self.model)
            kind = "fast"
        self.last_frame_render_time = max(1, int(self.clock() - start))   # the reference divides by it; never let it be 0
        self.log.append((mode, kind, self.model.maxSuperSampling, self.last_frame_render_time))
        st.step()                                                          # onRenderingDone
        return True

    # --- sessions ---------------------------------------------------------------------------------------------
    def run_zoom_session(self, where: Sequence[int], into: bool, frames: int) -> int:
        """mouse pressed at `where` for `frames` animator ticks, then released; progressive refinement until the FSM
        goes back to Waiting.  Returns the number of frames rendered."""
        self.last_mouse_position = tuple(where)
        self.state.startZooming(into)
        n = 0
        for _ in range(frames):
            n += int(self.display())
        self.state.stopZooming()
        self.state.step()                # the timer-fired repaint after release (RenderingController.java:103-106)
        while not self.state.isWaiting():
            if not self.display():
                break
            n += 1
        return n
