"""chaos-ultra B200 render backend -- Python mirror of the reference's renderer plugin API.

The reference's host is Java; no JDK exists in the build image, so the host-side mirror used by
the tests and the benchmark is this thin ``ctypes`` layer over ``libchaos_ultra.so``
(``include/chaos_ultra.h``).  Class and method names follow the Java interfaces they stand for:

* ``CudaFractalRendererProvider``  -- cudarenderer/CudaFractalRendererProvider.java:14-91
* ``CudaFractalRenderer``          -- cudarenderer/CudaFractalRenderer.java:32-430
  (``rendering/FractalRenderer.java:14-77``)
* ``RenderingModel``               -- rendering/model/RenderingModel.java:6-20 + rendering/Model.java
* exceptions                       -- the Java exception each ``chaos_status`` stands for

There is no CPU fallback: if the shared library or the module cubins are missing, importing
works but creating a provider raises, and on a GPU box the CUDA path is the only path.

The directory name contains a hyphen (the project is called chaos-ultra); import it with
``importlib.import_module("chaos-ultra_b200")`` or through the ``chaos_ultra_b200`` shim at the
repository root.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
DEFAULT_KERNELS_DIR = PKG_DIR / "cudaKernels"
LIB_PATH = PKG_DIR / "lib" / "libchaos_ultra.so"

MAX_SUPER_SAMPLING = 64  # rendering/FractalRenderer.java:15

# ---------------------------------------------------------------------------------------------
# exceptions (chaos_status -> the Java exception of the reference)
# ---------------------------------------------------------------------------------------------


class ChaosError(RuntimeError):
    status = -1


class IllegalStateException(ChaosError):
    status = 1


class IllegalArgumentException(ChaosError, ValueError):
    status = 2


class CudaInitializationException(ChaosError):
    status = 3


class CudaException(ChaosError):
    status = 4


class FractalRendererException(ChaosError):
    status = 5


_STATUS_TO_EXC = {c.status: c for c in (IllegalStateException, IllegalArgumentException, CudaInitializationException,
                                         CudaException, FractalRendererException)}

# ---------------------------------------------------------------------------------------------
# C structs (include/chaos_ultra.h)
# ---------------------------------------------------------------------------------------------


class _Params(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("max_iterations", C.c_int32),
        ("segment", C.c_double * 4),
        ("max_super_sampling", C.c_float),
        ("use_adaptive_super_sampling", C.c_uint8),
        ("visualise_sample_count", C.c_uint8),
        ("use_foveated_rendering", C.c_uint8),
        ("use_sample_reuse", C.c_uint8),
        ("is_zooming", C.c_uint8),
        ("is_zooming_in", C.c_uint8),
        ("sample_reuse_cache_dirty", C.c_uint8),
        ("force_precision", C.c_uint8),
        ("mouse_focus", C.c_int32 * 2),
        ("float_precision", C.c_int32),
    ]


class _Defaults(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("has_segment", C.c_uint8),
        ("has_max_iterations", C.c_uint8),
        ("has_max_super_sampling", C.c_uint8),
        ("reserved", C.c_uint8),
        ("center_x", C.c_double),
        ("center_y", C.c_double),
        ("zoom", C.c_double),
        ("max_iterations", C.c_int32),
        ("max_super_sampling", C.c_float),
        ("custom_params", C.c_char * 256),
    ]


class _Stats(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("kernel_launches", C.c_uint32),
        ("render_ms", C.c_float),
        ("compose_ms", C.c_float),
        ("reuse_ms", C.c_float),
        ("frame_ms", C.c_float),
        ("pixel_iterations", C.c_uint64),
        ("samples", C.c_uint64),
        ("launches_total", C.c_uint64),
        ("foreign_orbits", C.c_uint64),
        ("skipped_iterations", C.c_uint64),
    ]


# numpy view of the 16-byte pixel_info_t record (helpers.cuh:106-130)
RECORD_DTYPE = np.dtype([("value", "<f4"), ("weight", "<f4"), ("isReused", "u1"), ("pad", "u1", (3,)),
                         ("weightOfNewSamples", "<f4")])

OUTPUT_HOST, OUTPUT_DEVICE = 0, 1
PRECISION_SINGLE, PRECISION_DOUBLE, PRECISION_TOO_BIG = 0, 1, 2
STATE_NOT_INITIALIZED, STATE_READY_TO_RENDER = 0, 1

# every entry point include/chaos_ultra.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
_API = {
    "chaos_provider_create": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_VP)]),
    "chaos_provider_destroy": (C.c_int, [_VP]),
    "chaos_list_fractals": (C.c_int, [_VP, C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(C.c_uint32)]),
    "chaos_register_module": (C.c_int, [_VP, C.c_char_p, C.c_char_p]),
    "chaos_open": (C.c_int, [_VP, C.c_char_p, C.c_int, C.POINTER(_VP)]),
    "chaos_active_renderer": (_VP, [_VP]),
    "chaos_initialize": (C.c_int, [_VP, C.c_uint32, C.c_uint32, _VP, C.c_uint32, C.c_int]),
    "chaos_free_resources": (C.c_int, [_VP]),
    "chaos_render_quality": (C.c_int, [_VP, C.POINTER(_Params)]),
    "chaos_render_fast": (C.c_int, [_VP, C.POINTER(_Params)]),
    "chaos_debug": (C.c_int, [_VP]),
    "chaos_set_custom_params": (C.c_int, [_VP, C.c_char_p]),
    "chaos_write_constant": (C.c_int, [_VP, C.c_char_p, _VP, C.c_size_t]),
    "chaos_supply_defaults": (C.c_int, [_VP, C.POINTER(_Defaults)]),
    "chaos_close": (C.c_int, [_VP]),
    "chaos_get_state": (C.c_int, [_VP]),
    "chaos_get_width": (C.c_uint32, [_VP]),
    "chaos_get_height": (C.c_uint32, [_VP]),
    "chaos_fractal_name": (C.c_char_p, [_VP]),
    "chaos_output_rgba": (_VP, [_VP]),
    "chaos_output_rgba_device": (C.c_uint64, [_VP]),
    "chaos_download_rgba": (C.c_int, [_VP, _VP, C.c_size_t]),
    "chaos_download_records": (C.c_int, [_VP, _VP, C.c_size_t]),
    "chaos_get_stats": (C.c_int, [_VP, C.POINTER(_Stats)]),
    "chaos_debug_peek_counters": (C.c_int, [_VP, _VP, C.c_size_t]),
    "chaos_set_partition": (C.c_int, [_VP, C.c_uint32, C.c_uint32, C.c_uint32]),
    "chaos_set_output_target": (C.c_int, [_VP, C.c_uint64]),
    "chaos_ipc_export_frame": (C.c_int, [_VP, _VP]),
    "chaos_ipc_open_frame": (C.c_int, [_VP, _VP]),
    "chaos_ipc_export_records": (C.c_int, [_VP, _VP]),
    "chaos_ipc_open_records": (C.c_int, [_VP, C.c_uint32, _VP]),
    "chaos_set_host_target": (C.c_int, [_VP, _VP, C.c_size_t]),
    "chaos_set_frame_barrier": (C.c_int, [_VP, _VP, C.c_uint32]),
    "chaos_last_error": (C.c_char_p, []),
    "chaos_abi_version": (C.c_uint32, []),
}

_lib = None


def load_library(path: Optional[os.PathLike] = None) -> C.CDLL:
    """Load libchaos_ultra.so and bind every declared entry point.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise CudaInitializationException(
            f"{p} is missing: the native backend is not built (run `python __graft_entry__.py build`); "
            "there is no CPU fallback")
    lib = C.CDLL(str(p))
    for name, (res, args) in _API.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _check(lib, status: int):
    if status != 0:
        msg = (lib.chaos_last_error() or b"").decode("utf-8", "replace")
        raise _STATUS_TO_EXC.get(status, ChaosError)(msg)


# ---------------------------------------------------------------------------------------------
# model (rendering/Model.java, rendering/model/*.java)
# ---------------------------------------------------------------------------------------------


@dataclasses.dataclass
class RenderingModel:
    """The values a FractalRenderer reads from (and writes back to) the application's model."""
    canvasWidth: int = 0
    canvasHeight: int = 0
    planeSegment: List[float] = dataclasses.field(default_factory=lambda: [0.0, 0.0, 0.0, 0.0])  # lb.x lb.y rt.x rt.y
    maxIterations: int = 200
    maxSuperSampling: float = 2.0
    useAdaptiveSuperSampling: bool = True
    visualiseSampleCount: bool = False
    useFoveatedRendering: bool = True
    useSampleReuse: bool = True
    zooming: bool = False
    zoomingIn: bool = False
    mouseFocus: Sequence[int] = (0, 0)
    sampleReuseCacheDirty: bool = False
    floatingPointPrecision: int = PRECISION_SINGLE
    fractalCustomParams: str = ""
    forcePrecision: int = 0  # not in the reference: 0 = its rule, 1 = single, 2 = double (benchmarks/tests)

    def setMaxSuperSampling(self, v: float):
        # Model.java:166-168 clamps to [0, MAX_SUPER_SAMPLING]
        self.maxSuperSampling = max(0.0, min(float(v), float(MAX_SUPER_SAMPLING)))

    def setPlaneSegmentFromCenter(self, centerX: float, centerY: float, zoom: float):
        # Model.java:247-256
        relH = 1.0
        relW = relH / float(self.canvasHeight) * self.canvasWidth
        self.planeSegment = [centerX - relW * zoom / 2, centerY - relH * zoom / 2,
                             centerX + relW * zoom / 2, centerY + relH * zoom / 2]

    def resetRenderingValuesToDefault(self):
        # Model.java:284-294
        self.setPlaneSegmentFromCenter(0, 0, 4)
        self.maxIterations = 200
        self.setMaxSuperSampling(2)
        self.useAdaptiveSuperSampling = True
        self.useFoveatedRendering = True
        self.useSampleReuse = True
        self.visualiseSampleCount = False
        self.zooming = False

    def zoomAt(self, where: Sequence[int], into: bool):
        # RenderingController.zoomAt, rendering/RenderingController.java:130-150 (ZOOM_COEFF = 0.977f :19)
        zc32 = float(np.float32(0.977))
        lbx, lby, rtx, rty = self.planeSegment
        sw, sh = rtx - lbx, rty - lby
        relTop = where[1] / float(self.canvasHeight)
        relBtm = 1 - relTop
        relLeft = where[0] / float(self.canvasWidth)
        relRght = 1 - relLeft
        cx = lbx + sw * relLeft
        cy = lby + sh * relBtm
        zc = zc32 if into else float(np.float32(2.0) - np.float32(0.977))
        self.planeSegment = [cx - sw * relLeft * zc, cy - sh * relBtm * zc, cx + sw * relRght * zc, cy + sh * relTop * zc]

    def copy(self) -> "RenderingModel":
        m = dataclasses.replace(self)
        m.planeSegment = list(self.planeSegment)
        return m

    def _to_c(self) -> _Params:
        p = _Params()
        p.struct_size = C.sizeof(_Params)
        p.max_iterations = int(self.maxIterations)
        for i in range(4):
            p.segment[i] = float(self.planeSegment[i])
        p.max_super_sampling = float(self.maxSuperSampling)
        p.use_adaptive_super_sampling = int(bool(self.useAdaptiveSuperSampling))
        p.visualise_sample_count = int(bool(self.visualiseSampleCount))
        p.use_foveated_rendering = int(bool(self.useFoveatedRendering))
        p.use_sample_reuse = int(bool(self.useSampleReuse))
        p.is_zooming = int(bool(self.zooming))
        p.is_zooming_in = int(bool(self.zoomingIn))
        p.sample_reuse_cache_dirty = int(bool(self.sampleReuseCacheDirty))
        p.force_precision = int(self.forcePrecision)
        p.mouse_focus[0] = int(self.mouseFocus[0])
        p.mouse_focus[1] = int(self.mouseFocus[1])
        return p

    def _from_c(self, p: _Params):
        self.sampleReuseCacheDirty = bool(p.sample_reuse_cache_dirty)
        self.floatingPointPrecision = int(p.float_precision)


@dataclasses.dataclass
class RenderStats:
    kernel_launches: int
    render_ms: float
    compose_ms: float
    reuse_ms: float
    pixel_iterations: int
    samples: int
    launches_total: int
    skipped_iterations: int = 0   # part of pixel_iterations proven (exact recurrence) instead of executed
    frame_ms: float = 0.0         # device time of the whole frame (render kernels + compose)
    foreign_orbits: int = 0       # several GPUs: orbits of this rank's tiles that other ranks iterated (tile stealing)


# ---------------------------------------------------------------------------------------------
# palettes (util/ImageHelpers.java)
# ---------------------------------------------------------------------------------------------


def createDefaultColorPalette() -> np.ndarray:
    """ImageHelpers.createDefaultColorPalette, util/ImageHelpers.java:77-111: 6 x 256 linear ramps,
    R in the low byte, alpha 255."""
    mx, full = 256, 255
    p = np.zeros(mx * 6, dtype=np.uint32)

    def rgb(r, g, b):
        return (r & 0xFF) | ((g & 0xFF) << 8) | ((b & 0xFF) << 16) | 0xFF000000

    for i in range(0, mx):
        p[i] = rgb(i, 0, min(full, full // 2 + i))
    for i in range(mx, 2 * mx):
        p[i] = rgb(full, 0, full - i)
    for i in range(2 * mx, 3 * mx):
        p[i] = rgb(full, i, 0)
    for i in range(3 * mx, 4 * mx):
        p[i] = rgb(full - i, full, 0)
    for i in range(4 * mx, 5 * mx):
        p[i] = rgb(0, full, i)
    for i in range(5 * mx, 6 * mx):
        p[i] = rgb(0, full - i, full)
    return p


def loadColorPaletteFromFile(path) -> np.ndarray:
    """ImageHelpers.loadColorPaletteFromFile, util/ImageHelpers.java:58-72,118-130: first row of the
    image as RGBA with R in the low byte."""
    from PIL import Image  # optional dependency, only for PNG palettes

    img = Image.open(path).convert("RGBA")
    row = np.asarray(img)[0].astype(np.uint32)
    return (row[:, 0] | (row[:, 1] << 8) | (row[:, 2] << 16) | (row[:, 3] << 24)).astype(np.uint32)


# ---------------------------------------------------------------------------------------------
# renderer / provider
# ---------------------------------------------------------------------------------------------


class CudaFractalRenderer:
    """FractalRenderer over the C ABI.  Lifecycle (CudaFractalRenderer.java:27-30):
    notInitialized --initializeRendering()--> readyToRender --freeRenderingResources()--> notInitialized"""

    def __init__(self, lib, handle, provider):
        self._lib = lib
        self._h = handle
        self._provider = provider
        self._mode = OUTPUT_HOST

    # -- FractalRenderer interface ---------------------------------------------------------
    def initializeRendering(self, width: int, height: int, palette: Optional[np.ndarray] = None,
                            output_mode: int = OUTPUT_HOST):
        """initializeRendering(GLParams): the GL output texture and palette texture become a size, a palette
        array and an output mode."""
        pal = np.ascontiguousarray(createDefaultColorPalette() if palette is None else palette, dtype=np.uint32)
        _check(self._lib, self._lib.chaos_initialize(self._h, int(width), int(height), pal.ctypes.data, pal.size,
                                                     int(output_mode)))
        self._mode = int(output_mode)

    def freeRenderingResources(self):
        _check(self._lib, self._lib.chaos_free_resources(self._h))

    def getState(self) -> int:
        return int(self._lib.chaos_get_state(self._h))

    def getWidth(self) -> int:
        return int(self._lib.chaos_get_width(self._h))

    def getHeight(self) -> int:
        return int(self._lib.chaos_get_height(self._h))

    def launchDebugKernel(self):
        _check(self._lib, self._lib.chaos_debug(self._h))

    def renderFast(self, model: RenderingModel):
        p = model._to_c()
        st = self._lib.chaos_render_fast(self._h, C.byref(p))
        _check(self._lib, st)
        model._from_c(p)

    def renderQuality(self, model: RenderingModel):
        p = model._to_c()
        st = self._lib.chaos_render_quality(self._h, C.byref(p))
        _check(self._lib, st)
        model._from_c(p)

    def close(self):
        if self._h:
            _check(self._lib, self._lib.chaos_close(self._h))
            self._h = None
            if self._provider is not None and self._provider._active is self:
                self._provider._active = None

    def setFractalCustomParams(self, text: str):
        _check(self._lib, self._lib.chaos_set_custom_params(self._h, text.encode()))

    def getFractalName(self) -> str:
        return self._lib.chaos_fractal_name(self._h).decode()

    def supplyDefaultValues(self, model: RenderingModel):
        d = _Defaults()
        d.struct_size = C.sizeof(_Defaults)
        _check(self._lib, self._lib.chaos_supply_defaults(self._h, C.byref(d)))
        model.fractalCustomParams = d.custom_params.decode()
        if d.has_segment:
            model.setPlaneSegmentFromCenter(d.center_x, d.center_y, d.zoom)
        if d.has_max_iterations:
            model.maxIterations = int(d.max_iterations)
        if d.has_max_super_sampling:
            model.setMaxSuperSampling(d.max_super_sampling)

    # -- additions of this backend ---------------------------------------------------------
    def writeToConstantMemory(self, name: str, data: bytes):
        """FractalRenderingModule.writeToConstantMemory :163-227 (size-checked)."""
        buf = C.create_string_buffer(bytes(data), len(data))
        _check(self._lib, self._lib.chaos_write_constant(self._h, name.encode(), C.cast(buf, _VP), len(data)))

    def setPartition(self, part_index: int, part_count: int, band_rows: int = 64):
        _check(self._lib, self._lib.chaos_set_partition(self._h, part_index, part_count, band_rows))

    def outputRGBA(self) -> np.ndarray:
        """The composed frame as an (H, W) uint32 array (R in the low byte, row 0 = top).  HOST mode: a view
        of the library's pinned buffer (valid until freeRenderingResources); DEVICE mode: a download."""
        h, w = self.getHeight(), self.getWidth()
        if self._mode == OUTPUT_HOST:
            ptr = self._lib.chaos_output_rgba(self._h)
            if not ptr:
                raise IllegalStateException("Renderer has to be initialized first")
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=(h, w))
        out = np.empty((h, w), dtype=np.uint32)
        _check(self._lib, self._lib.chaos_download_rgba(self._h, out.ctypes.data, out.nbytes))
        return out

    def outputRGBADevicePointer(self) -> int:
        return int(self._lib.chaos_output_rgba_device(self._h))

    def setOutputTarget(self, device_ptr: int) -> None:
        """DEVICE mode: compose writes into ``device_ptr`` (e.g. rank 0's frame mapped with CUDA IPC) instead of the own frame; 0 resets."""
        _check(self._lib, self._lib.chaos_set_output_target(self._h, int(device_ptr)))

    def exportFrameHandle(self) -> bytes:
        """DEVICE mode: a 64-byte handle of this renderer's device frame that another process can open."""
        buf = C.create_string_buffer(64)
        _check(self._lib, self._lib.chaos_ipc_export_frame(self._h, C.cast(buf, _VP)))
        return buf.raw

    def openFrameHandle(self, handle: bytes) -> None:
        """DEVICE mode: compose into the frame another process exported (its bands cross NVLink as compose's stores)."""
        buf = C.create_string_buffer(bytes(handle), 64)
        _check(self._lib, self._lib.chaos_ipc_open_frame(self._h, C.cast(buf, _VP)))

    def exportRecordHandles(self) -> bytes:
        """192 bytes: handles of this renderer's two record buffers and of its scheduler counters (multi-GPU fast frames, tile stealing)."""
        buf = C.create_string_buffer(192)
        _check(self._lib, self._lib.chaos_ipc_export_records(self._h, C.cast(buf, _VP)))
        return buf.raw

    def openRecordHandles(self, peer_rank: int, handles: bytes) -> None:
        buf = C.create_string_buffer(bytes(handles), 192)
        _check(self._lib, self._lib.chaos_ipc_open_records(self._h, int(peer_rank), C.cast(buf, _VP)))

    def setHostTarget(self, address: int, nbytes: int) -> None:
        """DEVICE mode: compose into caller-owned, page-aligned host memory (e.g. shared memory all ranks mapped); 0 releases it."""
        _check(self._lib, self._lib.chaos_set_host_target(self._h, int(address) or None, int(nbytes)))

    def setFrameBarrier(self, address: int, world: int) -> None:
        """64 zeroed bytes of host memory shared by all ranks: every render call returns when all ranks' frames are done."""
        _check(self._lib, self._lib.chaos_set_frame_barrier(self._h, int(address) or None, int(world)))

    def downloadRecords(self) -> np.ndarray:
        """The primary pixel_info_t buffer as an (H, W) structured array."""
        h, w = self.getHeight(), self.getWidth()
        out = np.zeros((h, w), dtype=RECORD_DTYPE)
        _check(self._lib, self._lib.chaos_download_records(self._h, out.ctypes.data, out.nbytes))
        return out

    def stats(self) -> RenderStats:
        s = _Stats()
        s.struct_size = C.sizeof(_Stats)
        _check(self._lib, self._lib.chaos_get_stats(self._h, C.byref(s)))
        return RenderStats(s.kernel_launches, s.render_ms, s.compose_ms, s.reuse_ms, s.pixel_iterations, s.samples, s.launches_total, s.skipped_iterations, s.frame_ms, s.foreign_orbits)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class CudaFractalRendererProvider:
    """FractalRendererProvider over the C ABI: catalogue of fractals, at most one active renderer."""

    def __init__(self, kernels_dir: Optional[os.PathLike] = None, device: int = 0):
        self._lib = load_library()
        kd = Path(kernels_dir) if kernels_dir is not None else DEFAULT_KERNELS_DIR
        h = _VP()
        _check(self._lib, self._lib.chaos_provider_create(str(kd).encode(), int(device), C.byref(h)))
        self._h = h
        self._active: Optional[CudaFractalRenderer] = None

    def getAvailableFractals(self) -> List[str]:
        n = C.c_uint32(0)
        _check(self._lib, self._lib.chaos_list_fractals(self._h, None, 0, C.byref(n)))
        arr = (C.c_char_p * n.value)()
        _check(self._lib, self._lib.chaos_list_fractals(self._h, arr, n.value, C.byref(n)))
        return [a.decode() for a in arr]

    def registerModule(self, fractalName: str, fileStem: str) -> None:
        """a fractal author's module <kernels_dir>/<fileStem>.cubin under a display name of its own"""
        _check(self._lib, self._lib.chaos_register_module(self._h, fractalName.encode(), fileStem.encode()))

    def getDefaultRenderer(self) -> CudaFractalRenderer:
        return self.getRenderer("mandelbrot", False)

    def getRenderer(self, fractalName: str, forceReload: bool = False) -> CudaFractalRenderer:
        h = _VP()
        old = self._active
        same = old is not None and old._h is not None and not forceReload and old.getFractalName() == fractalName
        st = self._lib.chaos_open(self._h, fractalName.encode(), int(bool(forceReload)), C.byref(h))
        if st != 0 and old is not None and not self._lib.chaos_active_renderer(self._h):
            old._h = None      # the library closed it before the new module failed to load (:52)
            self._active = None
        _check(self._lib, st)
        if same:
            return old  # same name and !forceReload: the active renderer is returned (:50-51)
        if old is not None:
            old._h = None  # closed by the library (:52)
        self._active = CudaFractalRenderer(self._lib, h, self)
        return self._active

    def close(self):
        if self._h:
            if self._active is not None:
                self._active._h = None
                self._active = None
            _check(self._lib, self._lib.chaos_provider_destroy(self._h))
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


__all__ = [
    "CudaFractalRendererProvider", "CudaFractalRenderer", "RenderingModel", "RenderStats", "ChaosError",
    "IllegalStateException", "IllegalArgumentException", "CudaInitializationException", "CudaException",
    "FractalRendererException", "createDefaultColorPalette", "loadColorPaletteFromFile", "load_library",
    "RECORD_DTYPE", "OUTPUT_HOST", "OUTPUT_DEVICE", "PRECISION_SINGLE", "PRECISION_DOUBLE", "PRECISION_TOO_BIG",
    "STATE_NOT_INITIALIZED", "STATE_READY_TO_RENDER", "MAX_SUPER_SAMPLING", "DEFAULT_KERNELS_DIR", "LIB_PATH",
]
