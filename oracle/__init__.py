"""ctypes bindings of the CPU oracle (liboracle.so) and of the reference-kernel harness (refrun).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (chaos-ultra_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Optional

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_LIB = HERE / "liboracle.so"
REF_DIR = HERE / "_ref"
REFRUN_LIB = REF_DIR / "librefrun.so"

RECORD_DTYPE = np.dtype([("value", "<f4"), ("weight", "<f4"), ("isReused", "u1"), ("pad", "u1", (3,)),
                         ("weightOfNewSamples", "<f4")])

FRACTAL_KINDS = {"mandelbrot": 0, "julia": 1, "test": 2}
FLAG_ADAPTIVE_SS, FLAG_FOVEATION, FLAG_SAMPLE_REUSE, FLAG_IS_ZOOMING, FLAG_ZOOMING_IN = 1, 4, 8, 16, 32
VARIANT_NVCC129, VARIANT_SHIPPED_PTX = 0, 1


class _Fractal(C.Structure):
    _fields_ = [("kind", C.c_int), ("julia_c", C.c_double * 2), ("amplifier", C.c_int)]


class _Frame(C.Structure):
    _fields_ = [
        ("fractal", C.POINTER(_Fractal)),
        ("real_is_double", C.c_int),
        ("variant", C.c_int),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("image", C.c_double * 4),
        ("image_reused", C.c_double * 4),
        ("focus_x", C.c_uint32),
        ("focus_y", C.c_uint32),
        ("maxIter", C.c_uint32),
        ("maxSuperSampling", C.c_float),
        ("flags", C.c_uint32),
        ("trips", C.POINTER(C.c_uint64)),
        ("sample_hist", C.POINTER(C.c_uint64)),
        ("ub_reads", C.POINTER(C.c_uint64)),
    ]


def build(force: bool = False):
    """make -C oracle (liboracle.so always; _ref/ when /root/reference is present)."""
    if force or not ORACLE_LIB.exists() or ORACLE_LIB.stat().st_mtime < max(
            (HERE / n).stat().st_mtime for n in ("chaos_oracle.c", "chaos_oracle_real.inc", "chaos_oracle.h")):
        subprocess.run(["make", "-C", str(HERE), "liboracle.so"], check=True, capture_output=True)
    return ORACLE_LIB


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(str(ORACLE_LIB))
        l.ora_render_main.restype = C.c_int
        l.ora_render_main.argtypes = [C.POINTER(_Frame), C.c_void_p, C.c_size_t]
        l.ora_render_main_rows.restype = C.c_int
        l.ora_render_main_rows.argtypes = [C.POINTER(_Frame), C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32]
        l.ora_render_advanced.restype = C.c_int
        l.ora_render_advanced.argtypes = [C.POINTER(_Frame), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        l.ora_compose.restype = C.c_int
        l.ora_compose.argtypes = [C.POINTER(_Frame), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
        l.ora_default_palette.argtypes = [C.c_void_p]
        l.ora_choose_precision.restype = C.c_int
        l.ora_choose_precision.argtypes = [C.POINTER(C.c_double), C.c_uint32, C.c_uint32]
        l.ora_segment_from_center.argtypes = [C.c_double, C.c_double, C.c_double, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
        l.ora_zoom_at.argtypes = [C.POINTER(C.c_double), C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int]
        l.ora_scalar_rows.restype = C.c_uint64
        l.ora_scalar_rows.argtypes = [C.POINTER(_Frame), C.c_uint32, C.c_uint32, C.c_void_p]
        _lib = l
    return _lib


class Result:
    def __init__(self, records, trips, sample_hist, ub_reads):
        self.records = records
        self.pixel_iterations = trips
        self.sample_hist = sample_hist
        self.samples = int(sample_hist.sum())
        self.ub_reads = ub_reads


def _frame(fractal, W, H, image, maxIter, maxSS, flags, double, variant, julia_c, amplifier, image_reused=None,
           focus=(0, 0)):
    fr = _Fractal()
    fr.kind = FRACTAL_KINDS[fractal]
    fr.julia_c[0], fr.julia_c[1] = julia_c
    fr.amplifier = amplifier
    f = _Frame()
    f.fractal = C.pointer(fr)
    f.real_is_double = int(bool(double))
    f.variant = variant
    f.width, f.height = W, H
    for i in range(4):
        f.image[i] = float(image[i])
        f.image_reused[i] = float(image_reused[i]) if image_reused is not None else 0.0
    f.focus_x, f.focus_y = int(focus[0]) & 0xFFFFFFFF, int(focus[1]) & 0xFFFFFFFF
    f.maxIter = int(maxIter)
    f.maxSuperSampling = float(maxSS)
    f.flags = int(flags)
    trips = C.c_uint64(0)
    ub = C.c_uint64(0)
    hist = (C.c_uint64 * 64)()
    f.trips = C.pointer(trips)
    f.ub_reads = C.pointer(ub)
    f.sample_hist = C.cast(hist, C.POINTER(C.c_uint64))
    return f, (fr, trips, ub, hist)


def render_main(fractal, W, H, image, maxIter, maxSS, flags, double=True, variant=VARIANT_NVCC129,
                julia_c=(0.0, 0.0), amplifier=10) -> Result:
    f, keep = _frame(fractal, W, H, image, maxIter, maxSS, flags, double, variant, julia_c, amplifier)
    out = np.zeros((H, W), dtype=RECORD_DTYPE)
    rc = lib().ora_render_main(C.byref(f), out.ctypes.data, W * 16)
    if rc != 0:
        raise ValueError("oracle: maxSuperSampling must be >= 1 for the main kernel")
    _, trips, ub, hist = keep
    return Result(out, trips.value, np.array(list(hist), dtype=np.uint64), ub.value)


def render_main_rows_threaded(fractal, W, H, image, maxIter, maxSS, flags, double=True, row_stride=1, threads=1,
                              julia_c=(0.0, 0.0)):
    """Host baseline: the oracle's main kernel over every `row_stride`-th vote-tile row, rows dealt to `threads`
    Python threads (ctypes releases the GIL).  Returns (pixel_iterations, samples, seconds)."""
    import threading
    import time
    out = np.zeros((H, W), dtype=RECORD_DTYPE)
    n_rows = (H + 3) // 4
    picked = list(range(0, n_rows, row_stride))
    results = [None] * threads

    def work(t):
        f, keep = _frame(fractal, W, H, image, maxIter, maxSS, flags, double, VARIANT_NVCC129, julia_c, 10)
        for tr in picked[t::threads]:
            lib().ora_render_main_rows(C.byref(f), out.ctypes.data, W * 16, tr, tr + 1, 1)
        results[t] = (keep[1].value, int(np.array(list(keep[3]), dtype=np.uint64).sum()))

    lib()
    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    return sum(r[0] for r in results), sum(r[1] for r in results), dt


def render_advanced(fractal, W, H, image, maxIter, maxSS, flags, image_reused, in_records, focus, double=True,
                    variant=VARIANT_NVCC129, julia_c=(0.0, 0.0), amplifier=10) -> Result:
    f, keep = _frame(fractal, W, H, image, maxIter, maxSS, flags, double, variant, julia_c, amplifier, image_reused, focus)
    inp = np.ascontiguousarray(in_records)
    assert inp.dtype == RECORD_DTYPE and inp.shape == (H, W)
    out = np.zeros((H, W), dtype=RECORD_DTYPE)
    rc = lib().ora_render_advanced(C.byref(f), out.ctypes.data, W * 16, inp.ctypes.data, W * 16)
    assert rc == 0
    _, trips, ub, hist = keep
    return Result(out, trips.value, np.array(list(hist), dtype=np.uint64), ub.value)


def compose(fractal, records, palette, maxSS=1.0, visualize=False) -> np.ndarray:
    H, W = records.shape
    f, keep = _frame(fractal, W, H, (0, 0, 1, 1), 1, maxSS, 0, True, 0, (0, 0), 10)
    rec = np.ascontiguousarray(records)
    pal = np.ascontiguousarray(palette, dtype=np.uint32)
    out = np.zeros((H, W), dtype=np.uint32)
    rc = lib().ora_compose(C.byref(f), rec.ctypes.data, W * 16, out.ctypes.data, pal.ctypes.data, pal.size, int(visualize))
    assert rc == 0
    return out


def default_palette() -> np.ndarray:
    p = np.zeros(1536, dtype=np.uint32)
    lib().ora_default_palette(p.ctypes.data)
    return p


def choose_precision(image, W, H) -> int:
    arr = (C.c_double * 4)(*map(float, image))
    return int(lib().ora_choose_precision(arr, W, H))


def segment_from_center(cx, cy, zoom, W, H):
    arr = (C.c_double * 4)()
    lib().ora_segment_from_center(cx, cy, zoom, W, H, arr)
    return list(arr)


def zoom_at(image, W, H, where, into=True):
    arr = (C.c_double * 4)(*map(float, image))
    lib().ora_zoom_at(arr, W, H, int(where[0]), int(where[1]), int(bool(into)))
    return list(arr)


def scalar_rows(fractal, W, H, image, maxIter, y0, y1, julia_c=(0.0, 0.0), want_image=False):
    """Plain scalar escape loop, 1 sample/pixel, rows [y0,y1): the host baseline of BASELINE.md section 2."""
    f, keep = _frame(fractal, W, H, image, maxIter, 1.0, 0, True, 0, julia_c, 10)
    et = np.zeros((H, W), dtype=np.uint32) if want_image else None
    trips = lib().ora_scalar_rows(C.byref(f), y0, y1, et.ctypes.data if et is not None else None)
    return int(trips), et


# ---------------------------------------------------------------------------------------------
# reference kernels on the GPU (oracle/_ref, built from /root/reference by oracle/Makefile)
# ---------------------------------------------------------------------------------------------


class RefRun:
    """Runs the reference's own CUDA kernels (see refrun.cpp).  kind: 'src' (nvcc-12.9 build of the reference
    sources) or 'ptx92' (the shipped CUDA 9.2 PTX assembled for sm_100a)."""

    def __init__(self, fractal_file: str, kind: str = "src", device: int = 0):
        if not REFRUN_LIB.exists():
            raise FileNotFoundError(f"{REFRUN_LIB} missing: run `make -C oracle` where /root/reference is mounted")
        path = REF_DIR / f"{fractal_file}.{kind}.cubin"
        if not path.exists():
            raise FileNotFoundError(str(path))
        l = C.CDLL(str(REFRUN_LIB))
        l.refrun_last_error.restype = C.c_char_p
        l.refrun_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        l.refrun_close.argtypes = [C.c_void_p]
        l.refrun_write_constant.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
        l.refrun_main.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.c_uint32,
                                  C.c_float, C.c_uint32, C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        l.refrun_advanced.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.c_uint32,
                                      C.c_float, C.c_uint32, C.POINTER(C.c_double), C.c_void_p, C.c_uint32, C.c_uint32,
                                      C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        l.refrun_compose.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_float,
                                     C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        l.refrun_frames.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.c_uint32, C.c_float,
                                    C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                    C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]
        l.refrun_zoom.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_uint32,
                                  C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]
        self._l = l
        self._h = C.c_void_p()
        self._ok(l.refrun_open(str(path).encode(), device, C.byref(self._h)))
        self.last_ms = 0.0

    def _ok(self, rc):
        if rc != 0:
            raise RuntimeError("refrun: " + self._l.refrun_last_error().decode())

    def write_constant(self, name: str, data: bytes):
        buf = C.create_string_buffer(bytes(data), len(data))
        self._ok(self._l.refrun_write_constant(self._h, name.encode(), buf, len(data)))

    def main(self, W, H, image, maxIter, maxSS, flags, double=True, reps=1, want_records=True) -> Optional[np.ndarray]:
        img = (C.c_double * 4)(*map(float, image))
        out = np.zeros((H, W), dtype=RECORD_DTYPE) if want_records else None
        ms = C.c_float(0)
        self._ok(self._l.refrun_main(self._h, int(double), W, H, img, maxIter, maxSS, flags,
                                     out.ctypes.data if out is not None else None, reps, C.byref(ms)))
        self.last_ms = ms.value
        return out

    def advanced(self, W, H, image, maxIter, maxSS, flags, image_reused, in_records, focus, double=True, reps=1,
                 want_records=True):
        img = (C.c_double * 4)(*map(float, image))
        old = (C.c_double * 4)(*map(float, image_reused))
        inp = np.ascontiguousarray(in_records)
        out = np.zeros((H, W), dtype=RECORD_DTYPE) if want_records else None
        ms = C.c_float(0)
        self._ok(self._l.refrun_advanced(self._h, int(double), W, H, img, maxIter, maxSS, flags, old, inp.ctypes.data,
                                         int(focus[0]) & 0xFFFFFFFF, int(focus[1]) & 0xFFFFFFFF,
                                         out.ctypes.data if out is not None else None, reps, C.byref(ms)))
        self.last_ms = ms.value
        return out

    def compose(self, records, palette, maxSS=1.0, visualize=False, reps=1) -> np.ndarray:
        H, W = records.shape
        rec = np.ascontiguousarray(records)
        pal = np.ascontiguousarray(palette, dtype=np.uint32)
        out = np.zeros((H, W), dtype=np.uint32)
        ms = C.c_float(0)
        self._ok(self._l.refrun_compose(self._h, W, H, rec.ctypes.data, pal.ctypes.data, pal.size, maxSS, int(visualize),
                                        out.ctypes.data, reps, C.byref(ms)))
        self.last_ms = ms.value
        return out

    def frames(self, W, H, image, maxIter, maxSS, flags, palette, double=True, warmup=1, steps=1, to_host=True):
        """The reference host's quality-frame loop; returns (wall_ms, main_ms_sum, compose_ms_sum, rgba|None)."""
        img = (C.c_double * 4)(*map(float, image))
        pal = np.ascontiguousarray(palette, dtype=np.uint32)
        wall, m, c = C.c_double(0), C.c_float(0), C.c_float(0)
        out = np.zeros((H, W), dtype=np.uint32) if to_host else None
        self._ok(self._l.refrun_frames(self._h, int(double), W, H, img, maxIter, maxSS, flags, pal.ctypes.data, pal.size,
                                       warmup, steps, int(to_host), C.byref(wall), C.byref(m), C.byref(c),
                                       out.ctypes.data if out is not None else None))
        return wall.value, m.value, c.value, out

    def zoom(self, W, H, segments, doubles, maxIter, maxSS, flags, focus, palette, warmup=1, steps=1, to_host=True):
        """The reference host's zoom loop (frame 0 quality, then fast frames); segments: list of 1+warmup+steps segments.
        Returns (wall_ms, advanced_ms_sum, compose_ms_sum, last rgba|None)."""
        n = 1 + warmup + steps
        assert len(segments) == n and len(doubles) == n
        seg = (C.c_double * (4 * n))(*[float(v) for sgm in segments for v in sgm])
        dbl = (C.c_int * n)(*[int(bool(d)) for d in doubles])
        pal = np.ascontiguousarray(palette, dtype=np.uint32)
        wall, a, c = C.c_double(0), C.c_float(0), C.c_float(0)
        out = np.zeros((H, W), dtype=np.uint32) if to_host else None
        self._ok(self._l.refrun_zoom(self._h, W, H, seg, dbl, maxIter, maxSS, flags, int(focus[0]), int(focus[1]), pal.ctypes.data,
                                     pal.size, warmup, steps, int(to_host), C.byref(wall), C.byref(a), C.byref(c),
                                     out.ctypes.data if out is not None else None))
        return wall.value, a.value, c.value, out

    def close(self):
        if self._h:
            self._l.refrun_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
