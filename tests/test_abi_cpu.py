"""CPU-only checks: the C-ABI library loads without a driver, exports every symbol include/chaos_ultra.h
declares, fails loudly (no CPU fallback) when there is no device, and the host-side mirror of the reference
model (zoomAt, setPlaneSegmentFromCenter, default palette, precision rule) agrees with the oracle's
restatement of the Java code."""
import ctypes
import importlib
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import cases
import oracle

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "chaos_ultra.h"


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def declared_symbols():
    text = HEADER.read_text()
    return sorted(set(re.findall(r"CHAOS_API[^;(]*?\b(chaos_\w+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("chaos_provider_create", "chaos_open", "chaos_initialize", "chaos_render_quality", "chaos_render_fast",
                 "chaos_debug", "chaos_set_custom_params", "chaos_write_constant", "chaos_supply_defaults", "chaos_close",
                 "chaos_free_resources", "chaos_last_error"):
        assert must in syms
    assert len(syms) >= 25


def test_library_exports_every_declared_symbol(cu):
    lib = cu.load_library()
    out = subprocess.run(["nm", "-D", "--defined-only", str(cu.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (chaos_\w+)", out))
    for s in declared_symbols():
        assert s in exported, s
        assert hasattr(lib, s)
    # nothing but the declared boundary leaks out
    assert exported == set(declared_symbols())
    assert lib.chaos_abi_version() == 1
    # the python binding tables (renderer/provider in the package, frame driver in driver.py) cover the header
    drv = importlib.import_module("chaos-ultra_b200.driver")
    assert set(cu._API) | set(drv._DRIVER_API) == set(declared_symbols())


def test_library_has_no_torch_or_cudart_dependency(cu):
    out = subprocess.run(["ldd", str(cu.LIB_PATH)], capture_output=True, text=True).stdout
    assert "torch" not in out and "cudart" not in out and "libcuda" not in out  # libcuda is dlopen'ed on demand


def test_struct_layouts_match_header(cu):
    # compile a tiny C program that prints sizeof/offsetof from the real header and compare with ctypes
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "chaos_ultra.h"
    int main(void){
      printf("%zu %zu %zu %zu %zu %zu\n", sizeof(chaos_params), offsetof(chaos_params, segment), offsetof(chaos_params, max_super_sampling),
             offsetof(chaos_params, mouse_focus), offsetof(chaos_params, float_precision), offsetof(chaos_params, force_precision));
      printf("%zu %zu %zu\n", sizeof(chaos_defaults), offsetof(chaos_defaults, center_x), offsetof(chaos_defaults, custom_params));
      printf("%zu %zu %zu %zu\n", sizeof(chaos_stats), offsetof(chaos_stats, pixel_iterations), offsetof(chaos_stats, launches_total), offsetof(chaos_stats, reuse_ms));
      return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = Path(d) / "t.c"
        c.write_text(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), str(c), "-o", str(Path(d) / "t")], check=True)
        lines = subprocess.run([str(Path(d) / "t")], capture_output=True, text=True, check=True).stdout.split("\n")
    P, D, S = cu._Params, cu._Defaults, cu._Stats
    assert list(map(int, lines[0].split())) == [ctypes.sizeof(P), P.segment.offset, P.max_super_sampling.offset,
                                                P.mouse_focus.offset, P.float_precision.offset, P.force_precision.offset]
    assert list(map(int, lines[1].split())) == [ctypes.sizeof(D), D.center_x.offset, D.custom_params.offset]
    assert list(map(int, lines[2].split())) == [ctypes.sizeof(S), S.pixel_iterations.offset, S.launches_total.offset, S.reuse_ms.offset]


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a device")
def test_no_cpu_fallback_without_device(cu):
    with pytest.raises(cu.CudaInitializationException) as e:
        cu.CudaFractalRendererProvider()
    assert "CUDA" in str(e.value) or "Cuda" in str(e.value)


def test_null_and_bad_arguments_do_not_crash(cu):
    lib = cu.load_library()
    assert lib.chaos_provider_destroy(None) == 2
    assert b"NULL" in lib.chaos_last_error()
    assert lib.chaos_render_quality(None, None) == 2
    assert lib.chaos_close(None) == 2
    assert lib.chaos_get_width(None) == 0
    assert lib.chaos_get_state(None) == 0
    h = ctypes.c_void_p()
    assert lib.chaos_provider_create(None, 0, ctypes.byref(h)) == 2


def test_module_cubins_are_built_for_sm100a_and_export_the_contract(cu):
    # FractalRenderingModule.java:91-97: all seven kernels are looked up eagerly, by these names
    names = ["fractalRenderMainFloat", "fractalRenderMainDouble", "fractalRenderAdvancedFloat", "fractalRenderAdvancedDouble",
             "fractalRenderUnderSampled", "compose", "debug", "init"]
    for mod in ("mandelbrot", "julia", "test"):
        cubin = cu.DEFAULT_KERNELS_DIR / (mod + ".cubin")
        assert cubin.exists(), cubin
        elf = subprocess.run(["cuobjdump", "-elf", str(cubin)], capture_output=True, text=True).stdout
        assert "sm_100" in elf or "SM100" in elf.upper()
        for n in names:
            assert (".text." + n) in elf, (mod, n)
        assert "VISUALIZE_SAMPLE_COUNT" in elf
    assert "julia_c" in subprocess.run(["cuobjdump", "-elf", str(cu.DEFAULT_KERNELS_DIR / "julia.cubin")],
                                       capture_output=True, text=True).stdout
    assert "amplifier" in subprocess.run(["cuobjdump", "-elf", str(cu.DEFAULT_KERNELS_DIR / "test.cubin")],
                                         capture_output=True, text=True).stdout


def test_model_mirror_matches_java_restatement(cu):
    m = cu.RenderingModel(canvasWidth=3840, canvasHeight=2160)
    m.setPlaneSegmentFromCenter(-0.748, 0.1, 2.0)
    assert m.planeSegment == oracle.segment_from_center(-0.748, 0.1, 2.0, 3840, 2160)
    assert m.planeSegment == cases.seg(-0.748, 0.1, 2.0, 3840, 2160)
    seg = list(m.planeSegment)
    for k in range(5):
        m.zoomAt((1920 + 37 * k, 1080 - 11 * k), into=(k % 2 == 0))
        seg = oracle.zoom_at(seg, 3840, 2160, (1920 + 37 * k, 1080 - 11 * k), k % 2 == 0)
        assert m.planeSegment == seg
    m.setMaxSuperSampling(1000)
    assert m.maxSuperSampling == 64.0
    m.setMaxSuperSampling(-3)
    assert m.maxSuperSampling == 0.0
    c = m.copy()
    c.planeSegment[0] = 5
    assert m.planeSegment[0] != 5


def test_cases_zoom_matches_oracle():
    seg = cases.seg(-0.5, 0.0, 2.0, 203, 117)
    assert cases.zoom_at(seg, 203, 117, (31, 90), True) == oracle.zoom_at(seg, 203, 117, (31, 90), True)
    assert cases.zoom_at(seg, 203, 117, (31, 90), False) == oracle.zoom_at(seg, 203, 117, (31, 90), False)


def test_default_palette_matches_oracle(cu):
    p = cu.createDefaultColorPalette()
    assert p.shape == (1536,) and p.dtype == np.uint32
    assert (p == oracle.default_palette()).all()
    assert p[0] == 0xFF7F0000 and (p >> 24 == 0xFF).all()   # (r=0,g=0,b=127), alpha 255, R in the low byte


def test_precision_rule():
    # "M ex 2" at 4K is FP64 by the reference's own rule; the full view is FP32
    assert oracle.choose_precision(cases.seg(-0.235125, 0.827215, 4.0e-5, 3840, 2160), 3840, 2160) == 1
    assert oracle.choose_precision(cases.seg(-0.5, 0.0, 2.0, 3840, 2160), 3840, 2160) == 0
    assert oracle.choose_precision(cases.seg(-0.551042868375875, 0.62714332109057, 8.00592947491907e-9, 8192, 8192), 8192, 8192) == 1
    assert oracle.choose_precision(cases.seg(-0.55, 0.62, 1e-14, 3840, 2160), 3840, 2160) == 2
