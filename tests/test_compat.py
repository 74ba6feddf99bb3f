"""B1, source level: a module file written for the REFERENCE's per-fractal contract (src/main/cuda/fractals/fractal.cuh:7-28:
computeFractal / colorize / debugFractal, built by compile.sh:10-15) compiles, unmodified, into a module of this backend
(chaos-ultra_b200/csrc/compat/) and gives the reference's records and colours.

* CPU: an author's file written here builds and exports every entry name and its constants; the reference's own seven
  files build where /root/reference is mounted (nothing is copied: the wrapper includes them by path).
* GPU: the modules built from the reference's unmodified files (cudaKernels_compat/, built in this container, shipped to
  the GPU box like every other built file) == the reference's own kernels run live == the native ports, bit for bit."""
import importlib
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

import cases
import helpers
import oracle

ROOT = Path(__file__).resolve().parent.parent
REF_FRACTALS = Path("/root/reference/src/main/cuda/fractals")
ENTRY_NAMES = ["fractalRenderMainFloat", "fractalRenderMainDouble", "fractalRenderAdvancedFloat", "fractalRenderAdvancedDouble",
               "compose", "fractalRenderUnderSampled", "debug", "init", "VISUALIZE_SAMPLE_COUNT"]

# a module as a fractal author would write it against the reference's contract: "burning ship", palette lookup by surf2Dread
AUTHOR_MODULE = r'''
#include "fractal.cuh"

__constant__ double power_shift[2];

template <class Real> __device__
float computeFractal(unsigned int maxIterations, Point<Real> c){
  Point<Real> z(0);
  unsigned int i = 0;
  while(i < maxIterations && z.x * z.x + z.y * z.y < 4){
    Real xn = z.x * z.x - z.y * z.y + c.x + (Real) power_shift[0];
    z.y = abs(2 * z.x * z.y) + c.y + (Real) power_shift[1];
    z.x = xn;
    ++i;
  }
  return i;
}

__device__ __forceinline__
unsigned int colorize(cudaSurfaceObject_t colorPalette, unsigned int paletteLength, float iterationResult){
  unsigned int iterationResult_i = round(iterationResult);
  unsigned int paletteIdx = paletteLength - (iterationResult_i % paletteLength) - 1;
  ASSERT(paletteIdx < paletteLength);
  unsigned int resultColor;
  surf2Dread(&resultColor, colorPalette, paletteIdx * sizeof(unsigned int), 0);
  return resultColor;
}

__device__ void debugFractal(){
  printf("hello from the burning ship\n");
}
'''


def _symbols(cubin: Path) -> str:
    return subprocess.run(["cuobjdump", "-elf", str(cubin)], capture_output=True, text=True, check=True).stdout


def _have_nvcc():
    return shutil.which("nvcc") is not None or Path("/usr/local/cuda/bin/nvcc").exists()


@pytest.mark.skipif(not _have_nvcc(), reason="nvcc not available")
def test_an_authors_reference_style_module_builds_unmodified(tmp_path):
    build = importlib.import_module("chaos-ultra_b200.build")
    (tmp_path / "fractals").mkdir()
    src = tmp_path / "fractals" / "burning_ship.cu"
    src.write_text(AUTHOR_MODULE)
    out = build.build_compat_module(src, tmp_path / "kernels", force=True)
    assert out.name == "burning_ship.cubin" and out.stat().st_size > 100000
    sym = _symbols(out)
    for name in ENTRY_NAMES + ["power_shift", "chaosProbeDouble", "chaosReusePassFloat", "CHAOS_MODULE_ABI_VERSION"]:
        assert name in sym, name
    assert src.read_text() == AUTHOR_MODULE                       # the author's file is included by path, never rewritten
    assert not list((tmp_path / "kernels").glob("tmp_compiling_*"))


@pytest.mark.skipif(not (_have_nvcc() and REF_FRACTALS.is_dir()), reason="needs nvcc and /root/reference")
def test_every_reference_module_builds_unmodified(tmp_path):
    build = importlib.import_module("chaos-ultra_b200.build")
    names = sorted(p.stem for p in REF_FRACTALS.glob("*.cu"))
    assert names == ["goc", "julia", "mandelbrot", "newton_generic", "newton_iterations", "newton_wired", "test"]
    for stem, consts in (("test", ["amplifier"]), ("newton_generic", ["roots", "coefficients"]), ("julia", ["julia_c"])):
        out = build.build_compat_module(REF_FRACTALS / (stem + ".cu"), tmp_path, fused_plane_y=stem not in build.UNFUSED_PLANE_Y, force=True)
        sym = _symbols(out)
        for name in ENTRY_NAMES + consts:
            assert name in sym, (stem, name)


# ---- GPU ---------------------------------------------------------------------------------------------------------------
COMPAT_DIR = ROOT / "chaos-ultra_b200" / "cudaKernels_compat"
COMPAT_CASES = [c for c in cases.MAIN_CASES + cases.EXTRA_MAIN_CASES
                if c["name"] in ("m_full_a8_f64", "m_ex1_a3p6_f32", "j_def_a2_f64", "t_amp10_f64", "t_amp3_a5_f32", "nw_a2_f32", "nw_a4_f64",
                                 "ng_n3_a3_f64", "ni_def_a2_f32", "goc_a2_f32", "goc_1s_f64")]


def _have_compat():
    return (COMPAT_DIR / "test.cubin").exists() and oracle.REFRUN_LIB.exists() and (oracle.REF_DIR / "test.src.cubin").exists()


@pytest.fixture(scope="module")
def compat_provider(cu):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    with cu.CudaFractalRendererProvider(kernels_dir=COMPAT_DIR) as p:
        yield p


@pytest.mark.gpu
@pytest.mark.skipif(not _have_compat(), reason="cudaKernels_compat/ or oracle/_ref not built")
@pytest.mark.parametrize("case", COMPAT_CASES, ids=[c["name"] for c in COMPAT_CASES])
def test_unmodified_reference_module_gives_the_reference_records(cu, compat_provider, provider, case):
    r = helpers.open_renderer(cu, compat_provider, case)
    r.renderQuality(helpers.model_for(cu, case))
    got, rgba = r.downloadRecords(), r.outputRGBA().copy()
    with oracle.RefRun(case["fractal"], "src") as rr:
        helpers.setup_reference(rr, case)
        want = rr.main(case["W"], case["H"], case["image"], case["maxIter"], case["maxSS"], case["flags"], case["double"])
        want_rgba = rr.compose(want, cu.createDefaultColorPalette(), case["maxSS"], False)
    helpers.assert_records_equal(got, want, case["name"] + ": compat module vs reference kernels")
    assert (rgba == want_rgba).all()
    # and the native port of the same module
    n = helpers.open_renderer(cu, provider, case)
    n.renderQuality(helpers.model_for(cu, case))
    helpers.assert_records_equal(got, n.downloadRecords(), case["name"] + ": compat module vs native port")
    assert (rgba == n.outputRGBA()).all()


@pytest.mark.gpu
@pytest.mark.skipif(not _have_compat(), reason="cudaKernels_compat/ or oracle/_ref not built")
def test_unmodified_reference_module_in_a_fast_frame_and_debug(cu, compat_provider, capfd):
    case = [c for c in cases.ADV_CASES if c["name"] == "adv_test_f32"][0]
    img0, img1 = cases.adv_segments(case)
    r = helpers.open_renderer(cu, compat_provider, case)
    r.renderQuality(helpers.model_for(cu, case, image=img0, maxSS=case["maxSS0"]))
    r.renderFast(helpers.model_for(cu, case, image=img1))
    got = r.downloadRecords()
    with oracle.RefRun(case["fractal"], "src") as rr:
        helpers.setup_reference(rr, case)
        rec0 = rr.main(case["W"], case["H"], img0, case["maxIter"], case["maxSS0"], case["flags"], case["double"])
        want = rr.advanced(case["W"], case["H"], img1, case["maxIter"], case["maxSS"], case["flags"], img0, rec0, case["focus"], case["double"])
    helpers.assert_records_equal(got, want, "compat test module, fast frame")
    r.launchDebugKernel()
    assert "hello from test" in capfd.readouterr().out


@pytest.mark.gpu
@pytest.mark.skipif(not _have_nvcc(), reason="nvcc not available")
def test_an_authors_module_registers_and_renders(cu, tmp_path):
    """the burning ship written against the reference's contract: built unmodified, registered under its own name, rendered;
    checked against the same loop in numpy (the author's expressions leave contraction to nvcc, so a handful of boundary
    pixels differ: that freedom is the module's, not the backend's)"""
    build = importlib.import_module("chaos-ultra_b200.build")
    (tmp_path / "fractals").mkdir()
    src = tmp_path / "fractals" / "burning_ship.cu"
    src.write_text(AUTHOR_MODULE)
    kd = tmp_path / "kernels"
    build.build_compat_module(src, kd, force=True)
    W, H, max_iter = 160, 120, 60
    with cu.CudaFractalRendererProvider(kernels_dir=kd) as prov:
        with pytest.raises(cu.IllegalArgumentException, match="Unknown fractal"):
            prov.getRenderer("burning ship", False)
        prov.registerModule("burning ship", "burning_ship")
        assert "burning ship" in prov.getAvailableFractals() and len(prov.getAvailableFractals()) == 8
        with pytest.raises(cu.IllegalArgumentException, match="already registered"):
            prov.registerModule("mandelbrot", "burning_ship")
        r = prov.getRenderer("burning ship", False)
        r.writeToConstantMemory("power_shift", np.array([0.0, 0.0], dtype=np.float64).tobytes())
        r.initializeRendering(W, H)
        m = cu.RenderingModel(canvasWidth=W, canvasHeight=H)
        m.setPlaneSegmentFromCenter(-0.5, -0.5, 3.0)
        m.maxIterations, m.maxSuperSampling, m.useAdaptiveSuperSampling, m.forcePrecision = max_iter, 1.0, False, 2
        r.renderQuality(m)
        got = r.downloadRecords()["value"].astype(np.int64)
        lbx, lby, rtx, rty = m.planeSegment
        xs = lbx + (rtx - lbx) / W * np.arange(W)
        ys = rty - (rty - lby) / H * np.arange(H)
        cx, cy = np.meshgrid(xs, ys)
        zx, zy, it = np.zeros_like(cx), np.zeros_like(cy), np.zeros(cx.shape, dtype=np.int64)
        for _ in range(max_iter):
            live = zx * zx + zy * zy < 4
            xn = zx * zx - zy * zy + cx
            zy = np.where(live, np.abs(2 * zx * zy) + cy, zy)
            zx = np.where(live, xn, zx)
            it += live
        assert (got != it).mean() < 0.01      # (near the boundary one different rounding moves an escape by many trips)
        assert r.outputRGBA().shape == (H, W)
        r.launchDebugKernel()
