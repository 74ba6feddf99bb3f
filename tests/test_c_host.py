"""The C ABI used from C (tests/c_abi_example.c), without Python in between: it must build against the header alone
(CPU) and produce the same frames as the Python mirror (GPU)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "c_abi_example.c"


def _build(tmp_path):
    exe = tmp_path / "c_abi_example"
    subprocess.run(["gcc", "-std=gnu11", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"), str(SRC), "-o", str(exe), "-ldl"], check=True)
    return exe


def test_c_host_builds_against_the_header_and_fails_loudly_without_a_device(cu, tmp_path):
    exe = _build(tmp_path)
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("has a device")
    except ImportError:
        pass
    res = subprocess.run([str(exe), str(cu.LIB_PATH), str(cu.DEFAULT_KERNELS_DIR)], capture_output=True, text=True)
    assert res.returncode == 1 and "chaos_provider_create" in res.stderr and "CUDA" in res.stderr.upper()


@pytest.mark.gpu
def test_c_host_renders_the_same_frames_as_the_python_mirror(cu, provider, tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([str(exe), str(cu.LIB_PATH), str(cu.DEFAULT_KERNELS_DIR)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    lines = dict((ln.split()[0], ln.split()[1:]) for ln in res.stdout.strip().splitlines())

    def checksum(frame):
        p = frame.ravel().astype(np.uint64)
        k = (np.uint64(2654435761) + np.arange(p.size, dtype=np.uint64)) & np.uint64(0xFFFFFFFF)
        return int(np.bitwise_xor.reduce((p * k) & np.uint64(0xFFFFFFFF)))

    W, H = 320, 180
    i = np.arange(512, dtype=np.uint32)
    palette = (np.uint32(0xFF000000) | (i & 0xFF) | (((i * 3) & 0xFF) << 8) | ((255 - (i >> 1)) << 16)).astype(np.uint32)
    r = provider.getRenderer("mandelbrot", True)
    r.initializeRendering(W, H, palette)
    m = cu.RenderingModel(canvasWidth=W, canvasHeight=H)
    r.supplyDefaultValues(m)
    m.useAdaptiveSuperSampling = m.useFoveatedRendering = m.useSampleReuse = True
    m.mouseFocus = (W // 2, H // 2)
    r.renderQuality(m)
    assert lines["quality"] == ["%08x" % checksum(r.outputRGBA()), str(r.stats().pixel_iterations), str(m.floatingPointPrecision)]
    m.zoomAt(m.mouseFocus, True)
    m.zooming = m.zoomingIn = True
    r.renderFast(m)
    assert lines["fast"] == ["%08x" % checksum(r.outputRGBA()), str(r.stats().pixel_iterations), str(m.floatingPointPrecision)]
    r.freeRenderingResources()
