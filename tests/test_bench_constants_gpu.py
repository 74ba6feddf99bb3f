"""tests/golden/workloads.json pins, for every bench.py workload, the exact work count of a step and the checksum of the
composed frame.  bench.py asserts them on every run (every N), and its reference arm takes its work count from them, so
they are re-derived here on every GPU test run: this backend's frame and counters, and the reference kernels' frame."""
import importlib
import json
import sys
import zlib
from pathlib import Path

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
FILE = ROOT / "tests" / "golden" / "workloads.json"
sys.path.insert(0, str(ROOT))


def _constants():
    return json.loads(FILE.read_text()) if FILE.exists() else {}


def _crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


QUALITY = [n for n, c in _constants().items() if "rgba_crc32" in c]


@pytest.mark.skipif(not QUALITY, reason="tests/golden/workloads.json not generated yet")
@pytest.mark.parametrize("name", QUALITY)
def test_quality_workload_constants(cu, provider, name):
    import bench
    wl, want = bench.WORKLOADS[name], _constants()[name]
    r = provider.getRenderer(wl["fractal"], False)
    if wl["fractal"] == "julia":
        r.setFractalCustomParams("%r;%r" % tuple(wl["julia_c"]))
    if r.getState() == cu.STATE_READY_TO_RENDER:
        r.freeRenderingResources()
    r.initializeRendering(wl["W"], wl["H"], None, cu.OUTPUT_DEVICE)
    m = bench.make_model(cu, wl)
    r.renderQuality(m)
    st, rgba = r.stats(), r.outputRGBA()
    r.freeRenderingResources()
    assert st.pixel_iterations == want["pixel_iterations"] and st.samples == want["samples"]
    assert _crc(rgba) == want["rgba_crc32"] and int(np.bitwise_xor.reduce(rgba.ravel())) == want["rgba_xor"]
    if oracle.REFRUN_LIB.exists() and (oracle.REF_DIR / (wl["fractal"] + ".src.cubin")).exists():
        with oracle.RefRun(wl["fractal"], "src") as rr:
            if wl["fractal"] == "julia":
                rr.write_constant("julia_c", np.array(wl["julia_c"], dtype=np.float64).tobytes())
            _, _, _, ref = rr.frames(wl["W"], wl["H"], m.planeSegment, wl["maxIter"], wl["maxSS"], wl["flags"], oracle.default_palette(),
                                     wl["double"], 0, 1, True)
        assert _crc(ref) == want["rgba_crc32"], "the reference kernels' frame has another checksum"


@pytest.mark.skipif("c3" not in _constants(), reason="tests/golden/workloads.json not generated yet")
def test_zoom_workload_constants(cu, provider):
    import bench
    wl, want = bench.WORKLOADS["c3"], _constants()["c3"]
    n = len(want["frame_crc32"])
    assert n == 120
    segs = bench.zoom_segments(cu, wl, n)
    r = provider.getRenderer(wl["fractal"], False)
    if r.getState() == cu.STATE_READY_TO_RENDER:
        r.freeRenderingResources()
    r.initializeRendering(wl["W"], wl["H"], None, cu.OUTPUT_DEVICE)
    m = bench.zoom_model(cu, wl, segs[0]); m.maxSuperSampling = max(1.0, wl["maxSS"])
    r.renderQuality(m)
    assert _crc(r.outputRGBA()) == want["frame_crc32"][0] and r.stats().pixel_iterations == want["pixel_iterations"][0]
    for f in range(1, n):
        r.renderFast(bench.zoom_model(cu, wl, segs[f]))
        assert r.stats().pixel_iterations == want["pixel_iterations"][f], f
        if f in (1, 2, 30, 60, 90, 119):
            assert _crc(r.outputRGBA()) == want["frame_crc32"][f], "frame %d" % f
    r.freeRenderingResources()
