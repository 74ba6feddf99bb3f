"""Helpers shared by the GPU parity tests: run one parity case through the C ABI."""
from __future__ import annotations

import numpy as np

import cases

FIELDS = ("value", "weight", "isReused", "weightOfNewSamples")


def model_for(cu, case, image=None, maxSS=None):
    m = cu.RenderingModel(canvasWidth=case["W"], canvasHeight=case["H"])
    m.planeSegment = list(case["image"] if image is None else image)
    m.maxIterations = case["maxIter"]
    m.maxSuperSampling = case["maxSS"] if maxSS is None else maxSS
    fl = case["flags"]
    m.useAdaptiveSuperSampling = bool(fl & cases.A)
    m.useFoveatedRendering = bool(fl & cases.FOV)
    m.useSampleReuse = bool(fl & cases.REUSE)
    m.zooming = bool(fl & cases.ZOOMING)
    m.zoomingIn = bool(fl & cases.ZOOM_IN)
    m.mouseFocus = tuple(case.get("focus", (0, 0)))
    m.forcePrecision = 2 if case["double"] else 1
    return m


def open_renderer(cu, provider, case, palette=None, mode=None):
    r = provider.getRenderer(cases.DISPLAY_NAME[case["fractal"]], False)
    if r.getState() == cu.STATE_READY_TO_RENDER:
        r.freeRenderingResources()
    r.initializeRendering(case["W"], case["H"], palette, cu.OUTPUT_HOST if mode is None else mode)
    if case["fractal"] == "julia":
        r.setFractalCustomParams("%r;%r" % tuple(case["julia_c"]))
    if case["fractal"] == "test":
        r.setFractalCustomParams(str(case["amplifier"]))
    if "params" in case:
        r.setFractalCustomParams(case["params"])
    return r


def setup_reference(rr, case):
    """write the module constants the Java host would write for this case into a reference module"""
    import numpy as np
    if case["fractal"] == "julia":
        rr.write_constant("julia_c", np.array(case["julia_c"], dtype=np.float64).tobytes())
    if case["fractal"] == "test":
        rr.write_constant("amplifier", np.array([case["amplifier"]], dtype=np.int32).tobytes())
    if "params" in case:
        roots, coefs, mag = cases.newton_constants(case["params"])
        rr.write_constant("roots", np.array(roots, dtype=np.float64).tobytes())
        rr.write_constant("coefficients", np.array(coefs, dtype=np.float64).tobytes())
        if mag is not None:
            rr.write_constant("colorMagnifier", np.array([mag], dtype=np.int32).tobytes())


def assert_records_equal(got, want, what=""):
    """bit-exact on every field the kernels write; the 3 padding bytes are not compared (the reference never
    writes them, helpers.cuh:106-130)"""
    assert got.shape == want.shape, what
    for f in FIELDS:
        a, b = got[f], want[f]
        if a.dtype.kind == "f":
            a = a.view(np.uint32)
            b = b.view(np.uint32)
        bad = a != b
        assert not bad.any(), "%s: field %s differs at %d of %d pixels, first %s: got %r want %r" % (
            what, f, int(bad.sum()), bad.size, tuple(np.argwhere(bad)[0]), got[f][bad][0], want[f][bad][0])
