"""GPU parity: this backend, called through the C ABI, against (1) the CPU oracle, (2) the reference's own
kernels run live on the same device (oracle/_ref via refrun), (3) the committed golden fixtures.
Bar (BASELINE.json): records bit-exact, RGBA identical (<= 1 LSB allowed, 0 observed)."""
import os
from pathlib import Path

import numpy as np
import pytest

import cases
import helpers
import oracle

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
ENGINES = [0, 1, 2]


def _ids(cs):
    return [c["name"] for c in cs]


def _have_ref():
    return oracle.REFRUN_LIB.exists() and (oracle.REF_DIR / "mandelbrot.src.cubin").exists()


@pytest.fixture(params=ENGINES, ids=lambda e: "engine%d" % e)
def engine(request, provider):
    os.environ["CHAOS_ENGINE"] = str(request.param)
    # the engine is read when a renderer is opened: force a reload
    for name in ("mandelbrot",):
        provider.getRenderer(name, True)
    yield request.param
    os.environ.pop("CHAOS_ENGINE", None)


@pytest.mark.parametrize("case", cases.MAIN_CASES, ids=_ids(cases.MAIN_CASES))
def test_main_matches_oracle(cu, provider, engine, case):
    r = helpers.open_renderer(cu, provider, case)
    m = helpers.model_for(cu, case)
    r.renderQuality(m)
    got = r.downloadRecords()
    want = oracle.render_main(case["fractal"], case["W"], case["H"], case["image"], case["maxIter"], case["maxSS"],
                              case["flags"], case["double"], julia_c=case["julia_c"], amplifier=case["amplifier"])
    helpers.assert_records_equal(got, want.records, case["name"])
    st = r.stats()
    assert st.pixel_iterations == want.pixel_iterations
    assert st.samples == want.samples
    # engines 0/1: one launch + compose; or passes A, classify, order, B + compose; or A, classify, order, B, compose, C, D, composeTiles.
    # engine 2: passes A and C (and the single launch) are probe -> long -> finish chains
    assert st.kernel_launches in ((4, 7, 13) if engine == 2 else (2, 5, 9))   # (+ chaosExportAll when rounds can be exported)
    # compose: palette lookup must be identical
    pal = cu.createDefaultColorPalette()
    assert (r.outputRGBA() == oracle.compose(case["fractal"], want.records, pal, case["maxSS"])).all()
    assert not m.sampleReuseCacheDirty


ALL_MAIN = cases.MAIN_CASES + cases.EXTRA_MAIN_CASES


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", ALL_MAIN, ids=_ids(ALL_MAIN))
def test_main_matches_live_reference(cu, provider, case):
    r = helpers.open_renderer(cu, provider, case)
    m = helpers.model_for(cu, case)
    r.renderQuality(m)
    got = r.downloadRecords()
    rgba = r.outputRGBA().copy()
    with oracle.RefRun(case["fractal"], "src") as rr:
        helpers.setup_reference(rr, case)
        want = rr.main(case["W"], case["H"], case["image"], case["maxIter"], case["maxSS"], case["flags"], case["double"])
        want_rgba = rr.compose(want, cu.createDefaultColorPalette(), case["maxSS"], False)
    helpers.assert_records_equal(got, want, case["name"] + " vs reference kernels")
    assert (rgba == want_rgba).all()


ADV_ORACLE = [c for c in cases.ADV_CASES if c["fractal"] in oracle.FRACTAL_KINDS]


@pytest.mark.parametrize("case", ADV_ORACLE, ids=_ids(ADV_ORACLE))
def test_advanced_matches_oracle(cu, provider, case):
    img0, img1 = cases.adv_segments(case)
    r = helpers.open_renderer(cu, provider, case)
    m0 = helpers.model_for(cu, case, image=img0, maxSS=case["maxSS0"])
    m0.sampleReuseCacheDirty = True
    r.renderFast(m0)                      # nothing to reuse yet -> falls back to a quality render
    assert not m0.sampleReuseCacheDirty
    rec0 = r.downloadRecords()
    want0 = oracle.render_main(case["fractal"], case["W"], case["H"], img0, case["maxIter"], case["maxSS0"],
                               case["flags"], case["double"], julia_c=case["julia_c"], amplifier=case["amplifier"])
    helpers.assert_records_equal(rec0, want0.records, case["name"] + " frame 0")
    m1 = helpers.model_for(cu, case, image=img1)
    r.renderFast(m1)
    got = r.downloadRecords()
    want = oracle.render_advanced(case["fractal"], case["W"], case["H"], img1, case["maxIter"], case["maxSS"], case["flags"],
                                  img0, want0.records, case["focus"], case["double"], julia_c=case["julia_c"],
                                  amplifier=case["amplifier"])
    helpers.assert_records_equal(got, want.records, case["name"] + " frame 1")
    st = r.stats()
    assert st.pixel_iterations == want.pixel_iterations
    assert st.samples == want.samples
    pal = cu.createDefaultColorPalette()
    assert (r.outputRGBA() == oracle.compose(case["fractal"], want.records, pal, case["maxSS"])).all()


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", cases.ADV_CASES, ids=_ids(cases.ADV_CASES))
def test_advanced_matches_live_reference(cu, provider, case):
    img0, img1 = cases.adv_segments(case)
    r = helpers.open_renderer(cu, provider, case)
    m0 = helpers.model_for(cu, case, image=img0, maxSS=case["maxSS0"])
    r.renderQuality(m0)
    m1 = helpers.model_for(cu, case, image=img1)
    r.renderFast(m1)
    got = r.downloadRecords()
    with oracle.RefRun(case["fractal"], "src") as rr:
        helpers.setup_reference(rr, case)
        rec0 = rr.main(case["W"], case["H"], img0, case["maxIter"], case["maxSS0"], case["flags"], case["double"])
        want = rr.advanced(case["W"], case["H"], img1, case["maxIter"], case["maxSS"], case["flags"], img0, rec0,
                           case["focus"], case["double"])
    helpers.assert_records_equal(got, want, case["name"] + " vs reference kernels")


def test_visualise_sample_count(cu, provider):
    case = cases.MAIN_CASES[2]
    r = helpers.open_renderer(cu, provider, case)
    m = helpers.model_for(cu, case)
    m.visualiseSampleCount = True
    r.renderQuality(m)
    want = oracle.render_main(case["fractal"], case["W"], case["H"], case["image"], case["maxIter"], case["maxSS"],
                              case["flags"], case["double"])
    pal = cu.createDefaultColorPalette()
    assert (r.outputRGBA() == oracle.compose(case["fractal"], want.records, pal, case["maxSS"], True)).all()
    m.visualiseSampleCount = False
    r.renderQuality(m)
    assert (r.outputRGBA() == oracle.compose(case["fractal"], want.records, pal, case["maxSS"], False)).all()


def test_device_output_mode(cu, provider):
    case = cases.MAIN_CASES[0]
    r = helpers.open_renderer(cu, provider, case, mode=cu.OUTPUT_DEVICE)
    m = helpers.model_for(cu, case)
    r.renderQuality(m)
    want = oracle.render_main(case["fractal"], case["W"], case["H"], case["image"], case["maxIter"], case["maxSS"],
                              case["flags"], case["double"])
    assert r.outputRGBADevicePointer() != 0
    assert (r.outputRGBA() == oracle.compose(case["fractal"], want.records, cu.createDefaultColorPalette(), 1.0)).all()


@pytest.mark.parametrize("case", ALL_MAIN + cases.ADV_CASES, ids=_ids(ALL_MAIN + cases.ADV_CASES))
def test_matches_golden_fixture(cu, provider, case):
    f = GOLDEN / (case["name"] + ".npz")
    if not f.exists():
        pytest.skip("fixture not generated yet")
    g = np.load(f)
    r = helpers.open_renderer(cu, provider, case)
    if "focus" in case:
        img0, img1 = cases.adv_segments(case)
        r.renderQuality(helpers.model_for(cu, case, image=img0, maxSS=case["maxSS0"]))
        r.renderFast(helpers.model_for(cu, case, image=img1))
    else:
        r.renderQuality(helpers.model_for(cu, case))
    got = r.downloadRecords()
    assert (got["value"].view(np.uint32) == g["src_value"].view(np.uint32)).all()
    assert (got["weight"].view(np.uint32) == g["src_weight"].view(np.uint32)).all()
    assert (got["isReused"] == g["src_isReused"]).all()
    assert (got["weightOfNewSamples"].view(np.uint32) == g["src_wnew"].view(np.uint32)).all()
    assert (r.outputRGBA() == g["src_rgba"]).all()


# ---- full-size, size-independent property: the two engines are different programs (engine 0: one warp per tile in
# ---- lock step, the reference's 7-operation trip; engine 1: lane refill + scaled 6-operation trip) and must agree
# ---- bit for bit on every record and on the exact pixel-iteration total, at sizes the CPU oracle cannot reach.
FULL_CASES = [
    dict(name="full_c2_4k_a8_f64", fractal="mandelbrot", W=3840, H=2160, image=cases.seg(-0.5, 0.0, 2.0, 3840, 2160), maxIter=2000,
         maxSS=8.0, flags=cases.A, double=True, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="full_c2ex2_4k_a8_f64", fractal="mandelbrot", W=3840, H=2160, image=cases.seg(-0.235125, 0.827215, 4.0e-5, 3840, 2160),
         maxIter=3000, maxSS=8.0, flags=cases.A, double=True, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="full_c4_2k_1s_f64", fractal="mandelbrot", W=2048, H=2048,
         image=cases.seg(-0.551042868375875, 0.62714332109057, 8.00592947491907e-9, 2048, 2048), maxIter=20000, maxSS=1.0, flags=0,
         double=True, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="full_c1_1k_1s_f64", fractal="mandelbrot", W=1024, H=1024, image=cases.seg(-0.5, 0.0, 2.0, 1024, 1024), maxIter=500,
         maxSS=1.0, flags=0, double=True, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="full_c5_4k_a8_f64", fractal="julia", W=3840, H=2160, image=cases.seg(0.0, 0.0, 4.0, 3840, 2160), maxIter=900, maxSS=8.0,
         flags=cases.A, double=True, julia_c=(-0.4, 0.6), amplifier=10),
    dict(name="full_c2_4k_a5_f32", fractal="mandelbrot", W=3840, H=2160, image=cases.seg(-0.5, 0.0, 2.0, 3840, 2160), maxIter=1000,
         maxSS=5.0, flags=cases.A, double=False, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="full_ragged_n3_f64", fractal="mandelbrot", W=1531, H=1077, image=cases.seg(-0.748, 0.1, 0.0014, 1531, 1077), maxIter=1600,
         maxSS=3.0, flags=0, double=True, julia_c=(0.0, 0.0), amplifier=10),
]


@pytest.mark.parametrize("case", FULL_CASES, ids=_ids(FULL_CASES))
def test_engines_agree_at_full_size(cu, provider, case):
    out = {}
    for eng in (0, 1, 2):
        os.environ["CHAOS_ENGINE"] = str(eng)
        try:
            provider.getRenderer("test", False)   # drop the active renderer so the engine choice is re-read
            r = helpers.open_renderer(cu, provider, case, mode=cu.OUTPUT_DEVICE)
            r.renderQuality(helpers.model_for(cu, case))
            st = r.stats()
            out[eng] = (r.downloadRecords(), st.pixel_iterations, st.samples, r.outputRGBA())
        finally:
            os.environ.pop("CHAOS_ENGINE", None)
    for eng in (1, 2):
        helpers.assert_records_equal(out[eng][0], out[0][0], case["name"] + " engine %d vs engine 0" % eng)
        assert out[eng][1] == out[0][1] and out[eng][2] == out[0][2]
        assert (out[eng][3] == out[0][3]).all()
    # sanity of the exact counter: every sample contributes between 0 and maxIter trips
    assert out[1][2] >= case["W"] * case["H"]
    assert out[1][1] <= out[1][2] * case["maxIter"]


# ---- BASELINE.json's own shapes against the reference's own kernels, run live on the same device ---------------------
# (the reference has no CPU path, but its kernels are fast enough on a B200 to serve as the full-size checker)
REF_FULL = [
    dict(name="ref_c1_1024_f64", fractal="mandelbrot", W=1024, H=1024, image=cases.seg(-0.5, 0.0, 2.0, 1024, 1024), maxIter=500,
         maxSS=1.0, flags=0, double=True, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="ref_c2_4k_a8_f64", fractal="mandelbrot", W=3840, H=2160, image=cases.seg(-0.5, 0.0, 2.0, 3840, 2160), maxIter=10000,
         maxSS=8.0, flags=cases.A, double=True, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="ref_c2ex2_4k_a8_f64", fractal="mandelbrot", W=3840, H=2160, image=cases.seg(-0.235125, 0.827215, 4.0e-5, 3840, 2160),
         maxIter=10000, maxSS=8.0, flags=cases.A, double=True, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="ref_c2_4k_a8_f32", fractal="mandelbrot", W=3840, H=2160, image=cases.seg(-0.5, 0.0, 2.0, 3840, 2160), maxIter=2500,
         maxSS=8.0, flags=cases.A, double=False, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="ref_c4_2k_1s_f64", fractal="mandelbrot", W=2048, H=2048,
         image=cases.seg(-0.551042868375875, 0.62714332109057, 8.00592947491907e-9, 2048, 2048), maxIter=20000, maxSS=1.0, flags=0,
         double=True, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="ref_c5_4k_a8_f64", fractal="julia", W=3840, H=2160, image=cases.seg(0.0, 0.0, 4.0, 3840, 2160), maxIter=900, maxSS=8.0,
         flags=cases.A, double=True, julia_c=(-0.4, 0.6), amplifier=10),
    dict(name="ref_julia_real_c_4k_f64", fractal="julia", W=3840, H=2160, image=cases.seg(0.0, 0.0, 0.68, 3840, 2160), maxIter=900, maxSS=4.0,
         flags=cases.A, double=True, julia_c=(-1.77578, 0.0), amplifier=10),   # c.im == 0: every orbit takes the 7-operation form
    dict(name="ref_newton_4k_f32", fractal="newton_generic", W=3840, H=2160, image=cases.seg(0.0, 0.0, 4.0, 3840, 2160), maxIter=100, maxSS=3.0,
         flags=cases.A, double=False, julia_c=(0.0, 0.0), amplifier=10, params=cases.N3),
]


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", REF_FULL, ids=_ids(REF_FULL))
def test_full_size_frame_matches_live_reference(cu, provider, case):
    r = helpers.open_renderer(cu, provider, case, mode=cu.OUTPUT_DEVICE)
    r.renderQuality(helpers.model_for(cu, case))
    got, rgba = r.downloadRecords(), r.outputRGBA()
    with oracle.RefRun(case["fractal"], "src") as rr:
        helpers.setup_reference(rr, case)
        want = rr.main(case["W"], case["H"], case["image"], case["maxIter"], case["maxSS"], case["flags"], case["double"])
        want_rgba = rr.compose(want, cu.createDefaultColorPalette(), case["maxSS"], False)
    helpers.assert_records_equal(got, want, case["name"])
    assert (rgba == want_rgba).all()


# ---- config c4 at BASELINE.json's full size: 8192 x 8192, maxIter 200000 (2 M vote tiles, 1 GiB of records, 10^11-trip
# ---- counters), one sample per pixel and adaptive maxSS 4 (every pass of the multi-sample chain at that size)
C4_FULL = [
    dict(name="ref_c4_8k_1s_f64", fractal="mandelbrot", W=8192, H=8192,
         image=cases.seg(-0.551042868375875, 0.62714332109057, 8.00592947491907e-9, 8192, 8192), maxIter=200000, maxSS=1.0, flags=0,
         double=True, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="ref_c4_8k_a4_f64", fractal="mandelbrot", W=8192, H=8192,
         image=cases.seg(-0.551042868375875, 0.62714332109057, 8.00592947491907e-9, 8192, 8192), maxIter=200000, maxSS=4.0, flags=cases.A,
         double=True, julia_c=(0.0, 0.0), amplifier=10),
]


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", C4_FULL, ids=_ids(C4_FULL))
def test_c4_full_size_matches_live_reference(cu, provider, case):
    r = helpers.open_renderer(cu, provider, case, mode=cu.OUTPUT_DEVICE)
    r.renderQuality(helpers.model_for(cu, case))
    st = r.stats()
    got, rgba = r.downloadRecords(), r.outputRGBA()
    r.freeRenderingResources()                       # 2 GiB of records back before the reference allocates its own
    with oracle.RefRun(case["fractal"], "src") as rr:
        want = rr.main(case["W"], case["H"], case["image"], case["maxIter"], case["maxSS"], case["flags"], case["double"])
        for f in helpers.FIELDS:                     # field by field: a structured 1 GiB compare would need 1 GiB temporaries per field anyway
            a, b = got[f], want[f]
            if a.dtype.kind == "f":
                a, b = a.view(np.uint32), b.view(np.uint32)
            assert np.array_equal(a, b), "%s: field %s differs at %d pixels" % (case["name"], f, int((a != b).sum()))
        samples = int(want["weight"].astype(np.uint64).sum())
        del got
        want_rgba = rr.compose(want, cu.createDefaultColorPalette(), case["maxSS"], False)
    assert np.array_equal(rgba, want_rgba)
    # the exact counters at this size: every sample is counted once, an orbit contributes 1 .. maxIter trips, and for one
    # sample per pixel the total follows from the records (inside points are stored as 0 and count maxIter)
    assert st.samples == samples
    assert st.samples <= st.pixel_iterations <= st.samples * case["maxIter"]
    if case["maxSS"] == 1.0:
        v = want["value"].astype(np.uint64)
        assert st.pixel_iterations == int(v.sum()) + int((v == 0).sum()) * case["maxIter"]


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("double", [False, True], ids=["f32", "f64"])
def test_zoom_sequence_4k_matches_live_reference(cu, provider, double):
    """config c3 at full length: frame 0 quality, then 119 fast frames.  Every frame's RGBA is compared with the
    reference's kernels fed with the reference's own previous frame (so errors cannot hide by accumulating
    identically); the records are compared bit for bit at frames 1, 2, 3, 4, 30, 60 and 119."""
    W, H, focus = 3840, 2160, (1920, 1080)
    flags = cases.A | cases.FOV | cases.REUSE | cases.ZOOMING | cases.ZOOM_IN
    base = dict(name="zoom4k", fractal="mandelbrot", W=W, H=H, maxIter=1600, maxSS=2.0, flags=flags, double=double,
                julia_c=(0.0, 0.0), amplifier=10, focus=focus)
    seg = cases.seg(-0.748, 0.1, 2.0 if not double else 1.0e-4, W, H)
    pal = cu.createDefaultColorPalette()
    r = helpers.open_renderer(cu, provider, dict(base, image=seg), mode=cu.OUTPUT_DEVICE)
    r.renderQuality(helpers.model_for(cu, dict(base, image=seg)))
    with oracle.RefRun("mandelbrot", "src") as rr:
        ref_prev = rr.main(W, H, seg, 1600, 2.0, flags, double)
        helpers.assert_records_equal(r.downloadRecords(), ref_prev, "frame 0")
        for f in range(1, 120):
            new_seg = cases.zoom_at(seg, W, H, focus, True)
            r.renderFast(helpers.model_for(cu, dict(base, image=new_seg)))
            want = rr.advanced(W, H, new_seg, 1600, 2.0, flags, seg, ref_prev, focus, double)
            assert np.array_equal(r.outputRGBA(), rr.compose(want, pal, 2.0, False)), "RGBA of frame %d" % f
            if f in (1, 2, 3, 4, 30, 60, 119):
                got = r.downloadRecords()
                helpers.assert_records_equal(got, want, "frame %d" % f)
                assert got["isReused"].mean() > 0.9 and (got["weightOfNewSamples"] > 0).sum() > 10000
            seg, ref_prev = new_seg, want


# ---- count-preserving shortcuts of the escape loop (quadratic.cuh items 3 and 4): every trip executed and tested
# ---- (CHAOS_SHORTCUTS=0) vs deferred test (1) vs deferred test + exact recurrence (3, the default) must give the same
# ---- records, the same reference-equivalent pixel-iteration total and the same RGBA, on frames dominated by
# ---- never-escaping orbits, odd iteration limits (tails shorter than a group), both precisions, both kernels.
SHORTCUT_CASES = [
    dict(name="sc_full_set_f64", fractal="mandelbrot", W=1920, H=1080, image=cases.seg(-0.5, 0.0, 2.0, 1920, 1080), maxIter=10007,
         maxSS=8.0, flags=cases.A, double=True, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="sc_full_set_f32", fractal="mandelbrot", W=1920, H=1080, image=cases.seg(-0.5, 0.0, 2.0, 1920, 1080), maxIter=5003,
         maxSS=4.0, flags=cases.A, double=False, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="sc_bulb_boundary_f64", fractal="mandelbrot", W=1280, H=720, image=cases.seg(-0.75, 0.05, 0.2, 1280, 720), maxIter=30000,
         maxSS=1.0, flags=0, double=True, julia_c=(0.0, 0.0), amplifier=10),    # parabolic point: slow, non-closing orbits
    dict(name="sc_tip_minus2_f64", fractal="mandelbrot", W=1280, H=720, image=cases.seg(-1.9, 0.0, 0.3, 1280, 720), maxIter=4001,
         maxSS=3.0, flags=cases.A, double=True, julia_c=(0.0, 0.0), amplifier=10),    # |c|^2 crosses 3.6 inside the frame
    dict(name="sc_julia_basilica_f64", fractal="julia", W=1920, H=1080, image=cases.seg(0.0, 0.0, 3.0, 1920, 1080), maxIter=5000,
         maxSS=4.0, flags=cases.A, double=True, julia_c=(-1.0, 0.0), amplifier=10),   # c.im == 0: unscaled 6-operation form
    dict(name="sc_julia_rabbit_f32", fractal="julia", W=1920, H=1080, image=cases.seg(0.0, 0.0, 3.0, 1920, 1080), maxIter=3001,
         maxSS=2.0, flags=cases.A, double=False, julia_c=(-0.123, 0.745), amplifier=10),
    dict(name="sc_sync_kernel_f64", fractal="mandelbrot", W=1920, H=1080, image=cases.seg(-0.5, 0.0, 2.0, 1920, 1080), maxIter=1001,
         maxSS=6.0, flags=cases.A, double=True, julia_c=(0.0, 0.0), amplifier=10),    # below the refill threshold: tile-synchronous kernel
    dict(name="sc_julia_far_c_f64", fractal="julia", W=640, H=360, image=cases.seg(0.0, 0.0, 5.0, 640, 360), maxIter=500,
         maxSS=2.0, flags=cases.A, double=True, julia_c=(1.5, 1.5), amplifier=10),    # |c|^2 > 3.6: shortcuts must stay off
]


@pytest.mark.parametrize("case", SHORTCUT_CASES, ids=_ids(SHORTCUT_CASES))
def test_shortcuts_change_nothing(cu, provider, case):
    out = {}
    for sc in (0, 1, 3):
        os.environ["CHAOS_SHORTCUTS"] = str(sc)
        try:
            provider.getRenderer("test", False)   # drop the active renderer so the knob is re-read
            r = helpers.open_renderer(cu, provider, case, mode=cu.OUTPUT_DEVICE)
            r.renderQuality(helpers.model_for(cu, case))
            st = r.stats()
            out[sc] = (r.downloadRecords(), st.pixel_iterations, st.samples, r.outputRGBA(), st.skipped_iterations)
        finally:
            os.environ.pop("CHAOS_SHORTCUTS", None)
    for sc in (1, 3):
        helpers.assert_records_equal(out[sc][0], out[0][0], "%s shortcuts %d vs 0" % (case["name"], sc))
        assert out[sc][1] == out[0][1] and out[sc][2] == out[0][2]
        assert (out[sc][3] == out[0][3]).all()
    assert out[0][4] == 0 and out[1][4] == 0
    assert out[3][4] <= out[3][1]
    if case["name"] == "sc_julia_far_c_f64":
        assert out[3][4] == 0
    elif case["name"] in ("sc_bulb_boundary_f64", "sc_tip_minus2_f64", "sc_sync_kernel_f64"):
        assert out[3][4] > 0
    else:
        assert out[3][4] > out[3][1] // 2, "most of these frames' work is never-escaping orbits that close exactly"


def test_shortcuts_in_fast_frames_change_nothing(cu, provider):
    """zoom steps with resampling in the fovea: pass S runs escape loops through the tile-synchronous sampler"""
    W, H, focus = 1920, 1080, (700, 500)
    flags = cases.A | cases.FOV | cases.REUSE | cases.ZOOMING | cases.ZOOM_IN
    base = dict(name="sczoom", fractal="mandelbrot", W=W, H=H, maxIter=3000, maxSS=4.0, flags=flags, double=True,
                julia_c=(0.0, 0.0), amplifier=10, focus=focus)
    frames = {}
    for sc in (0, 3):
        os.environ["CHAOS_SHORTCUTS"] = str(sc)
        try:
            provider.getRenderer("test", False)
            seg = cases.seg(-0.6, 0.0, 2.0, W, H)
            r = helpers.open_renderer(cu, provider, dict(base, image=seg), mode=cu.OUTPUT_DEVICE)
            r.renderQuality(helpers.model_for(cu, dict(base, image=seg)))
            got = []
            for _ in range(3):
                seg = cases.zoom_at(seg, W, H, focus, True)
                r.renderFast(helpers.model_for(cu, dict(base, image=seg)))
                st = r.stats()
                got.append((r.downloadRecords(), st.pixel_iterations, st.samples, st.skipped_iterations))
            frames[sc] = got
        finally:
            os.environ.pop("CHAOS_SHORTCUTS", None)
    for f, (a, b) in enumerate(zip(frames[3], frames[0])):
        helpers.assert_records_equal(a[0], b[0], "fast frame %d" % (f + 1))
        assert a[1] == b[1] and a[2] == b[2] and b[3] == 0
    assert frames[3][0][3] > 0


# ---- pass C/D (render_refill.cuh): tiles that leave pass B and get their remaining rounds as independent orbits must end
# ---- up with the records, counters and colours they would have had staying in pass B (CHAOS_EXPORT=0), whatever the
# ---- sample budget (3 = smallest that can export, 10 = largest, 11 and 64 = never exported), frame shape and module.
EXPORT_CASES = [
    dict(name="ex_full_set_a3_f64", fractal="mandelbrot", W=1920, H=1080, image=cases.seg(-0.5, 0.0, 2.0, 1920, 1080), maxIter=4000,
         maxSS=3.0, flags=cases.A, double=True, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="ex_full_set_a10_f64", fractal="mandelbrot", W=1531, H=1077, image=cases.seg(-0.5, 0.0, 2.0, 1531, 1077), maxIter=3000,
         maxSS=10.0, flags=cases.A, double=True, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="ex_full_set_a11_f64", fractal="mandelbrot", W=1280, H=720, image=cases.seg(-0.5, 0.0, 2.0, 1280, 720), maxIter=2500,
         maxSS=11.0, flags=cases.A, double=True, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="ex_noise_a8_f32", fractal="mandelbrot", W=1920, H=1080, image=cases.seg(-0.235125, 0.827215, 4.0e-4, 1920, 1080), maxIter=2500,
         maxSS=8.0, flags=cases.A, double=False, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="ex_fixed_n6_f64", fractal="mandelbrot", W=1280, H=720, image=cases.seg(-0.748, 0.1, 0.0014, 1280, 720), maxIter=2200,
         maxSS=6.0, flags=0, double=True, julia_c=(0.0, 0.0), amplifier=10),      # not adaptive: one decision, at i == S / 2
    dict(name="ex_julia_a8_f64", fractal="julia", W=1920, H=1080, image=cases.seg(0.0, 0.0, 3.2, 1920, 1080), maxIter=2500,
         maxSS=8.0, flags=cases.A, double=True, julia_c=(-0.8, 0.156), amplifier=10),
    dict(name="ex_a64_f64", fractal="mandelbrot", W=640, H=360, image=cases.seg(-0.5, 0.0, 2.0, 640, 360), maxIter=2100,
         maxSS=64.0, flags=cases.A, double=True, julia_c=(0.0, 0.0), amplifier=10),
]


@pytest.mark.parametrize("case", EXPORT_CASES, ids=_ids(EXPORT_CASES))
def test_exported_rounds_change_nothing(cu, provider, case):
    out = {}
    for ex in (0, 1):
        os.environ["CHAOS_EXPORT"] = str(ex)
        os.environ["CHAOS_STRANDS"] = "1"         # one pass chain: the launch counts below are per chain
        try:
            provider.getRenderer("test", False)   # drop the active renderer so the knob is re-read
            r = helpers.open_renderer(cu, provider, case, mode=cu.OUTPUT_DEVICE)
            r.renderQuality(helpers.model_for(cu, case))
            st = r.stats()
            out[ex] = (r.downloadRecords(), st.pixel_iterations, st.samples, r.outputRGBA(), st.skipped_iterations, st.kernel_launches)
        finally:
            os.environ.pop("CHAOS_EXPORT", None)
            os.environ.pop("CHAOS_STRANDS", None)
    helpers.assert_records_equal(out[1][0], out[0][0], case["name"] + " exported vs kept")
    assert out[1][1] == out[0][1] and out[1][2] == out[0][2]
    assert (out[1][3] == out[0][3]).all()
    assert out[0][5] == 7                         # probe, long, finish, classify, order, pass B, compose
    assert out[1][5] == (13 if 3 <= round(case["maxSS"]) <= 10 else 7)   # + export-all, compose of the final tiles, probe, long, finish, replay


# ---- strands (chaos_abi.cpp): a multi-pass frame cut into interleaved sets of row bands whose pass chains run next to
# ---- each other on their own streams must give the records, counters and colours of the single chain, for every strand
# ---- count, CTA size of the pass kernels, frame shape (ragged last band, fewer bands than strands) and module.
STRAND_CASES = [EXPORT_CASES[0], EXPORT_CASES[1], EXPORT_CASES[3], EXPORT_CASES[5], EXPORT_CASES[6],
                dict(name="st_small_a4_f64", fractal="mandelbrot", W=333, H=130, image=cases.seg(-0.5, 0.0, 2.0, 333, 130), maxIter=2100,
                     maxSS=4.0, flags=cases.A, double=True, julia_c=(0.0, 0.0), amplifier=10)]


@pytest.mark.parametrize("case", STRAND_CASES, ids=_ids(STRAND_CASES))
def test_strands_change_nothing(cu, provider, case):
    out = {}
    for key, (strands, threads) in {"one": (1, 256), "two": (2, 256), "three": (3, 128), "eight": (8, 64)}.items():
        os.environ["CHAOS_STRANDS"] = str(strands)
        os.environ["CHAOS_PASS_THREADS"] = str(threads)
        os.environ["CHAOS_STRAND_MIN_TILES"] = "0"        # (small frames would otherwise run as one chain)
        try:
            provider.getRenderer("test", False)   # drop the active renderer so the knobs are re-read
            r = helpers.open_renderer(cu, provider, case, mode=cu.OUTPUT_DEVICE)
            r.renderQuality(helpers.model_for(cu, case))
            st = r.stats()
            out[key] = (r.downloadRecords(), st.pixel_iterations, st.samples, r.outputRGBA(), st.kernel_launches)
        finally:
            os.environ.pop("CHAOS_STRANDS", None)
            os.environ.pop("CHAOS_PASS_THREADS", None)
            os.environ.pop("CHAOS_STRAND_MIN_TILES", None)
    for key in ("two", "three", "eight"):
        helpers.assert_records_equal(out[key][0], out["one"][0], case["name"] + " strands " + key)
        assert out[key][1] == out["one"][1] and out[key][2] == out["one"][2]
        assert (out[key][3] == out["one"][3]).all()
    if case["H"] >= 1000:
        assert out["two"][4] > out["one"][4]      # the frame really ran as several chains


# ---- engine 2's lists (render_streams.cuh): a survivor that finds the long list full, or an orbit that finds the finish
# ---- list full, is finished in place by the kernel that holds it. CHAOS_LIST_SHRINK=n gives the lists 1/n of their entries
# ---- so that these paths run; records, counters and colours must not notice.
@pytest.mark.parametrize("case", [EXPORT_CASES[0], EXPORT_CASES[5], SHORTCUT_CASES[0]],
                         ids=_ids([EXPORT_CASES[0], EXPORT_CASES[5], SHORTCUT_CASES[0]]))
def test_full_lists_change_nothing(cu, provider, case):
    out = {}
    for shrink in (1, 16, 4096):
        os.environ["CHAOS_LIST_SHRINK"] = str(shrink)
        os.environ["CHAOS_ENGINE"] = "2"
        try:
            provider.getRenderer("test", False)   # drop the active renderer so the knobs are re-read
            r = helpers.open_renderer(cu, provider, case, mode=cu.OUTPUT_DEVICE)
            r.renderQuality(helpers.model_for(cu, case))
            st = r.stats()
            out[shrink] = (r.downloadRecords(), st.pixel_iterations, st.samples, r.outputRGBA())
        finally:
            os.environ.pop("CHAOS_LIST_SHRINK", None)
            os.environ.pop("CHAOS_ENGINE", None)
    for shrink in (16, 4096):
        helpers.assert_records_equal(out[shrink][0], out[1][0], "%s lists / %d" % (case["name"], shrink))
        assert out[shrink][1] == out[1][1] and out[shrink][2] == out[1][2]
        assert (out[shrink][3] == out[1][3]).all()



# ---- one-launch frames on their way to host memory are rendered in parts (chaos_renderer::host_parts), each part composed
# ---- while the next one is rendered: frame, records and counters must be those of the single launch, for the
# ---- tile-synchronous kernel and for the lane-refill kernel, whole frames and one rank's bands.
PARTS_CASES = [
    dict(name="parts_julia_a8_f64", fractal="julia", W=3840, H=2160, image=cases.seg(0.0, 0.0, 4.0, 3840, 2160), maxIter=900,
         maxSS=8.0, flags=cases.A, double=True, julia_c=(-0.4, 0.6), amplifier=10),
    dict(name="parts_mandel_n1_f64", fractal="mandelbrot", W=3840, H=2161, image=cases.seg(-0.5, 0.0, 2.0, 3840, 2161), maxIter=700,
         maxSS=1.0, flags=0, double=True, julia_c=(0.0, 0.0), amplifier=10),
    dict(name="parts_mandel_n1_f32", fractal="mandelbrot", W=4093, H=2051, image=cases.seg(-0.7, 0.2, 1.0, 4093, 2051), maxIter=300,
         maxSS=1.0, flags=0, double=False, julia_c=(0.0, 0.0), amplifier=10),
]


@pytest.mark.parametrize("case", PARTS_CASES, ids=_ids(PARTS_CASES))
def test_one_launch_frames_in_parts_change_nothing(cu, provider, case):
    out = {}
    for parts, part in ((1, None), (2, None), (4, None), (3, (1, 2))):
        os.environ["CHAOS_HOST_PARTS"] = str(parts)
        try:
            provider.getRenderer("test", False)   # drop the active renderer so the knob is re-read
            r = helpers.open_renderer(cu, provider, case, mode=cu.OUTPUT_HOST)
            if part:
                r.setPartition(part[0], part[1], 32)
            r.renderQuality(helpers.model_for(cu, case))
            st = r.stats()
            out[(parts, part)] = (r.downloadRecords(), st.pixel_iterations, st.samples, r.outputRGBA().copy(), st.kernel_launches)
            if part:
                r.setPartition(0, 1, 32)
        finally:
            os.environ.pop("CHAOS_HOST_PARTS", None)
    whole = out[(1, None)]
    for key in ((2, None), (4, None)):
        helpers.assert_records_equal(out[key][0], whole[0], "%s in %d parts" % (case["name"], key[0]))
        assert out[key][1] == whole[1] and out[key][2] == whole[2]
        assert (out[key][3] == whole[3]).all()
        assert out[key][4] == 2 * key[0]          # a render and a compose launch per part
    # one rank of two, its bands in three parts: its rows equal the whole frame's
    rows = np.zeros(case["H"], dtype=bool)
    for b in range(1, (case["H"] + 31) // 32, 2):
        rows[b * 32:(b + 1) * 32] = True
    rank = out[(3, (1, 2))]
    helpers.assert_records_equal(rank[0][rows], whole[0][rows], case["name"] + " rank 1 of 2 in 3 parts")
    assert (rank[3][rows] == whole[3][rows]).all()
