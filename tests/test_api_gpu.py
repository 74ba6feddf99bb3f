"""GPU: behaviour of the renderer/provider contract through the C ABI -- state machine, argument validation, module
loading by name, reload, constants by name, fast->quality fallback, precision rule, partitioned rendering.
Each check cites the reference behaviour it mirrors (paths under src/main/java/.../chaosultra/)."""
import shutil
import struct

import numpy as np
import pytest

import cases
import helpers
import oracle

pytestmark = pytest.mark.gpu

CASE = cases.MAIN_CASES[2]          # mandelbrot 203x117, adaptive, FP64


def test_available_fractals_are_the_reference_registry(cu, provider):
    # cudarenderer/CudaFractalRendererProvider.java:21-27
    assert sorted(provider.getAvailableFractals()) == sorted(
        ["julia", "mandelbrot", "newton wired", "newton generic", "newton colored by iterations", "test", "goc"])


def test_default_renderer_is_mandelbrot_and_is_reused(cu, provider):
    a = provider.getDefaultRenderer()                      # :43-45
    assert a.getFractalName() == "mandelbrot"
    b = provider.getRenderer("mandelbrot", False)          # same name, no reload -> the active renderer (:49-51)
    assert b is a
    c = provider.getRenderer("mandelbrot", True)           # forceReload -> old one closed, new one created
    assert c is not a and a._h is None
    with pytest.raises(cu.IllegalArgumentException, match="Unknown fractal"):
        provider.getRenderer("no such fractal", False)     # :56-58


def test_state_machine(cu, provider):
    r = provider.getRenderer("mandelbrot", True)
    m = helpers.model_for(cu, CASE)
    assert r.getState() == cu.STATE_NOT_INITIALIZED and r.getWidth() == 0
    with pytest.raises(cu.IllegalStateException, match="initialized first"):
        r.renderQuality(m)                                 # CudaFractalRenderer.java:188
    with pytest.raises(cu.IllegalStateException, match="initialized first"):
        r.renderFast(m)                                    # :160
    with pytest.raises(cu.IllegalStateException, match="Already free"):
        r.freeRenderingResources()                         # :104
    r.initializeRendering(CASE["W"], CASE["H"])
    assert r.getState() == cu.STATE_READY_TO_RENDER and (r.getWidth(), r.getHeight()) == (CASE["W"], CASE["H"])
    with pytest.raises(cu.IllegalStateException, match="Already initialized"):
        r.initializeRendering(CASE["W"], CASE["H"])        # :86
    r.renderQuality(m)
    r.freeRenderingResources()
    assert r.getState() == cu.STATE_NOT_INITIALIZED
    r.initializeRendering(64, 48)                          # re-initialisable after a resize (GLRenderer.java:264-268)
    assert (r.getWidth(), r.getHeight()) == (64, 48)


def test_argument_validation(cu, provider):
    r = helpers.open_renderer(cu, provider, CASE)
    m = helpers.model_for(cu, CASE)
    m.maxIterations = 0
    with pytest.raises(cu.IllegalArgumentException, match="maxIterations must be a positive number, but is : 0"):
        r.renderQuality(m)                                 # RenderingKernel.java:69
    m = helpers.model_for(cu, CASE)
    m.planeSegment[2] = float("nan")
    with pytest.raises(cu.IllegalArgumentException, match="right_top_x must be a finite float"):
        r.renderQuality(m)                                 # RenderingKernel.java:143-146
    m = helpers.model_for(cu, CASE)
    m.planeSegment[1] = float("inf")
    with pytest.raises(cu.IllegalArgumentException, match="left_bottom_y"):
        r.renderFast(m)
    m = helpers.model_for(cu, CASE)
    m.maxSuperSampling = 0.5
    with pytest.raises(cu.FractalRendererException, match="maxSuperSampling must be >= 1"):
        r.renderQuality(m)                                 # device assert fractalRendererGeneric.cu:174
    # the renderer is still usable after every rejected call
    r.renderQuality(helpers.model_for(cu, CASE))
    with pytest.raises(cu.IllegalArgumentException):
        r.freeRenderingResources() or r.initializeRendering(0, 10)
    r.initializeRendering(CASE["W"], CASE["H"])


def test_fast_falls_back_to_quality_when_there_is_nothing_to_reuse(cu, provider):
    case = cases.ADV_CASES[1]
    img0, img1 = cases.adv_segments(case)
    r = helpers.open_renderer(cu, provider, case)
    m0 = helpers.model_for(cu, case, image=img0, maxSS=case["maxSS0"])
    r.renderFast(m0)                                       # lastRendering == null -> renderQuality (:164-168)
    q = r.downloadRecords()
    assert not q["isReused"].any()
    m1 = helpers.model_for(cu, case, image=img1)
    r.renderFast(m1)
    assert r.downloadRecords()["isReused"].any()           # now a real fast frame
    assert r.stats().reuse_ms > 0
    m2 = helpers.model_for(cu, case, image=img1, maxSS=case["maxSS0"])
    m2.sampleReuseCacheDirty = True                        # dirty cache -> quality again, flag cleared (:205)
    r.renderFast(m2)
    assert not r.downloadRecords()["isReused"].any() and not m2.sampleReuseCacheDirty
    r.freeRenderingResources()
    r.initializeRendering(case["W"], case["H"])            # reallocation marks the primary buffer dirty (:45-49)
    r.renderFast(helpers.model_for(cu, case, image=img1, maxSS=case["maxSS0"]))
    assert not r.downloadRecords()["isReused"].any()


def test_precision_rule_is_reported_and_followed(cu, provider):
    W, H = 3840, 2160
    r = provider.getRenderer("mandelbrot", False)
    if r.getState() == cu.STATE_READY_TO_RENDER:
        r.freeRenderingResources()
    r.initializeRendering(W, H, None, cu.OUTPUT_DEVICE)
    for center, zoom in [((-0.5, 0.0), 2.0), ((-0.235125, 0.827215), 4.0e-5), ((-0.55, 0.62), 1e-14)]:
        m = cu.RenderingModel(canvasWidth=W, canvasHeight=H)
        m.setPlaneSegmentFromCenter(center[0], center[1], zoom)
        m.maxIterations, m.maxSuperSampling = 50, 1.0
        r.renderQuality(m)
        assert m.floatingPointPrecision == oracle.choose_precision(m.planeSegment, W, H)   # CudaFractalRenderer.java:409-419
    r.freeRenderingResources()


def test_constants_are_written_by_name_and_size_checked(cu, provider):
    r = provider.getRenderer("julia", True)
    r.writeToConstantMemory("julia_c", struct.pack("<dd", -0.8, 0.156))
    with pytest.raises(cu.IllegalArgumentException, match="allocated to size 16"):
        r.writeToConstantMemory("julia_c", b"\0" * 24)     # FractalRenderingModule.java:176-179
    with pytest.raises(cu.IllegalArgumentException, match="no constant named"):
        r.writeToConstantMemory("does_not_exist", b"\0" * 4)
    with pytest.raises(cu.IllegalArgumentException, match="NumberFormatException"):
        r.setFractalCustomParams("abc;0.6")                # Double.parseDouble
    with pytest.raises(cu.IllegalArgumentException):
        r.setFractalCustomParams("0.3")                    # vals[1] missing
    r.setFractalCustomParams(" -0.4 ; 0.6 ")               # tokens are trimmed by Double.parseDouble
    g = provider.getRenderer("newton generic", False)
    with pytest.raises(cu.IllegalArgumentException, match="expecting 4 coefficients"):
        g.setFractalCustomParams('{"coefficients":[1,0,-1],"roots":[[1,0],[0,1],[0,-1]]}')   # ModuleNewtonGeneric.java:38-39
    with pytest.raises(cu.IllegalArgumentException, match="expecting 3 roots"):
        g.setFractalCustomParams('{"coefficients":[1,0,0,-1],"roots":[[1,0],[0,1]]}')
    with pytest.raises(cu.IllegalArgumentException, match="not represented as"):
        g.setFractalCustomParams('{"coefficients":[1,0,0,-1],"roots":[[1,0,3],[0,1],[0,-1]]}')
    g.setFractalCustomParams(cases.N3)


def test_default_values_of_every_module(cu, provider):
    want = {   # modules/Module*.java supplyDefaultValues
        "mandelbrot": dict(maxIterations=1600, maxSuperSampling=5.0, params="", seg=cases.seg(-0.5, 0.0, 2.0, 320, 180)),
        "julia": dict(maxIterations=900, params="-0.4;0.6"),
        "test": dict(params="10"),
        "newton wired": dict(maxIterations=200, params=""),
        "newton generic": dict(maxIterations=200, params=cases.N_DEFAULT),
        "newton colored by iterations": dict(maxIterations=200, params=cases.N_ITER),
        "goc": dict(maxIterations=900, params="", seg=cases.seg(1.1, -0.2, 0.20000000000000004, 320, 180)),
    }
    for name, w in want.items():
        r = provider.getRenderer(name, False)
        m = cu.RenderingModel(canvasWidth=320, canvasHeight=180)
        m.resetRenderingValuesToDefault()
        before = m.copy()
        r.supplyDefaultValues(m)
        assert m.fractalCustomParams == w["params"], name
        assert m.maxIterations == w.get("maxIterations", before.maxIterations), name
        assert m.maxSuperSampling == w.get("maxSuperSampling", before.maxSuperSampling), name
        assert m.planeSegment == w.get("seg", before.planeSegment), name
        r.setFractalCustomParams(m.fractalCustomParams) if m.fractalCustomParams else None   # GLRenderer.java:260-261


def test_modules_are_files_found_by_name_and_reload_rereads_the_file(cu, tmp_path):
    # FractalRenderingModule.java:63,73-88 ; README.md:82-86 "reload"
    for f in ("mandelbrot.cubin", "julia.cubin"):
        shutil.copy(cu.DEFAULT_KERNELS_DIR / f, tmp_path / f)
    case = cases.MAIN_CASES[0]
    with cu.CudaFractalRendererProvider(kernels_dir=tmp_path) as prov:
        with pytest.raises(cu.IllegalArgumentException, match="test.cubin"):
            prov.getRenderer("test", False)                # registered, but its file is not there: the message names the path
        r = prov.getRenderer("mandelbrot", False)
        r.initializeRendering(case["W"], case["H"])
        r.renderQuality(helpers.model_for(cu, case))
        inside_before = int((r.downloadRecords()["value"] == 0).sum())
        assert inside_before > 0                           # mandelbrot reports 0 for points that never escape
        shutil.copy(cu.DEFAULT_KERNELS_DIR / "julia.cubin", tmp_path / "mandelbrot.cubin")   # "recompile" the module
        r.renderQuality(helpers.model_for(cu, case))
        assert int((r.downloadRecords()["value"] == 0).sum()) == inside_before   # still the loaded module
        r2 = prov.getRenderer("mandelbrot", True)          # reload picks up the new file
        r2.initializeRendering(case["W"], case["H"])
        r2.renderQuality(helpers.model_for(cu, case))
        # the file now holds the julia code (its constant c is still 0): exactly what the oracle's julia gives
        want = oracle.render_main("julia", case["W"], case["H"], case["image"], case["maxIter"], case["maxSS"], case["flags"],
                                  case["double"], julia_c=(0.0, 0.0))
        helpers.assert_records_equal(r2.downloadRecords(), want.records, "reloaded module")
        assert int((want.records["value"] == 0).sum()) != inside_before
    with pytest.raises(cu.IllegalArgumentException, match="Invalid module file name"):
        with cu.CudaFractalRendererProvider(kernels_dir=tmp_path / "nowhere") as prov:
            prov.getRenderer("mandelbrot", False)


def test_corrupt_module_is_rejected(cu, tmp_path):
    (tmp_path / "mandelbrot.cubin").write_bytes(b"this is not a cubin")
    with cu.CudaFractalRendererProvider(kernels_dir=tmp_path) as prov:
        with pytest.raises(cu.CudaInitializationException):
            prov.getRenderer("mandelbrot", False)


def test_debug_kernel_runs(cu, provider, capfd):
    provider.getRenderer("mandelbrot", False).launchDebugKernel()   # CudaFractalRenderer.java:147-154
    assert "hello from mandelbrot" in capfd.readouterr().out


@pytest.mark.parametrize("world,band", [(2, 32), (3, 4), (4, 16), (8, 8)])
def test_partitioned_rendering_equals_whole_frame(cu, provider, world, band):
    """multi-GPU partition on one GPU: every part renders + composes only its row bands; their union is the frame"""
    case = cases.MAIN_CASES[2]
    r = helpers.open_renderer(cu, provider, case, mode=cu.OUTPUT_DEVICE)
    r.renderQuality(helpers.model_for(cu, case))
    whole_rec, whole_rgba, whole_it = r.downloadRecords(), r.outputRGBA(), r.stats().pixel_iterations
    part = __import__("importlib").import_module("chaos-ultra_b200.partition")
    rec = np.zeros_like(whole_rec)
    rgba = np.zeros_like(whole_rgba)
    total_it = 0
    for rank in range(world):
        r.freeRenderingResources()
        r.initializeRendering(case["W"], case["H"], None, cu.OUTPUT_DEVICE)
        r.setPartition(rank, world, band)
        r.renderQuality(helpers.model_for(cu, case))
        total_it += r.stats().pixel_iterations
        got_rec, got_rgba = r.downloadRecords(), r.outputRGBA()
        for r0, r1 in part.rows_owned(rank, world, case["H"], band):
            rec[r0:r1] = got_rec[r0:r1]
            rgba[r0:r1] = got_rgba[r0:r1]
    r.setPartition(0, 1, 32)
    helpers.assert_records_equal(rec, whole_rec, "union of %d partitions" % world)
    assert (rgba == whole_rgba).all() and total_it == whole_it
    with pytest.raises(cu.IllegalArgumentException, match="multiple of 4"):
        r.setPartition(0, 2, 6)
    with pytest.raises(cu.IllegalArgumentException):
        r.setPartition(2, 2, 8)


def test_partitions_compose_into_one_shared_target(cu, provider):
    """chaos_set_output_target: every part composes its bands straight into ONE frame (what the ranks do with rank 0's
    frame mapped over CUDA IPC); after the last part that frame is the whole picture, no gather step"""
    import torch
    case = cases.EXTRA_MAIN_CASES[0] if hasattr(cases, "EXTRA_MAIN_CASES") and cases.EXTRA_MAIN_CASES else cases.MAIN_CASES[2]
    case = dict(case, maxSS=8.0, flags=cases.A)            # several samples: passes A..D and both compose launches
    r = helpers.open_renderer(cu, provider, case, mode=cu.OUTPUT_DEVICE)
    r.renderQuality(helpers.model_for(cu, case))
    whole = r.outputRGBA()
    target = torch.zeros((case["H"], case["W"]), dtype=torch.int32, device="cuda")
    world, band = 3, 8
    for rank in range(world):
        r.freeRenderingResources()
        r.initializeRendering(case["W"], case["H"], None, cu.OUTPUT_DEVICE)
        r.setPartition(rank, world, band)
        r.setOutputTarget(target.data_ptr())
        r.renderQuality(helpers.model_for(cu, case))
    torch.cuda.synchronize()
    assert (target.cpu().numpy().view(np.uint32) == whole).all()
    r.setOutputTarget(0)
    r.setPartition(0, 1, 32)
    r.freeRenderingResources()
    r.initializeRendering(case["W"], case["H"], None, cu.OUTPUT_HOST)
    with pytest.raises(cu.IllegalStateException, match="DEVICE"):
        r.setOutputTarget(target.data_ptr())


def test_fast_frame_of_a_partitioned_renderer_renders_its_bands_afresh(cu, provider):
    """a rank that owns only its row bands has no valid previous frame around its pixels (the other bands were never
    rendered here), so chaos_render_fast must not reproject: it renders a quality frame of the own bands"""
    case = cases.ADV_CASES[2]
    img0, img1 = cases.adv_segments(case)
    part = __import__("importlib").import_module("chaos-ultra_b200.partition")
    r = helpers.open_renderer(cu, provider, case, mode=cu.OUTPUT_DEVICE)
    r.renderQuality(helpers.model_for(cu, case, image=img1, maxSS=case["maxSS0"]))
    whole = r.downloadRecords()                            # quality frame of the new segment, unpartitioned
    for rank in range(2):
        r.freeRenderingResources()
        r.initializeRendering(case["W"], case["H"], None, cu.OUTPUT_DEVICE)
        r.setPartition(rank, 2, 8)
        r.renderQuality(helpers.model_for(cu, case, image=img0, maxSS=case["maxSS0"]))   # the cache is clean now ...
        r.renderFast(helpers.model_for(cu, case, image=img1, maxSS=case["maxSS0"]))      # ... and still nothing is reprojected
        got = r.downloadRecords()
        assert r.stats().reuse_ms == 0
        for r0, r1 in part.rows_owned(rank, 2, case["H"], 8):
            helpers.assert_records_equal(got[r0:r1], whole[r0:r1], "rank %d rows %d..%d" % (rank, r0, r1))
            assert not got["isReused"][r0:r1].any()
    r.setPartition(0, 1, 32)


def test_custom_parameter_texts_are_validated_like_the_java_parsers(cu, provider):
    g = provider.getRenderer("newton generic", False)
    for bad in ('{"coefficients":["a",null,true,1],"roots":[[1,0],[0,1],[0,-1]]}',       # Gson getAsDouble throws
                '{"coefficients":[1,0,0,-1],"roots":[[1,"x"],[0,1],[0,-1]]}',
                '"just a string"', "[" * 5000):                                           # nesting limit, no stack overflow
        with pytest.raises(cu.IllegalArgumentException):
            g.setFractalCustomParams(bad)
    g.setFractalCustomParams(cases.N3)
    n = provider.getRenderer("newton colored by iterations", False)
    with pytest.raises(cu.IllegalArgumentException, match="colorMagnifier"):
        n.setFractalCustomParams('{"colorMagnifier": "7",' + cases.N3[1:])
    t = provider.getRenderer("test", False)
    for bad in (" 7", "+", "7 ", "99999999999", "0x10", ""):                              # Integer.parseInt
        with pytest.raises(cu.IllegalArgumentException, match="NumberFormatException"):
            t.setFractalCustomParams(bad)
    t.setFractalCustomParams("-3")
    t.setFractalCustomParams("+12")


def test_failed_open_does_not_leave_a_dangling_renderer(cu, tmp_path):
    shutil.copy(cu.DEFAULT_KERNELS_DIR / "mandelbrot.cubin", tmp_path / "mandelbrot.cubin")
    with cu.CudaFractalRendererProvider(kernels_dir=tmp_path) as prov:
        a = prov.getRenderer("mandelbrot", False)
        with pytest.raises(cu.IllegalArgumentException, match="Unknown fractal"):
            prov.getRenderer("nope", False)
        assert a._h is not None and a.getFractalName() == "mandelbrot"     # an unknown name closes nothing
        with pytest.raises(cu.IllegalArgumentException, match="julia.cubin"):
            prov.getRenderer("julia", False)                                # registered, file missing: the old one is closed first (:52)
        assert a._h is None and prov._active is None
        b = prov.getRenderer("mandelbrot", False)                           # and the provider still works
        assert b.getFractalName() == "mandelbrot"


def test_frame_driver_zoom_session_on_the_gpu(cu, provider, tmp_path):
    """the reference's closed loop (FSM + automatic quality) around the real renderer; frames end up as PNG"""
    import importlib
    drv = importlib.import_module("chaos-ultra_b200.driver")
    io = importlib.import_module("chaos-ultra_b200.imageio")
    W, H = 640, 360
    r = provider.getRenderer("mandelbrot", False)
    if r.getState() == cu.STATE_READY_TO_RENDER:
        r.freeRenderingResources()
    r.initializeRendering(W, H)
    m = cu.RenderingModel(canvasWidth=W, canvasHeight=H)
    m.resetRenderingValuesToDefault()
    r.supplyDefaultValues(m)
    d = drv.FrameDriver(r, m)                                 # native controller, frames timed on the device (CUDA events)
    h0 = m.planeSegment[3] - m.planeSegment[1]
    n = d.run_zoom_session((W // 2, H // 2), True, frames=20)
    assert n >= 21 and d.isWaiting()
    assert [k for _, k, _, _ in d.log[:20]] == ["fast"] * 20 and d.log[-1][1] == "quality"
    assert all(0.0 < ms < 1000.0 for _, _, _, ms in d.log)
    assert abs((m.planeSegment[3] - m.planeSegment[1]) / h0 - float(np.float32(0.977)) ** 20) < 1e-9
    assert 1.0 <= m.maxSuperSampling <= 64.0
    io.save_png(tmp_path / "last.png", r.outputRGBA())
    assert (io.decode_png_rgba8((tmp_path / "last.png").read_bytes()) == r.outputRGBA()).all()
    r.freeRenderingResources()
