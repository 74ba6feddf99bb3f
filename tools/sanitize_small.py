"""compute-sanitizer target: small frames through every kernel family (streams incl. pass C and the pool, refill engine,
tile-synchronous, fast frame, Newton) -- `compute-sanitizer --tool memcheck python tools/sanitize_small.py`"""
import importlib, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import cases, helpers
cu = importlib.import_module("chaos-ultra_b200")
A = cases.A
frames = [
    dict(name="streams_a8", fractal="mandelbrot", W=333, H=130, image=cases.seg(-0.5, 0.0, 2.0, 333, 130), maxIter=2500, maxSS=8.0, flags=A, double=True, julia_c=(0, 0), amplifier=10),
    dict(name="streams_a4_f32", fractal="mandelbrot", W=203, H=117, image=cases.seg(-0.748, 0.1, 0.0014, 203, 117), maxIter=2100, maxSS=4.0, flags=A, double=False, julia_c=(0, 0), amplifier=10),
    dict(name="one_sample", fractal="mandelbrot", W=203, H=117, image=cases.seg(-0.5, 0.0, 2.0, 203, 117), maxIter=3000, maxSS=1.0, flags=0, double=True, julia_c=(0, 0), amplifier=10),
    dict(name="sync_a6", fractal="julia", W=160, H=96, image=cases.seg(0.0, 0.0, 4.0, 160, 96), maxIter=300, maxSS=6.0, flags=A, double=True, julia_c=(-0.4, 0.6), amplifier=10),
    dict(name="newton", fractal="newton_generic", W=120, H=72, image=cases.seg(0.0, 0.0, 4.0, 120, 72), maxIter=100, maxSS=3.0, flags=A, double=True, julia_c=(0, 0), amplifier=10, params=cases.N3),
]
with cu.CudaFractalRendererProvider() as prov:
    for eng in os.environ.get("SANITIZE_ENGINES", "3,2,1").split(","):
        os.environ["CHAOS_ENGINE"] = eng
        prov.getRenderer("test", False)
        for c in frames:
            r = helpers.open_renderer(cu, prov, c, mode=cu.OUTPUT_DEVICE)
            r.renderQuality(helpers.model_for(cu, c))
            print(eng, c["name"], r.stats().pixel_iterations, flush=True)
    os.environ.pop("CHAOS_ENGINE")
    if os.environ.get("SANITIZE_PARTS"):     # one-launch frames on their way to host memory, rendered in parts (4K: a part needs 60 000 tiles)
        for c in (dict(name="parts_sync", fractal="julia", W=3840, H=2160, image=cases.seg(0.0, 0.0, 4.0, 3840, 2160), maxIter=300, maxSS=4.0, flags=A, double=True, julia_c=(-0.4, 0.6), amplifier=10),
                  dict(name="parts_one_sample", fractal="mandelbrot", W=3840, H=2160, image=cases.seg(-0.5, 0.0, 2.0, 3840, 2160), maxIter=200, maxSS=1.0, flags=0, double=False, julia_c=(0, 0), amplifier=10)):
            r = helpers.open_renderer(cu, prov, c, mode=cu.OUTPUT_HOST)
            r.renderQuality(helpers.model_for(cu, c))
            print("parts", c["name"], r.stats().pixel_iterations, r.stats().kernel_launches, flush=True)
    case = cases.ADV_CASES[2]
    img0, img1 = cases.adv_segments(case)
    r = helpers.open_renderer(cu, prov, case)
    r.renderQuality(helpers.model_for(cu, case, image=img0, maxSS=case["maxSS0"]))
    r.renderFast(helpers.model_for(cu, case, image=img1))
    print("fast frame", r.stats().pixel_iterations)
print("sanitize ok")
