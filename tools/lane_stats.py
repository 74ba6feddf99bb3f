"""Where the lanes of the escape loop go: render one frame of a bench workload with a module built with
-DCHAOS_LANE_STATS (tools/ls_kernels/, built by this script when nvcc is there) and print the counters.
usage: python tools/lane_stats.py [workload ...]     (diagnostics; not part of the product or the tests)"""
import importlib, os, shutil, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
KD = ROOT / "tools" / "ls_kernels"

def build():
    KD.mkdir(exist_ok=True)
    src = ROOT / "chaos-ultra_b200" / "csrc"
    for f in ("mandelbrot", "julia"):
        subprocess.run(["nvcc", "-cubin", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                        "-DCHAOS_LANE_STATS", *os.environ.get("LS_EXTRA", "").split(), "-I", str(src), str(src / "fractals" / (f + ".cu")), "-o", str(KD / (f + ".cubin"))], check=True)

if __name__ == "__main__":
    if sys.argv[1:2] == ["build"]:
        build(); sys.exit(0)
    frames = int(os.environ.get("LS_FRAMES", "2"))
    if frames <= 2:
        os.environ["CHAOS_LANE_STATS"] = "1"
    cu = importlib.import_module("chaos-ultra_b200")
    import bench
    for w in sys.argv[1:] or ["c2"]:
        wl = dict(bench.WORKLOADS[w.split("@")[0]])
        if "@" in w:        # c2@256x144: the workload's view at another size
            wl["W"], wl["H"] = map(int, w.split("@")[1].split("x"))
        with cu.CudaFractalRendererProvider(kernels_dir=os.environ.get("CHAOS_KERNELS_DIR", KD), device=0) as prov:
            r = prov.getRenderer(wl["fractal"], False)
            r.initializeRendering(wl["W"], wl["H"], output_mode=cu.OUTPUT_DEVICE)
            m = cu.RenderingModel(canvasWidth=wl["W"], canvasHeight=wl["H"])
            m.planeSegment = bench.seg(wl["center"][0], wl["center"][1], wl["zoom"], wl["W"], wl["H"])
            m.maxIterations, m.maxSuperSampling = wl["maxIter"], wl["maxSS"]
            m.useAdaptiveSuperSampling = bool(wl["flags"] & 1)
            m.forcePrecision = 2 if wl["double"] else 1
            for f in range(frames):
                if frames <= 2 or f % 50 == 0:
                    print("== %s frame %d" % (w, f), file=sys.stderr); sys.stderr.flush()
                r.renderQuality(m)
            st = r.stats()
            print(w, "render_ms %.3f pixel_iterations %d skipped %d" % (st.render_ms, st.pixel_iterations, st.skipped_iterations))
