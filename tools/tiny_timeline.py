"""Diagnostic: the launch timeline (CHAOS_TIMELINE) of a c2-like frame at a small size -- the long kernels then hold few orbits and their
duration is the latency of one orbit of maxIterations trips.  usage: python tools/tiny_timeline.py [W H] ..."""
import importlib, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["CHAOS_TIMELINE"] = str(ROOT / "gpurun_out" / "tiny_tl.txt")
import bench
cu = importlib.import_module("chaos-ultra_b200")
sizes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)] or [(256, 144), (640, 360)]
with cu.CudaFractalRendererProvider() as prov:
    for W, H in sizes:
        wl = dict(bench.WORKLOADS[os.environ.get("TINY_WL", "c2")], W=W, H=H)
        if "TINY_CENTER" in os.environ:      # (a view without pixels on the axes: x,y)
            wl["center"] = tuple(float(v) for v in os.environ["TINY_CENTER"].split(","))
        r = prov.getRenderer(wl["fractal"], True)
        if wl["fractal"] == "julia":
            r.setFractalCustomParams("%r;%r" % tuple(wl["julia_c"]))
        r.initializeRendering(W, H, None, cu.OUTPUT_HOST if os.environ.get("TINY_HOST") else cu.OUTPUT_DEVICE)
        if "TINY_PART" in os.environ:        # (what one rank of an N-GPU frame does: rank:world)
            pi, pn = map(int, os.environ["TINY_PART"].split(":"))
            r.setPartition(pi, pn, 32)
        m = bench.make_model(cu, wl)
        for _ in range(int(os.environ.get("TINY_FRAMES", "4"))):     # (many frames: the clocks are up by the last one)
            r.renderQuality(m)
        st = r.stats()
        print("== %dx%d frame %.3f ms, executed %.4g G, launches %d" % (W, H, st.frame_ms, (st.pixel_iterations - st.skipped_iterations) / 1e9, st.kernel_launches))
        print(open(os.environ["CHAOS_TIMELINE"]).read())
        r.freeRenderingResources()
