"""Quick device-side sweep (not the benchmark of record): render kernel time per workload / engine / block size.
usage: python tools/sweep.py [workload ...] [--engines 0,1] [--nb 64,128,256] [--reps 3]"""
import argparse
import importlib
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

cu = importlib.import_module("chaos-ultra_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="*", default=["c1", "c2", "c2ex2", "c5"])
    ap.add_argument("--engines", default="0,1")
    ap.add_argument("--nb", default="128")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    prov = cu.CudaFractalRendererProvider()
    print("%-8s %3s %5s %10s %10s %14s %12s" % ("workload", "eng", "nb", "render_ms", "compose_ms", "pixel_iters", "Gpi/s"))
    for name in a.workloads:
        wl = bench.WORKLOADS[name]
        for eng in map(int, a.engines.split(",")):
            for nb in map(int, a.nb.split(",")):
                if eng == 0 and nb != int(a.nb.split(",")[0]):
                    continue
                os.environ["CHAOS_ENGINE"] = str(eng)
                os.environ["CHAOS_BLOCK_ITERS"] = str(nb)
                r = prov.getRenderer(wl["fractal"], True)
                if wl["fractal"] == "julia":
                    r.setFractalCustomParams("%r;%r" % tuple(wl["julia_c"]))
                r.initializeRendering(wl["W"], wl["H"], None, cu.OUTPUT_DEVICE)
                m = bench.make_model(cu, wl)
                best = None
                for _ in range(a.reps):
                    r.renderQuality(m)
                    st = r.stats()
                    if best is None or st.render_ms < best.render_ms:
                        best = st
                print("%-8s %3d %5d %10.3f %10.3f %14d %12.1f" % (name, eng, nb, best.render_ms, best.compose_ms, best.pixel_iterations,
                                                                best.pixel_iterations / best.render_ms / 1e6), flush=True)
                r.freeRenderingResources()
    prov.close()


if __name__ == "__main__":
    main()
