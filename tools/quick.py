"""Quick A/B in ONE process (not the benchmark of record): for every setting (environment variables, joined by +) and
workload: a fresh renderer, warm-up frames, then the mean device-timed frame (chaos_stats.frame_ms) -- and the frame's
CRC-32, exact pixel-iteration and sample totals checked against tests/golden/workloads.json, so that a variant that
changes a result is seen at once.
usage: python tools/quick.py [--settings "X=0 CHAOS_KERNELS_DIR=tools/variants/a A=1+B=2"] [--workloads "c2 c4"] [--steps 10]"""
import argparse
import importlib
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

cu = importlib.import_module("chaos-ultra_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--settings", default="X=0")
    ap.add_argument("--workloads", default="c2 c2f32 c2ex2 c4 c5 c1")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    consts = json.loads((ROOT / "tests" / "golden" / "workloads.json").read_text())
    for setting in a.settings.split():
        env = dict(kv.split("=", 1) for kv in setting.split("+"))
        os.environ.update(env)
        prov = cu.CudaFractalRendererProvider(kernels_dir=env.get("CHAOS_KERNELS_DIR"))      # (diagnostic builds: tools/build_variant.sh)
        try:
            for name in a.workloads.split():
                wl = bench.WORKLOADS[name]
                r = prov.getRenderer(wl["fractal"], True)       # reopened: the knobs are read when a renderer is opened
                if wl["fractal"] == "julia":
                    r.setFractalCustomParams("%r;%r" % tuple(wl["julia_c"]))
                r.initializeRendering(wl["W"], wl["H"], None, cu.OUTPUT_DEVICE)
                m = bench.make_model(cu, wl)
                steps = max(3, a.steps // 3) if name == "c4" else a.steps
                ms, rms = [], []
                for f in range(a.warmup + steps):
                    r.renderQuality(m)
                    st = r.stats()
                    if f >= a.warmup:
                        ms.append(st.frame_ms)
                        rms.append(st.render_ms)
                crc = bench.frame_crc(r.outputRGBA())
                want = consts.get(name, {})
                ok = (crc == want.get("rgba_crc32") and st.pixel_iterations == want.get("pixel_iterations") and
                      st.samples == want.get("samples"))
                print("%-40s %-6s frame %8.3f ms (min %8.3f) render %8.3f  executed %.4g G  launches %d  %s" % (
                    setting, name, sum(ms) / len(ms), min(ms), sum(rms) / len(rms),
                    (st.pixel_iterations - st.skipped_iterations) / 1e9, st.kernel_launches, "ok" if ok else "WRONG FRAME"), flush=True)
                r.freeRenderingResources()
        finally:
            prov.close()
            for k in env:
                os.environ.pop(k, None)


if __name__ == "__main__":
    main()
