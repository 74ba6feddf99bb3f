"""Per-workload constants of bench.py's workloads (tests/golden/workloads.json): the exact work count of one step and
the checksum of the composed frame, so that

* bench.py --impl reference can state its throughput without loading this backend's library into that process,
* every bench run -- in particular every N > 1 run -- can ASSERT that the frame it timed is the right one.

Generated on a B200 (python tests/golden/make_workload_constants.py); for every workload the frame of this backend must
equal the frame of the reference's own kernels (oracle/_ref, src build) before anything is written.  The parity tests
(tests/test_bench_constants_gpu.py) re-check the committed values against both on every GPU test run."""
import importlib
import json
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
OUT = Path(__file__).resolve().parent / "workloads.json"


def frame_sums(rgba: np.ndarray) -> dict:
    flat = np.ascontiguousarray(rgba).ravel()
    return {"rgba_crc32": zlib.crc32(flat.tobytes()) & 0xFFFFFFFF, "rgba_xor": int(np.bitwise_xor.reduce(flat))}


def quality_constants(cu, bench, oracle, prov, name, wl, check_reference=True):
    r = prov.getRenderer(wl["fractal"], False)
    if wl["fractal"] == "julia":
        r.setFractalCustomParams("%r;%r" % tuple(wl["julia_c"]))
    if r.getState() == cu.STATE_READY_TO_RENDER:
        r.freeRenderingResources()
    r.initializeRendering(wl["W"], wl["H"], None, cu.OUTPUT_DEVICE)
    m = bench.make_model(cu, wl)
    r.renderQuality(m)
    st, rgba = r.stats(), r.outputRGBA()
    r.freeRenderingResources()
    if check_reference:
        with oracle.RefRun(wl["fractal"], "src") as rr:
            if wl["fractal"] == "julia":
                rr.write_constant("julia_c", np.array(wl["julia_c"], dtype=np.float64).tobytes())
            rec = rr.main(wl["W"], wl["H"], m.planeSegment, wl["maxIter"], wl["maxSS"], wl["flags"], wl["double"])
            want = rr.compose(rec, cu.createDefaultColorPalette(), wl["maxSS"], False)
        assert np.array_equal(rgba, want), name + ": frame differs from the reference kernels' frame"
        assert st.samples == int(rec["weight"].astype(np.uint64).sum()), name
    return dict(frame_sums(rgba), pixel_iterations=int(st.pixel_iterations), samples=int(st.samples))


def zoom_constants(cu, bench, oracle, prov, name, wl, frames=120):
    """crc of every frame of the zoom sequence (frame 0 quality, then fast frames) + the work of the fast frames"""
    W, H = wl["W"], wl["H"]
    segs = bench.zoom_segments(cu, wl, frames)
    r = prov.getRenderer(wl["fractal"], False)
    if r.getState() == cu.STATE_READY_TO_RENDER:
        r.freeRenderingResources()
    r.initializeRendering(W, H, None, cu.OUTPUT_DEVICE)
    m = bench.zoom_model(cu, wl, segs[0]); m.maxSuperSampling = max(1.0, wl["maxSS"])
    r.renderQuality(m)
    crcs, its = [frame_sums(r.outputRGBA())["rgba_crc32"]], [int(r.stats().pixel_iterations)]
    pal = cu.createDefaultColorPalette()
    with oracle.RefRun(wl["fractal"], "src") as rr:
        dbl = oracle.choose_precision(segs[0], W, H) != 0
        prev = rr.main(W, H, segs[0], wl["maxIter"], max(1.0, wl["maxSS"]), wl["flags"], dbl)
        for f in range(1, frames):
            r.renderFast(bench.zoom_model(cu, wl, segs[f]))
            rgba = r.outputRGBA()
            dbl = oracle.choose_precision(segs[f], W, H) != 0
            prev = rr.advanced(W, H, segs[f], wl["maxIter"], wl["maxSS"], wl["flags"], segs[f - 1], prev, wl["focus"], dbl)
            assert np.array_equal(rgba, rr.compose(prev, pal, wl["maxSS"], False)), "%s frame %d differs from the reference" % (name, f)
            crcs.append(frame_sums(rgba)["rgba_crc32"]); its.append(int(r.stats().pixel_iterations))
    r.freeRenderingResources()
    return {"frame_crc32": crcs, "pixel_iterations": its}


def main():
    cu = importlib.import_module("chaos-ultra_b200")
    import bench
    import oracle
    out = {}
    with cu.CudaFractalRendererProvider(device=0) as prov:
        for name, wl in bench.WORKLOADS.items():
            out[name] = zoom_constants(cu, bench, oracle, prov, name, wl) if wl.get("kind") == "zoom" else quality_constants(cu, bench, oracle, prov, name, wl)
            print(name, {k: (v if not isinstance(v, list) else "[%d values]" % len(v)) for k, v in out[name].items()}, flush=True)
    OUT.write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
