"""Produces the committed fixtures under tests/golden/ by running the REFERENCE's own kernels.

Run on a B200 box (the reference has no CPU path):
    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'
then copy gpurun_out/golden/*.npz into tests/golden/.  The modules it runs are oracle/_ref/<name>.src.cubin,
built by oracle/Makefile from /root/reference/src/main/cuda (nvcc 12.9, sm_100a), launched by oracle/refrun.cpp
with the Java host's launch geometry.  Each fixture stores the inputs (case dict) and, per field, the records the
reference wrote, plus the records of the shipped CUDA-9.2 PTX ("ptx92") for the report of how the two reference
builds differ from each other.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import cases  # noqa: E402
import oracle  # noqa: E402

import helpers  # noqa: E402

FILE_OF = {k: k for k in cases.DISPLAY_NAME}


def pack(rec):
    return dict(value=rec["value"].copy(), weight=rec["weight"].copy(), isReused=rec["isReused"].copy(),
                wnew=rec["weightOfNewSamples"].copy())


def setup(rr, case):
    helpers.setup_reference(rr, case)


def main(outdir):
    out = Path(outdir)
    out.mkdir(parents=True, exist_ok=True)
    pal = oracle.default_palette()
    report = {}
    for kind in ("src", "ptx92"):
        for case in cases.MAIN_CASES + cases.EXTRA_MAIN_CASES:
            with oracle.RefRun(FILE_OF[case["fractal"]], kind) as rr:
                setup(rr, case)
                rec = rr.main(case["W"], case["H"], case["image"], case["maxIter"], case["maxSS"], case["flags"], case["double"])
                rec2 = rr.main(case["W"], case["H"], case["image"], case["maxIter"], case["maxSS"], case["flags"], case["double"])
                det = all((rec[f] == rec2[f]).all() for f in ("value", "weight", "isReused", "weightOfNewSamples"))
                rgba = rr.compose(rec, pal, case["maxSS"], False)
                vis = rr.compose(rec, pal, case["maxSS"], True)
            d = {kind + "_" + k: v for k, v in pack(rec).items()}
            d[kind + "_rgba"] = rgba
            d[kind + "_vis"] = vis
            report.setdefault(case["name"], {})[kind + "_deterministic"] = bool(det)
            f = out / (case["name"] + ".npz")
            prev = dict(np.load(f)) if f.exists() else {}
            prev.update(d)
            prev["case_json"] = np.frombuffer(json.dumps(case).encode(), dtype=np.uint8)
            np.savez_compressed(f, **prev)
        for case in cases.ADV_CASES:
            img0, img1 = cases.adv_segments(case)
            with oracle.RefRun(FILE_OF[case["fractal"]], kind) as rr:
                setup(rr, case)
                rec0 = rr.main(case["W"], case["H"], img0, case["maxIter"], case["maxSS0"], case["flags"], case["double"])
                rec1 = rr.advanced(case["W"], case["H"], img1, case["maxIter"], case["maxSS"], case["flags"], img0, rec0,
                                   case["focus"], case["double"])
                rec1b = rr.advanced(case["W"], case["H"], img1, case["maxIter"], case["maxSS"], case["flags"], img0, rec0,
                                    case["focus"], case["double"])
                det = all((rec1[f] == rec1b[f]).all() for f in ("value", "weight", "isReused", "weightOfNewSamples"))
                rgba = rr.compose(rec1, pal, case["maxSS"], False)
                vis = rr.compose(rec1, pal, case["maxSS"], True)
            d = {kind + "_f0_" + k: v for k, v in pack(rec0).items()}
            d.update({kind + "_" + k: v for k, v in pack(rec1).items()})
            d[kind + "_rgba"] = rgba
            d[kind + "_vis"] = vis
            report.setdefault(case["name"], {})[kind + "_deterministic"] = bool(det)
            f = out / (case["name"] + ".npz")
            prev = dict(np.load(f)) if f.exists() else {}
            prev.update(d)
            prev["case_json"] = np.frombuffer(json.dumps(case).encode(), dtype=np.uint8)
            np.savez_compressed(f, **prev)
    (out / "REPORT.json").write_text(json.dumps(report, indent=1, sort_keys=True))
    print(json.dumps(report, sort_keys=True))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
