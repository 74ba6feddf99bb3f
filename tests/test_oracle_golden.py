"""CPU: pins the oracle (oracle/chaos_oracle.c) to what the reference's own kernels produced on a B200
(tests/golden/*.npz, made by tests/golden/make_golden.py).  Bit-exact on every written field, RGBA identical."""
import json
from pathlib import Path

import numpy as np
import pytest

import cases
import oracle

GOLDEN = Path(__file__).resolve().parent / "golden"
ALL = cases.MAIN_CASES + cases.EXTRA_MAIN_CASES + cases.ADV_CASES


def _load(case):
    f = GOLDEN / (case["name"] + ".npz")
    assert f.exists(), "fixture missing: %s (see tests/golden/README.md)" % f
    g = np.load(f)
    stored = json.loads(bytes(g["case_json"]).decode())
    # the fixture was produced from exactly this case definition
    for k in ("W", "H", "maxIter", "maxSS", "flags", "double", "fractal", "image"):
        assert stored[k] == (list(case[k]) if isinstance(case[k], (list, tuple)) else case[k]), (case["name"], k)
    return g


def _eq(a, b):
    if a.dtype.kind == "f":
        a, b = a.view(np.uint32), b.view(np.uint32)
    return (a == b).all()


def _check(rec, g, prefix):
    assert _eq(rec["value"], g[prefix + "value"])
    assert _eq(rec["weight"], g[prefix + "weight"])
    assert _eq(rec["isReused"], g[prefix + "isReused"])
    assert _eq(rec["weightOfNewSamples"], g[prefix + "wnew"])


@pytest.mark.parametrize("case", cases.MAIN_CASES, ids=[c["name"] for c in cases.MAIN_CASES])
def test_oracle_main_equals_reference_kernels(case):
    g = _load(case)
    res = oracle.render_main(case["fractal"], case["W"], case["H"], case["image"], case["maxIter"], case["maxSS"], case["flags"],
                             case["double"], julia_c=case["julia_c"], amplifier=case["amplifier"])
    assert res.ub_reads == 0          # no case steps into the reference's out-of-bounds samples[] read
    _check(res.records, g, "src_")
    pal = oracle.default_palette()
    assert (oracle.compose(case["fractal"], res.records, pal, case["maxSS"], False) == g["src_rgba"]).all()
    assert (oracle.compose(case["fractal"], res.records, pal, case["maxSS"], True) == g["src_vis"]).all()


ADV_ORACLE = [c for c in cases.ADV_CASES if c["fractal"] in oracle.FRACTAL_KINDS]


@pytest.mark.parametrize("case", ADV_ORACLE, ids=[c["name"] for c in ADV_ORACLE])
def test_oracle_advanced_equals_reference_kernels(case):
    g = _load(case)
    img0, img1 = cases.adv_segments(case)
    r0 = oracle.render_main(case["fractal"], case["W"], case["H"], img0, case["maxIter"], case["maxSS0"], case["flags"],
                            case["double"], julia_c=case["julia_c"], amplifier=case["amplifier"])
    _check(r0.records, g, "src_f0_")
    r1 = oracle.render_advanced(case["fractal"], case["W"], case["H"], img1, case["maxIter"], case["maxSS"], case["flags"], img0,
                                r0.records, case["focus"], case["double"], julia_c=case["julia_c"], amplifier=case["amplifier"])
    _check(r1.records, g, "src_")
    pal = oracle.default_palette()
    assert (oracle.compose(case["fractal"], r1.records, pal, case["maxSS"], False) == g["src_rgba"]).all()
    assert (oracle.compose(case["fractal"], r1.records, pal, case["maxSS"], True) == g["src_vis"]).all()
    # every branch of the advanced kernel is exercised somewhere in the set
    assert r1.records.shape == (case["H"], case["W"])


def test_advanced_cases_cover_all_branches():
    seen = {"reuse_plain": 0, "reuse_resampled": 0, "fresh": 0}
    for case in cases.ADV_CASES:
        g = np.load(GOLDEN / (case["name"] + ".npz"))
        reused, wnew = g["src_isReused"] != 0, g["src_wnew"]
        seen["reuse_plain"] += int((reused & (wnew == 0)).sum())
        seen["reuse_resampled"] += int((reused & (wnew > 0)).sum())
        seen["fresh"] += int((~reused).sum())
    assert all(v > 1000 for v in seen.values()), seen


def test_report_reference_is_deterministic():
    rep = json.loads((GOLDEN / "REPORT.json").read_text())
    assert len(rep) == len(ALL)
    assert all(v["src_deterministic"] and v["ptx92_deterministic"] for v in rep.values())


def test_shipped_ptx_differs_only_slightly_from_source_build():
    """Informational bound on the difference between the two reference builds (c.y mapping: one FMA in the
    CUDA-9.2 PTX, MUL+SUB in the nvcc-12.9 build)."""
    worst = 0.0
    for case in cases.MAIN_CASES:
        g = np.load(GOLDEN / (case["name"] + ".npz"))
        diff = g["src_value"] != g["ptx92_value"]
        worst = max(worst, diff.mean())
    assert worst < 0.05
