"""Shared helpers of the test-suite: golden fixtures, result-path lookup, array comparison."""
import gzip
import json
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "wire_cases.json.gz")
_cache = {}


def golden_cases() -> list:
    if "cases" not in _cache:
        with gzip.open(GOLDEN, "rt", encoding="utf-8") as fh:
            _cache["cases"] = json.load(fh)
    return _cache["cases"]


def golden_names(small_only: bool = False) -> list:
    """Cases that are expected to solve (cases that must be refused: golden_error_names)."""
    names = [c["name"] for c in golden_cases() if "expect_error" not in c]
    if small_only:
        names = [n for n in names if n not in ("nafems_LE1_72x8", "nafems_LE6_32x32")]
    return names


def golden_error_names() -> list:
    return [c["name"] for c in golden_cases() if "expect_error" in c]


def golden(name: str) -> dict:
    for c in golden_cases():
        if c["name"] == name:
            return c
    raise KeyError(name)


def lookup(results: dict, path: str):
    """'loadcases/NAME/displacement_nodes/2/dy' -> value ('__sum_reaction_fz' = derived)."""
    parts = path.split("/")
    cur = results
    for i, p in enumerate(parts):
        if p == "__sum_reaction_fz":
            return abs(sum(n["nodal_forces"]["fz"] for n in cur["reaction_nodes"].values()))
        if isinstance(cur, list) and "=" in p:          # "label=Bending z (6.2.5)": the record whose field has that value
            key, val = p.split("=", 1)
            cur = next(item for item in cur if str(item.get(key)) == val)
        elif isinstance(cur, list):
            cur = cur[int(p)]
        else:
            cur = cur[p]
    return cur


def nafems_value(results: dict, spec: dict) -> float:
    """Stress extraction of the reference's NAFEMS builders (tests/nafems/benchmarks/LE1_build.py:123-135,
    LE6_build.py:109-123): plate whose centroid is nearest the target point."""
    if spec["nafems"] == "LE5":
        # tests/nafems/benchmarks/LE5_build.py:70-100: warping stress from the bimoment at x = 2.5 m (end of member 1 /
        # start of member 2): sigma = 0.5 (|bw1| + |bw2|) omega_tip / C_w
        mr = results["loadcases"]["Torque"]["member_results"]
        bw = 0.5 * (abs(mr["1"]["end_node_forces"]["bw"]) + abs(mr["2"]["start_node_forces"]["bw"]))
        return bw * spec["omega"] / spec["cw"]
    case = "Edge" if spec["nafems"] == "LE1" else "P"
    pr = results["loadcases"][case]["plate_results"]
    best, bestd = None, 1e9
    for p in pr.values():
        dx = p["centroid"]["X"] - spec["point"][0]
        dy = p["centroid"]["Y"] - spec["point"][1]
        d = dx * dx + dy * dy
        if d < bestd:
            bestd, best = d, p
    r, t = best["resultants"], spec["thickness"]
    if spec["nafems"] == "LE1":
        return r["ny"] / t
    m1 = (r["mx"] + r["my"]) / 2 + math.sqrt(((r["mx"] - r["my"]) / 2) ** 2 + r["mxy"] ** 2)
    return 6.0 * m1 / (t * t)


def check_expectations(results: dict, expect: list) -> None:
    for ex in expect:
        if "nafems" in ex:
            got = nafems_value(results, ex)
        else:
            if ex["value"] is None:
                continue
            got = lookup(results, ex["path"])
        tol = ex["abs"] + ex.get("rel", 0.0) * abs(ex["value"])
        assert abs(got - ex["value"]) <= tol, (ex, got)


RESULT_FIELDS = ("U", "R", "f_local", "f_global", "u_local", "sec", "lmax", "lmin", "gmax", "gmin",
                 "p_res", "p_nodal", "p_cdg", "p_cdl", "bw")


# fields that are derived from the member end forces share their scale: an extremum or a section
# force that is analytically zero carries round-off of the size eps * |f_local|
_SCALE_OF = {"bw": "f_local", "sec": "f_local", "lmax": "f_local", "lmin": "f_local", "gmax": "f_local", "gmin": "f_local",
             "f_global": "f_local", "u_local": "U", "p_cdg": "U", "p_cdl": "U"}


def assert_results_close(got, want, rtol=1e-9, fields=RESULT_FIELDS, label=""):
    """Per result column: max |got - want| <= rtol * max |want| of that field (fp64 parity bar)."""
    for f in fields:
        a, b = getattr(got, f), getattr(want, f)
        assert a.shape == b.shape, (label, f, a.shape, b.shape)
        if a.size == 0:
            continue
        ref = getattr(want, _SCALE_OF.get(f, f))
        for col in range(a.shape[0]):
            scale = max(np.abs(b[col]).max(), np.abs(ref[col]).max() if ref.size else 0.0)
            err = np.abs(a[col] - b[col]).max()
            assert np.isfinite(a[col]).all(), (label, f, col, "non-finite")
            assert err <= rtol * scale + 1e-300, (label, f, col, err, scale, err / max(scale, 1e-300))
