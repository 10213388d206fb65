"""results.results_to_json (text written from the arrays by fers_b200_json_records) against the dict document of
results.results_to_wire -- the ``ResultsBundle`` contract, fers_core/types/pydantic_models.py:2171-2185."""
import json
import math

import numpy as np
import pytest

from fers_core_b200 import _lib, hostlogic, results, synth, wire
from oracle import fers_oracle
from tests import helpers


def _same(x, y, path=""):
    if isinstance(x, dict):
        assert isinstance(y, dict) and list(x) == list(y), path
        for k in x:
            _same(x[k], y[k], path + "/" + str(k))
    elif isinstance(x, list):
        assert isinstance(y, list) and len(x) == len(y), path
        for i, (p, q) in enumerate(zip(x, y)):
            _same(p, q, f"{path}[{i}]")
    elif isinstance(x, float) and math.isnan(x):
        assert isinstance(y, float) and math.isnan(y), path
    else:
        assert type(x) is type(y) and x == y, (path, x, y)


@pytest.mark.parametrize("name", helpers.golden_names(small_only=True))
def test_text_equals_the_dict_document_on_every_golden_case(name):
    model = wire.parse_document(helpers.golden(name)["doc"])
    res = fers_oracle.analyze(model)
    _same(json.loads(results.results_to_json(model, res)), json.loads(json.dumps(results.results_to_wire(model, res))))


def test_number_forms_survive_the_round_trip():
    m = synth.grid_frame(2, 2, 2, n_combos=2)
    plan = hostlogic.rhs_plan(m)
    res = results.ResultArrays.empty(m, plan)
    rng = np.random.default_rng(1)
    for f in ("U", "R", "f_local", "f_global", "u_local", "sec", "lmax", "lmin", "gmax", "gmin"):
        a = getattr(res, f)
        a[...] = rng.standard_normal(a.shape) * 10.0 ** rng.integers(-300, 300, a.shape)
    res.U[0, 3, 2] = 5.0; res.U[0, 4, 1] = float("nan"); res.R[1, 0, 0] = -0.0
    res.f_local[0, 7, 3] = float("inf"); res.f_local[0, 8, 3] = -float("inf"); res.f_global[1, 2, 0] = 1e22; res.lmax[0, 0, 0] = 123456789012345680.0
    res.method = "Linear"
    _same(json.loads(results.results_to_json(m, res)), json.loads(json.dumps(results.results_to_wire(m, res))))


def test_writer_rejects_bad_templates():
    import ctypes as C
    lib = _lib.load()
    out = np.empty(1024, np.uint8)
    w = C.c_size_t()
    vals = np.zeros((1, 2))
    assert lib.fers_b200_json_records(b"{\x01}", 3, 2, 1, b"1\0", vals.ctypes.data, out.ctypes.data, 1024, C.byref(w)) != 0
    assert b"slot" in lib.fers_b200_last_error()
    assert lib.fers_b200_json_records(b"[\x01,\x02]", 5, 2, 1, b"k\0", np.array([[1.5, 7.0]]).ctypes.data, out.ctypes.data, 1024, C.byref(w)) == 0
    assert out[: w.value].tobytes() == b'{"k":[1.5,7]}'
    assert lib.fers_b200_json_records(b"[\x01]", 3, 1, 1, b"k\0", np.array([[1.0]]).ctypes.data, out.ctypes.data, 8, C.byref(w)) != 0


@pytest.mark.gpu
def test_gpu_calculate_from_json_text_equals_solve_document():
    import fers_calculations
    for name in ("portal_2d_combinations", "plate_strip_surface_load", "nafems_LE5_4el"):
        doc = helpers.golden(name)["doc"]
        text = fers_calculations.calculate_from_json(json.dumps(doc))
        got = json.loads(text)
        want = json.loads(json.dumps(fers_calculations.solve_document(doc)))
        assert list(got) == list(want)
        for part in ("loadcases", "loadcombinations"):
            for cname, col in want["results"][part].items():
                for key in ("displacement_nodes", "reaction_nodes", "member_results", "summary"):
                    _same(got["results"][part][cname][key], col[key], f"{name}/{cname}/{key}")
