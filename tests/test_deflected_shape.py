"""MemberResult.member_displacements (pydantic_models.py:1622-1625, option include_member_deflected_shape 1350-1353):
the sampled deflected shape of every member.  Pinned on the reference's own client-side Hermite reconstruction
(fers_core/fers/deformation_utils.py:19-114; fixture tests/golden/hermite_cases.json, made by
tests/golden/make_hermite_golden.py), then checked through the results document."""
import copy
import json
import os

import numpy as np
import pytest

from fers_core_b200 import hostlogic, results, wire
from oracle import fers_oracle
from tests import helpers
from tests.test_results_json import _same

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hermite_cases.json")


def _fixture():
    with open(FIXTURE) as fh:
        return json.load(fh)


def test_matches_the_reference_hermite_reconstruction():
    fx = _fixture()
    model = wire.parse_document(fx["doc"])
    assert fx["stations"] == hostlogic.DEFLECTED_STATIONS
    u = np.zeros((model.n_nodes, 6))
    for i, nid in enumerate(model.node_ids):
        u[i] = fx["node_displacements"][str(int(nid))]
    got = hostlogic.member_deflected_shape(model, u)
    assert got.shape == (model.n_members, hostlogic.DEFLECTED_STATIONS, 3)
    for m, mid in enumerate(model.m_ids):
        want = np.array(fx["member_displacements"][str(int(mid))])
        np.testing.assert_allclose(got[m], want, rtol=0, atol=1e-15 + 1e-13 * np.abs(want).max(), err_msg=f"member {mid}")


def test_end_stations_follow_the_rigid_arms_of_offset_members():
    doc = copy.deepcopy(_fixture()["doc"])
    member = doc["model"]["members"][3]
    member["start_offset"] = {"X": 0.1, "Y": -0.2, "Z": 0.05}
    model = wire.parse_document(doc)
    rng = np.random.default_rng(3)
    u = rng.normal(0, 0.01, (model.n_nodes, 6))
    got = hostlogic.member_deflected_shape(model, u)
    a = u[model.m_n1[3]]
    np.testing.assert_allclose(got[3, 0], a[:3] + np.cross(a[3:], [0.1, -0.2, 0.05]), atol=1e-15)
    np.testing.assert_allclose(got[3, -1], u[model.m_n2[3], :3], atol=1e-15)


def test_truss_members_stay_straight():
    doc = copy.deepcopy(_fixture()["doc"])
    doc["model"]["members"][0]["member_type"] = "Truss"
    model = wire.parse_document(doc)
    u = np.random.default_rng(4).normal(0, 0.01, (model.n_nodes, 6))
    got = hostlogic.member_deflected_shape(model, u)[0]
    a, b = u[model.m_n1[0], :3], u[model.m_n2[0], :3]
    xi = np.linspace(0, 1, hostlogic.DEFLECTED_STATIONS)[:, None]
    np.testing.assert_allclose(got, a * (1 - xi) + b * xi, atol=1e-15)


def _with_option(name, value=True):
    doc = copy.deepcopy(helpers.golden(name)["doc"])
    doc["analysis"]["options"]["include_member_deflected_shape"] = value
    return doc


@pytest.mark.parametrize("name", ["cantilever_end_load", "space_frame_3d", "portal_2d_combinations"])
def test_document_carries_the_polyline_only_when_asked(name):
    plain = wire.parse_document(helpers.golden(name)["doc"])
    res = fers_oracle.analyze(plain)
    for mr in results.results_to_wire(plain, res)["loadcases"].values():
        assert all("member_displacements" not in m for m in mr["member_results"].values())
    model = wire.parse_document(_with_option(name))
    assert model.options.include_member_deflected_shape
    res = fers_oracle.analyze(model)
    doc = results.results_to_wire(model, res)
    _same(json.loads(results.results_to_json(model, res)), json.loads(json.dumps(doc)))
    for part in ("loadcases", "loadcombinations"):
        for col in doc[part].values():
            for mid, mr in col["member_results"].items():
                samples = mr["member_displacements"]
                assert [s["x_frac"] for s in samples] == [i / 10 for i in range(11)] or np.allclose(
                    [s["x_frac"] for s in samples], np.linspace(0, 1, 11), atol=1e-15)
                m = int(np.nonzero(model.m_ids == int(mid))[0][0])
                for s, node in ((samples[0], model.m_n1[m]), (samples[-1], model.m_n2[m])):
                    dn = col["displacement_nodes"][str(int(model.node_ids[node]))]
                    # the consistency check of examples/104_visual_member_deflected_shape.py:85-93
                    np.testing.assert_allclose(s["displacement"], [dn["dx"], dn["dy"], dn["dz"]], atol=1e-15)


def test_cantilever_polyline_is_the_exact_cubic():
    """Tip load on a cantilever: no span load, so the Hermite cubic IS the elastic curve v = -P x^2 (3L - x) / 6EI."""
    model = wire.parse_document(_with_option("cantilever_end_load"))
    res = fers_oracle.analyze(model)
    mr = next(iter(results.results_to_wire(model, res)["loadcases"].values()))["member_results"]["1"]
    L, P, EI = 5.0, 1000.0, 210e9 * 13.21e-6
    for s in mr["member_displacements"]:
        x = s["x_frac"] * L
        assert s["displacement"][1] == pytest.approx(-P * x * x * (3 * L - x) / (6 * EI), rel=1e-5, abs=1e-9)


@pytest.mark.gpu
def test_gpu_document_matches_oracle_polyline():
    import fers_calculations
    doc = _with_option("space_frame_3d")
    got = json.loads(fers_calculations.calculate_from_json(json.dumps(doc)))["results"]
    model = wire.parse_document(doc)
    want = results.results_to_wire(model, fers_oracle.analyze(model))
    for part in ("loadcases", "loadcombinations"):
        for cname, col in want[part].items():
            for mid, mr in col["member_results"].items():
                g = np.array([s["displacement"] for s in got[part][cname]["member_results"][mid]["member_displacements"]])
                w = np.array([s["displacement"] for s in mr["member_displacements"]])
                np.testing.assert_allclose(g, w, rtol=1e-9, atol=1e-9 * np.abs(w).max() + 1e-300)


def test_lazy_view_builds_one_member_at_a_time():
    model = wire.parse_document(_with_option("space_frame_3d"))
    res = fers_oracle.analyze(model)
    eager = results.results_to_wire(model, res)
    lazy = results.lazy_bundle(model, res)
    for part in ("loadcases", "loadcombinations"):
        for cname, col in eager[part].items():
            for mid, mr in col["member_results"].items():
                assert lazy[part][cname]["member_results"][mid]["member_displacements"] == mr["member_displacements"]
