"""Lazy ResultsBundle view (fers_core_b200.results.lazy_bundle) against the eager wire document
(results_to_wire, contract fers_core/types/pydantic_models.py:2171-2185): same keys, same numbers, built on access."""
import json

import numpy as np
import pytest

from fers_core_b200 import results, wire
from oracle import fers_oracle
from tests import helpers

CASES = ["cantilever_end_load", "portal_2d_combinations", "hinge_spring", "plate_strip_surface_load", "nafems_LE5_4el"]


@pytest.mark.parametrize("name", CASES)
def test_lazy_view_equals_the_eager_document(name):
    model = wire.parse_document(helpers.golden(name)["doc"])
    res = fers_oracle.analyze(model)
    eager = results.results_to_wire(model, res)
    lazy = results.lazy_bundle(model, res)
    assert list(lazy) == list(eager)
    assert json.dumps(lazy.to_dict(), sort_keys=True) == json.dumps(eager, sort_keys=True)
    for group in ("loadcases", "loadcombinations"):
        assert list(lazy[group]) == list(eager[group]) and len(lazy[group]) == len(eager[group])
        for cname, col in eager[group].items():
            view = lazy[group][cname]
            assert isinstance(view["member_results"], results.LazyMap)
            for key in list(col["member_results"])[:3]:
                assert view["member_results"][key] == col["member_results"][key]
            for key in list(col["displacement_nodes"])[-2:]:
                assert view["displacement_nodes"][key] == col["displacement_nodes"][key]
            assert dict(view["reaction_nodes"].items()) == col["reaction_nodes"]
            assert view["summary"] == col["summary"] and view["solver_diagnostics"] == col["solver_diagnostics"]
            if "plate_results" in col:
                k = next(iter(col["plate_results"]))
                assert view["plate_results"][k] == col["plate_results"][k]


def test_lazy_view_builds_only_what_is_read(monkeypatch):
    model = wire.parse_document(helpers.golden("portal_2d_combinations")["doc"])
    res = fers_oracle.analyze(model)
    calls = []
    real = results._Column.member
    monkeypatch.setattr(results._Column, "member", lambda self, m: (calls.append((self.col, m)), real(self, m))[1])
    lazy = results.lazy_bundle(model, res)
    name = next(iter(lazy["loadcombinations"]))
    key = str(int(model.m_ids[1]))
    mz = lazy["loadcombinations"][name]["member_results"][key]["maximums"]["mz"]
    assert calls == [(res.plan.names.index(name), 1)] and np.isfinite(mz)
    with pytest.raises(KeyError):
        lazy["loadcombinations"][name]["member_results"]["no such member"]
    with pytest.raises(TypeError):
        lazy["loadcases"]["x"] = 1


@pytest.mark.gpu
def test_gpu_solve_lazy_matches_solve_document():
    import fers_calculations
    doc = helpers.golden("portal_2d_combinations")["doc"]
    eager = fers_calculations.solve_document(doc)["results"]
    lazy = fers_calculations.solve_lazy(doc)
    a, b = lazy.to_dict(), eager
    for grp in ("loadcases", "loadcombinations"):
        for cname in b[grp]:
            for part in ("displacement_nodes", "reaction_nodes", "member_results"):
                assert json.dumps(a[grp][cname][part], sort_keys=True) == json.dumps(b[grp][cname][part], sort_keys=True)
