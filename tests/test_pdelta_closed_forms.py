"""P-Delta / imperfection pins (SURVEY 8c listed them as unpinned).

The numbers live in the golden fixture (tests/golden/make_golden.py, section 12b): a cantilever beam-column whose
expected displacements and reactions come from a self-contained plane solver inside the GENERATOR (its own element
matrices, its own fixed-point loop, sway-perturbed coordinates formed there -- nothing of fers_core_b200 or oracle/)
and from the classical formulas
    delta = H L^3 / 3EI * 3 (tan kL - kL) / (kL)^3,   k^2 = P / EI          (consistent K_g, 8 elements: 1e-5)
    delta = H / (3 EI / L^3 - P / L)                                          (simplified K_g, one element: exact)
    sway imperfection phi  ==  notional load phi P                            (to O(phi^2): 2e-4)
    P > P_cr = pi^2 EI / 4 L^2  ->  refused ("not positive definite")
test_oracle_golden.py / test_gpu_parity.py run every such case through the oracle and through calculate_from_json;
this file holds what does not fit the (path, value) form."""
import json

import pytest

from fers_core_b200 import wire
from oracle import fers_oracle, oracle_engine
from tests import helpers

PD_CASES = [n for n in helpers.golden_names() if n.startswith("pdelta_")]


def test_fixture_holds_the_pdelta_pins():
    assert set(PD_CASES) >= {"pdelta_beam_column_consistent", "pdelta_simplified_one_element", "pdelta_consistent_one_element",
                             "pdelta_sway_imperfection", "pdelta_notional_load", "pdelta_near_euler_load"}
    assert helpers.golden_error_names() == ["pdelta_above_euler_load"]
    for n in PD_CASES:
        assert helpers.golden(n)["expect"], n          # non-empty expectations


def _tip_dx(out):
    nodes = out["results"]["loadcombinations"]["C"]["displacement_nodes"]
    return nodes[str(max(int(k) for k in nodes))]["dx"]


def test_oracle_orders_simplified_consistent_exact():
    """One element, P = 0.3 P_cr: simplified ("P/L-only") < consistent (Przemieniecki) < beam-column formula."""
    ds = _tip_dx(oracle_engine.solve_document(helpers.golden("pdelta_simplified_one_element")["doc"]))
    dc = _tip_dx(oracle_engine.solve_document(helpers.golden("pdelta_consistent_one_element")["doc"]))
    lin_doc = json.loads(json.dumps(helpers.golden("pdelta_consistent_one_element")["doc"]))
    lin_doc["analysis"]["options"]["order"] = "LINEAR"
    dl = _tip_dx(oracle_engine.solve_document(lin_doc))
    assert dl < ds < dc
    assert ds / dl == pytest.approx(1.0 / (1.0 - 0.3 * (3.141592653589793 ** 2 / 4) / 3.0), rel=1e-9)


def test_oracle_sway_equals_notional_load():
    a = _tip_dx(oracle_engine.solve_document(helpers.golden("pdelta_sway_imperfection")["doc"]))
    b = _tip_dx(oracle_engine.solve_document(helpers.golden("pdelta_notional_load")["doc"]))
    assert a < 0.0 and a == pytest.approx(b, rel=2e-4)


def test_oracle_refuses_loads_above_the_euler_load():
    case = helpers.golden("pdelta_above_euler_load")
    with pytest.raises(fers_oracle.NotPositiveDefinite, match=case["expect_error"]):
        oracle_engine.solve_document(case["doc"])


def test_oracle_iteration_counts_are_reported():
    model = wire.parse_document(helpers.golden("pdelta_beam_column_consistent")["doc"])
    res = fers_oracle.analyze(model)
    assert res.converged.all() and (res.iterations >= 2).all() and (res.residual < 1e-12).all()


@pytest.mark.gpu
def test_engine_refuses_loads_above_the_euler_load():
    from fers_core_b200 import engine
    case = helpers.golden("pdelta_above_euler_load")
    with pytest.raises(engine.EngineError, match=case["expect_error"]):
        engine.calculate_from_json(json.dumps(case["doc"]))


@pytest.mark.gpu
def test_engine_sway_equals_notional_load_and_ordering():
    from fers_core_b200 import engine
    run = lambda n: _tip_dx(json.loads(engine.calculate_from_json(json.dumps(helpers.golden(n)["doc"]))))
    assert run("pdelta_sway_imperfection") == pytest.approx(run("pdelta_notional_load"), rel=2e-4)
    assert run("pdelta_simplified_one_element") < run("pdelta_consistent_one_element")
