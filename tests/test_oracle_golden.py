"""Pin the CPU oracle: every golden document (built by the reference's own front end, see
tests/golden/make_golden.py) must reproduce the expectations the reference's tests hold --
closed-form beam theory to rel 1e-5 / abs 1e-6 m, 1e-3 N.m, 1e-6 N (tests/common_functions.py:35-43)
and the NAFEMS known-answer numbers to their printed digit (LE1_result.json, LE6_result.json)."""
import numpy as np
import pytest

from fers_core_b200 import wire
from fers_core_b200.hostlogic import rhs_plan
from oracle import fers_oracle, oracle_engine
from tests import helpers


@pytest.mark.parametrize("name", helpers.golden_names())
def test_oracle_meets_reference_expectations(name):
    case = helpers.golden(name)
    out = oracle_engine.solve_document(case["doc"])
    helpers.check_expectations(out["results"], case["expect"])


@pytest.mark.parametrize("name", ["space_frame_3d", "portal_2d_combinations"])
def test_oracle_superposition_and_equilibrium(name):
    """Linear combinations equal the factored sum of their load cases (reference
    examples/091_Load_Combinations_Cantilever.py:99-112) and reactions balance the applied load."""
    model = wire.parse_document(helpers.golden(name)["doc"])
    res = fers_oracle.analyze(model)
    plan = rhs_plan(model)
    nlc = len(model.load_cases)
    for col in range(nlc, plan.n_rhs):
        want = np.einsum("l,lnd->nd", plan.factors[col], res.U[:nlc])
        assert np.abs(res.U[col] - want).max() <= 1e-12 * np.abs(want).max()
    asm = fers_oracle.Assembly(model)
    F, _ = fers_oracle.load_vectors(asm)
    for li in range(nlc):
        tot = F[li].reshape(-1, 6)[:, :3].sum(0) + res.R[li][:, :3].sum(0)
        assert np.abs(tot).max() <= 1e-8 * max(np.abs(F[li]).max(), 1.0)


def test_oracle_pattern_is_sorted_unique():
    model = wire.parse_document(helpers.golden("space_frame_3d")["doc"])
    rowptr, colidx = fers_oracle.bsr_pattern(model)
    assert rowptr[0] == 0 and rowptr[-1] == len(colidx)
    for i in range(model.n_nodes):
        row = colidx[rowptr[i]:rowptr[i + 1]]
        assert np.all(np.diff(row) > 0) and i in row
