"""MemberHinge.rotational_release_warp (pydantic_models.py:2078): warping released at a member end.  On the NAFEMS LE5 cantilever
(Z section, warping restrained at the clamped end, torque at the tip) a released start of the first member takes the support's
warping restraint away: the bar twists uniformly, phi = T L / (G J) exactly (a uniform twist rate lies in the element's cubic
space), and no bimoment develops; without the release the restrained warping stiffens it."""
import copy

import numpy as np
import pytest

from fers_core_b200 import wire
from oracle import fers_oracle
from tests import helpers


def _doc(release):
    doc = copy.deepcopy(helpers.golden("nafems_LE5_4el")["doc"])
    if release is not None:
        doc["model"]["member_hinges"] = [{"id": 1, "hinge_type": "warping", "rotational_release_warp": release}]
        doc["model"]["members"][0]["start_hinge"] = 1
    return doc


def test_released_end_gets_its_own_warping_dof():
    m = wire.parse_document(_doc(0.0))
    assert m.n_nodes == 11 and m.w_aux.tolist() == [[10, 6], [6, 7], [7, 8], [8, 9]]
    assert m.bc_code[10].tolist() == [0, 1, 1, 1, 1, 1] and m.bc_code[5].tolist() == [1] * 6
    assert len(set(m.node_ids.tolist())) == m.n_nodes and (m.node_ids[5:] < 0).all()
    spring = wire.parse_document(_doc(5.0e3))
    assert spring.n_nodes == 10 and any(code == "W_WARP_RELEASE_SPRING" for code, _ in spring.warnings)


def test_uniform_torsion_when_the_restraint_is_released():
    free = wire.parse_document(_doc(0.0))
    held = wire.parse_document(_doc(None))
    sec, mat = free.member_raw(0)
    T, L = 1.2e6, float(free.xyz[4, 0] - free.xyz[0, 0])
    gj = float(mat["g_mod"]) * float(sec["j"])
    r_free, r_held = fers_oracle.analyze(free), fers_oracle.analyze(held)
    assert r_free.U[0, 4, 3] == pytest.approx(T * L / gj, rel=1e-9)
    assert np.abs(r_free.bw[0]).max() < 1e-6 * T
    assert 0.0 < r_held.U[0, 4, 3] < 0.9 * r_free.U[0, 4, 3]
    assert np.abs(r_held.bw[0]).max() > 1e-2 * T


@pytest.mark.gpu
def test_gpu_matches_oracle():
    from fers_core_b200 import engine
    m = wire.parse_document(_doc(0.0))
    helpers.assert_results_close(engine.analyze(m), fers_oracle.analyze(m), rtol=1e-9, label="warp release")
