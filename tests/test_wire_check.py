"""The shared wire parser against a second, independent reading of the document (oracle/wire_check.py): every golden document
and the two building-sized example models.  ("What's weak" 2: the oracle solves on the parser's arrays.)"""
import gzip
import json
import math
import os

import numpy as np
import pytest

from fers_core_b200 import wire
from oracle import wire_check
from tests import helpers

with gzip.open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "example_models.json.gz"), "rt", encoding="utf-8") as fh:
    EXAMPLES = json.load(fh)
CASES = [(n, helpers.golden(n)["doc"]) for n in helpers.golden_names()] + sorted(EXAMPLES.items())


@pytest.mark.parametrize("name,doc", CASES, ids=[c[0] for c in CASES])
def test_parser_arrays_equal_the_second_reading(name, doc):
    m = wire.parse_document(doc)
    r = wire_check.read(doc)
    n = len(r["node_ids"])                                   # the parser may append auxiliary (warping) nodes after the real ones
    assert m.node_ids[:n].tolist() == r["node_ids"]
    assert np.array_equal(m.xyz[:n], np.array(r["xyz"], float).reshape(n, 3))
    assert m.bc_code[:n].tolist() == r["bc_code"]
    want_k = np.array(r["bc_k"], float).reshape(n, 6)
    known = ~np.isnan(want_k)
    assert np.array_equal(m.bc_k[:n][known], want_k[known])
    assert m.m_ids.tolist() == r["m_ids"] and m.m_n1.tolist() == r["m_n1"] and m.m_n2.tolist() == r["m_n2"]
    assert m.m_type.tolist() == r["m_type"]
    assert np.allclose(m.m_roll, r["m_roll"], rtol=0, atol=1e-15)
    for i, row in enumerate(r["m_stiffness"]):
        if row is None:
            assert m.m_prop[i] < 0
            continue
        got = m.prop[m.m_prop[i], :6]
        for c, v in enumerate(row):
            if not math.isnan(v):
                assert got[c] == pytest.approx(v, rel=1e-15), (name, i, c)
    for i, row in enumerate(r["m_release"]):
        if row is None or m.m_type[i] == wire.MT_CABLE:
            continue
        assert m.m_rel[i] >= 0
        got = m.rel[m.m_rel[i], :12]
        for c, v in enumerate(row):
            if not math.isnan(v):
                assert got[c] == v, (name, i, c)
    if not (doc["model"].get("plate_surfaces") and not doc["model"].get("plate_elements")):
        assert m.n_plates == len(r["plates"])
        for e, (pid, tri, th, E, nu) in enumerate(r["plates"]):
            assert int(m.p_ids[e]) == pid and m.p_nodes[e].tolist() == tri
            assert m.p_data[e, 0] == th and m.p_data[e, 1] == E and m.p_data[e, 2] == pytest.approx(nu, rel=1e-15)
    cable = set(m.c_member.tolist())
    generated = len(m.load_cases) - len(r["load_cases"])     # a self-weight case the parser may have appended
    assert generated in (0, 1)
    for lc, want in zip(m.load_cases, r["load_cases"]):
        assert lc.id == want["id"]
        got_nodal = {}
        for i, v in zip(lc.nod_node.tolist(), lc.nod_vec):
            acc = got_nodal.setdefault(i, np.zeros(6))
            acc += v
        assert set(got_nodal) == set(want["nodal"])
        for i, v in want["nodal"].items():
            assert np.allclose(got_nodal[i], v, rtol=1e-15, atol=0.0)
        own = [(mi, row) for mi, row in want["distributed"] if mi not in cable]
        got_rows = [(int(mi), row.tolist()) for mi, row in zip(lc.dl_member, lc.dl_data)]
        if not lc.pp_plate.size and len(got_rows) == len(own):        # (surface loads may add member rows: own tests)
            for (gm, grow), (wm, wrow) in zip(got_rows, own):
                assert gm == wm and np.allclose(grow, wrow, rtol=1e-15, atol=0.0)
    assert m.combo_ids.tolist() == [c[0] for c in r["combinations"]]
    for row, (_cid, want) in zip(m.combo_factors, r["combinations"]):
        assert np.allclose(row[: len(want)], want, rtol=1e-15, atol=0.0)
