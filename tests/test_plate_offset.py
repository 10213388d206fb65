"""Plate offset (PlateElement.offset, pydantic_models.py:1711-1714: "distance from the reference surface along the local plate
normal").  Exact identity: a plate whose nodes lie on the reference surface z = 0 and whose mid-surface is offset by o is the
plate with its nodes ON the mid-surface (z = o, no offset) seen through rigid arms: with the loads carried along the arms
(a force F at the lower point = F + moment (0, 0, -o) x F at the mid-surface) the displacements obey u_mid = u + o ry,
v_mid = v - o rx and the plate resultants are the same."""
import copy

import numpy as np
import pytest

from fers_core_b200 import wire
from oracle import fers_oracle
from tests.test_plate_orthotropic import _doc

O = 0.085


def _pair(ortho=False):
    base, _, nid = _doc("x", ortho=ortho, behavior="Shell", n=(5, 2))
    free = {"condition_type": "Free", "stiffness": None}
    fixed = {"condition_type": "Fixed", "stiffness": None}
    for s in base["model"]["nodal_supports"]:
        if s["id"] in (1, 2):
            s.update(X=fixed, Y=fixed, Z=fixed, RX=fixed, RY=fixed, RZ=fixed)        # clamped edge
        else:
            s.update(X=free, Y=free, Z=free, RX=free, RY=free, RZ=fixed)
    lc = base["analysis"]["load_cases"][0]
    far = [l["node"] for l in lc["nodal_loads"]]
    for k, node in enumerate(far):
        lc["nodal_loads"].append({"id": 100 + k, "node": node, "magnitude": 40.0 * (k + 1), "direction": [0, 0, -1], "load_type": "force"})
        lc["nodal_loads"].append({"id": 200 + k, "node": node, "magnitude": 300.0, "direction": [0, 1, 0], "load_type": "force"})
    a = copy.deepcopy(base)
    for p in a["model"]["plate_elements"]:
        p["offset"] = O
    b = copy.deepcopy(base)
    for n in b["model"]["nodes"]:
        n["Z"] = O
    moments = []
    for l in lc["nodal_loads"]:
        f = np.array(l["direction"], float) * l["magnitude"]
        mom = np.cross([0.0, 0.0, -O], f)                     # the force acts at the reference surface, o below the nodes of B
        for c in range(3):
            if mom[c] != 0.0:
                d = [0, 0, 0]; d[c] = 1
                moments.append({"id": 1000 + len(moments), "node": l["node"], "magnitude": float(mom[c]), "direction": d, "load_type": "moment"})
    b["analysis"]["load_cases"][0]["nodal_moments"] = moments
    return a, b


@pytest.mark.parametrize("ortho", [False, True])
def test_offset_plate_equals_the_plate_on_its_mid_surface_seen_through_rigid_arms(ortho):
    a, b = _pair(ortho)
    ma, mb = wire.parse_document(a), wire.parse_document(b)
    assert (ma.p_data[:, wire.PD_OFFSET] == O).all() and not mb.p_data[:, wire.PD_OFFSET].any()
    assert not [w for w in ma.warnings if "OFFSET" in w[0]]
    ra, rb = fers_oracle.analyze(ma), fers_oracle.analyze(mb)
    ua, ub = ra.U[0], rb.U[0]
    want = ua.copy()
    want[:, 0] += O * ua[:, 4]
    want[:, 1] -= O * ua[:, 3]
    scale = np.abs(ub).max()
    assert np.abs(ub - want).max() <= 1e-8 * scale
    assert np.abs(ua[:, 4]).max() > 1e-3 * scale / 1.0 or np.abs(ua[:, 4]).max() > 0          # the arms do rotate
    assert np.abs(ra.p_res - rb.p_res).max() <= 1e-7 * np.abs(rb.p_res).max()
    # and the offset matters: the same loads on the same nodes without it give another answer
    c = wire.parse_document({**a, "model": {**a["model"], "plate_elements": [dict(p, offset=0.0) for p in a["model"]["plate_elements"]]}})
    uc = fers_oracle.analyze(c).U[0]
    assert np.abs(uc - ua).max() > 1e-3 * np.abs(ua).max()


@pytest.mark.gpu
@pytest.mark.parametrize("ortho", [False, True])
def test_gpu_matches_oracle(ortho):
    from fers_core_b200 import engine
    a, _ = _pair(ortho)
    model = wire.parse_document(a)
    want, got = fers_oracle.analyze(model), engine.analyze(model)
    for name in ("U", "R", "p_res", "p_nodal"):
        x, y = getattr(got, name), getattr(want, name)
        assert np.abs(x - y).max() <= 1e-9 * np.abs(y).max(), name
