"""Turning a structure about the global Y axis -- the reference direction of the member local-axes rule
(fers_core/members/member.py:309-354) -- turns its answer with it: the reference states this with examples 084 / 085
(085_Triple_Cantilever_3D_Rotated_45Y.py:6-22, "direct comparison of global displacements with the base case").  Here on the
golden 3D frame (rolled members, distributed / point loads and moments in global and local axes): displacements, rotations and
reactions turn as vectors, member end forces and section forces in LOCAL axes do not change at all."""
import copy

import numpy as np
import pytest

from fers_core_b200 import wire
from oracle import fers_oracle
from tests import helpers

ANGLES = [33.0, -45.0, 180.0]


def _turned(doc, deg):
    c, s = np.cos(np.radians(deg)), np.sin(np.radians(deg))
    Q = np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])            # rotation about global Y
    out = copy.deepcopy(doc)
    # a member parallel to Y takes global Z as its reference direction, which does not turn with the structure: its section
    # is turned along by its roll angle (local y = Z, local z = +-X for a member pointing up / down)
    xyz = {n["id"]: np.array([n["X"], n["Y"], n["Z"]], float) for n in doc["model"]["nodes"]}
    for m in out["model"]["members"]:
        d = xyz[m["end_node_id"]] - xyz[m["start_node_id"]]
        if abs(d[1]) > (1.0 - 1e-6) * np.linalg.norm(d):
            m["rotation_angle"] = float(m.get("rotation_angle") or 0.0) + deg * np.sign(d[1])
    for n in out["model"]["nodes"]:
        n["X"], n["Y"], n["Z"] = (Q @ np.array([n["X"], n["Y"], n["Z"]], float)).tolist()
    for lc in out["analysis"]["load_cases"]:
        for key in ("nodal_loads", "nodal_moments", "distributed_loads", "member_point_loads", "member_point_moments"):
            for rec in lc.get(key) or []:
                if str(rec.get("axes") or "global").lower() == "local":
                    continue                                               # given in member axes: turns with the member
                rec["direction"] = (Q @ np.array(rec["direction"], float)).tolist()
    return out, Q


@pytest.mark.parametrize("deg", ANGLES)
def test_oracle_answer_turns_with_the_structure(deg):
    doc = helpers.golden("space_frame_3d")["doc"]
    turned, Q = _turned(doc, deg)
    a = fers_oracle.analyze(wire.parse_document(doc))
    b = fers_oracle.analyze(wire.parse_document(turned))
    scale = np.abs(a.U).max()
    for block in (slice(0, 3), slice(3, 6)):
        assert np.abs(b.U[..., block] - a.U[..., block] @ Q.T).max() < 1e-9 * scale
        assert np.abs(b.R[..., block] - a.R[..., block] @ Q.T).max() < 1e-9 * np.abs(a.R).max()
    assert np.abs(b.f_local - a.f_local).max() < 1e-8 * np.abs(a.f_local).max()
    assert np.abs(b.sec - a.sec).max() < 1e-8 * np.abs(a.f_local).max()
