"""Randomised sweep: small 3D frames with jittered geometry, rolled members, trusses, end releases and springs, shear-centre
and end offsets, products of inertia, pinned and sprung supports, slabs (Mindlin / thin-plate, orthotropic, offset), pretensioned
cable braces, linear and P-Delta -- every result field of the GPU path
against the oracle.  The fixed cases elsewhere pin single features; this one pins their combinations."""
import numpy as np
import pytest

from fers_core_b200 import synth, wire
from oracle import fers_oracle

SEEDS = list(range(10))


def random_model(seed: int):
    rng = np.random.default_rng(1000 + seed)
    nx, nz, ny = (int(v) for v in rng.integers(2, 5, size=3))
    second_order = seed % 3 == 0
    model = synth.grid_frame(nx, nz, ny, n_combos=5, timoshenko=bool(seed % 2), order="NONLINEAR" if second_order else "LINEAR",
                             solve_loadcases=True)
    nn, nm = model.n_nodes, model.n_members
    base = model.node_support_id >= 0
    jitter = rng.normal(size=(nn, 3)) * np.array([0.25, 0.12, 0.25])
    jitter[base] = 0.0
    model.xyz = model.xyz + jitter
    model.m_roll = np.where(rng.random(nm) < 0.3, rng.uniform(-1.2, 1.2, nm), 0.0)
    column = np.isclose(model.xyz[model.m_n1, 0], model.xyz[model.m_n2, 0], atol=1.0) & \
        (np.abs(model.xyz[model.m_n2, 1] - model.xyz[model.m_n1, 1]) > 2.0)
    model.m_type = np.where(~column & (rng.random(nm) < 0.08), wire.MT_TRUSS, model.m_type).astype(np.int32)
    # member modifier records: releases / springs, shear-centre offsets, end offsets, product of inertia
    rel = np.full((5, wire.REL_STRIDE), 0.0)
    rel[:, :12] = wire.RIGID_RELEASE
    rel[0, 11] = 0.0                                   # pinned about local z at the end
    rel[1, [4, 5]] = (2.0e6, 0.0)                      # spring about y, pin about z at the start
    rel[2, 12:14] = (0.02, -0.015)                     # shear-centre offset
    rel[3, 14:20] = (0.05, -0.03, 0.02, -0.04, 0.0, 0.03)     # end offsets
    rel[3, 5] = 5.0e5
    rel[4, wire.REL_EIYZ] = 0.2 * float(np.sqrt(model.prop[1, wire.P_EIY] * model.prop[1, wire.P_EIZ]))
    pick = (rng.random(nm) < 0.18) & ~column & (model.m_type == wire.MT_NORMAL)
    model.rel = rel
    model.m_rel = np.where(pick, rng.integers(0, 5, nm), -1).astype(np.int32)
    # supports: a third of the base nodes pinned, one of them on a vertical spring
    base_idx = np.nonzero(base)[0]
    pinned = base_idx[rng.random(len(base_idx)) < 0.33]
    model.bc_code = model.bc_code.copy(); model.bc_k = model.bc_k.copy()
    model.bc_code[pinned, 3:] = wire.BC_FREE
    if len(pinned):
        model.bc_code[pinned[0], 1] = wire.BC_SPRING
        model.bc_k[pinned[0], 1] = 4.0e8
    if seed % 2 and not second_order:
        # a floor slab on the top storey of some bays: two triangles per bay, random thickness, behaviour, local-x angle,
        # orthotropic material and offset; uniform pressure in the gravity case
        px, pz = nx + 1, nz + 1
        nid = lambda i, k, j: i + px * (k + pz * j)
        tris, data, flags = [], [], []
        for i in range(nx):
            for k in range(nz):
                if rng.random() < 0.6:
                    a, b, c, d = nid(i, k, ny), nid(i + 1, k, ny), nid(i + 1, k + 1, ny), nid(i, k + 1, ny)
                    for t3 in ((a, b, c), (a, c, d)):
                        rec = np.zeros(wire.NPDATA)
                        E = 30e9
                        rec[:7] = (rng.uniform(0.12, 0.3), E, 0.2, 1.0, rng.uniform(0.5, 1.0), 1.0, rng.uniform(-1.0, 1.0))
                        if rng.random() < 0.5:
                            rec[wire.PD_ORTHO] = 1.0
                            rec[wire.PD_EX:wire.PD_GYZ + 1] = (E, 0.4 * E, 0.2 * E, 0.15, 0.25 * E, 0.1 * E)
                        if rng.random() < 0.5:
                            rec[wire.PD_OFFSET] = rng.uniform(-0.15, 0.15)
                        fl = int(rng.choice([1, 2, 3]))
                        if fl & 2 and rng.random() < 0.4:
                            fl |= 8                                   # thin-plate (DKT) bending
                        tris.append(t3); data.append(rec); flags.append(fl)
        if tris:
            model.p_ids = np.arange(1, len(tris) + 1, dtype=np.int64)
            model.p_nodes = np.array(tris, np.int32)
            model.p_data = np.array(data)
            model.p_flags = np.array(flags, np.int32)
            lc = model.load_cases[0]
            lc.pp_plate = np.arange(len(tris), dtype=np.int32)
            lc.pp_vec = np.tile(np.array([0.0, -4.0e3, 0.0]), (len(tris), 1))
    if seed % 5 == 3:
        _add_cable_braces(model, rng, nx, nz, ny)
    return model


def _add_cable_braces(model, rng, nx, nz, ny):
    """Pretensioned cable diagonals (elastic catenaries with a little own weight) across some bays of the first storeys."""
    px, pz = nx + 1, nz + 1
    nid = lambda i, k, j: i + px * (k + pz * j)          # noqa: E731
    pairs = []
    for i in range(nx):
        for k in range(pz):
            if rng.random() < 0.4:
                j = int(rng.integers(0, ny))
                pairs.append((nid(i, k, j), nid(i + 1, k, j + 1)) if rng.random() < 0.5 else (nid(i + 1, k, j), nid(i, k, j + 1)))
    if not pairs:
        pairs = [(nid(0, 0, 0), nid(1, 0, 1))]
    nc, nm = len(pairs), model.n_members
    n1 = np.array([p[0] for p in pairs], np.int32); n2 = np.array([p[1] for p in pairs], np.int32)
    prop_row = 0
    model.m_ids = np.concatenate([model.m_ids, model.m_ids.max() + 1 + np.arange(nc)])
    model.m_n1 = np.concatenate([model.m_n1, n1]).astype(np.int32)
    model.m_n2 = np.concatenate([model.m_n2, n2]).astype(np.int32)
    model.m_type = np.concatenate([model.m_type, np.full(nc, wire.MT_CABLE)]).astype(np.int32)
    model.m_roll = np.concatenate([model.m_roll, np.zeros(nc)])
    model.m_prop = np.concatenate([model.m_prop, np.full(nc, prop_row)]).astype(np.int32)
    model.m_rel = np.concatenate([model.m_rel, np.full(nc, -1)]).astype(np.int32)
    for name in ("m_ref_member", "m_ref_node"):
        arr = getattr(model, name)
        if len(arr) == nm:
            setattr(model, name, np.concatenate([arr, np.full(nc, -1, arr.dtype)]))
    for name in ("m_classification", "m_section_raw"):
        lst = getattr(model, name)
        if len(lst) == nm:
            setattr(model, name, list(lst) + [None] * nc)
    chord = np.linalg.norm(model.xyz[n2] - model.xyz[n1], axis=1)
    ea = np.full(nc, 0.02 * float(model.prop[prop_row, 0]))          # a rod of 2 % of the column's axial stiffness
    model.c_member = (nm + np.arange(nc)).astype(np.int32)
    model.c_ea = ea
    model.c_l0 = chord / (1.0 + rng.uniform(2e3, 3e4, nc) / ea)
    model.c_w = np.tile(np.array([0.0, -25.0, 0.0]), (nc, 1))
    for li, lc in enumerate(model.load_cases):
        lc.cab_q = np.zeros((nc, 3))
        if li == 1:
            lc.cab_q[:, 0] = 40.0                                      # a wind-like uniform load on the cables in one case


def _fields(res):
    out = {n: getattr(res, n) for n in ("U", "R", "f_local", "f_global", "sec", "lmax", "lmin")}
    if res.p_res.size:
        out.update(p_res=res.p_res, p_nodal=res.p_nodal)
    return out


@pytest.mark.parametrize("seed", SEEDS[:4])
def test_oracle_solves_the_random_models(seed):
    model = random_model(seed)
    res = fers_oracle.analyze(model)
    assert np.isfinite(res.U).all() and np.abs(res.U).max() < 1.0           # a structure, not a mechanism
    assert res.converged.all()
    # global equilibrium of every column: support reactions (R = K u - f at the supported nodes holds the force of a support
    # spring as well) + applied loads (the nodal load vectors of the cases, the own weight and the uniform loads of the cables)
    # sum to zero
    asm = fers_oracle.Assembly(model, model.xyz)
    F_lc, _ = fers_oracle.load_vectors(asm)
    plan = fers_oracle.rhs_plan(model)
    for col in range(plan.n_rhs):
        applied = (plan.factors[col] @ F_lc).reshape(-1, 6)[:, :3].sum(0)
        if len(model.c_member):
            q = model.c_w + sum(plan.factors[col, li] * lc.cab_q for li, lc in enumerate(model.load_cases))
            applied = applied + (q * model.c_l0[:, None]).sum(0)
        held = res.R[col, :, :3].sum(0)
        assert np.abs(held + applied).max() < 1e-7 * max(np.abs(applied).max(), 1.0), (seed, col, held, applied)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", SEEDS)
def test_gpu_matches_oracle_on_random_models(seed):
    from fers_core_b200 import engine
    model = random_model(seed)
    want = fers_oracle.analyze(model)
    got = engine.analyze(model, solver="cholesky" if seed % 4 == 1 else "auto")
    assert np.array_equal(got.converged, want.converged) and np.array_equal(got.iterations, want.iterations)
    for name, y in _fields(want).items():
        x = getattr(got, name)
        assert np.abs(x - y).max() <= 2e-9 * np.abs(y).max(), (seed, name, np.abs(x - y).max() / np.abs(y).max())
