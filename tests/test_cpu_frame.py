"""The C++ / OpenMP restatement of the oracle for plain frames (oracle/cpu_frame.cpp, the CPU baseline of bench.py)
against oracle/fers_oracle.py (numpy / scipy, pinned on the reference's own tests): stiffness values, load vectors,
displacements through the multifrontal Cholesky, member end forces and reactions on seeded frames."""
import numpy as np
import pytest

from fers_core_b200 import synth, wire
from fers_core_b200.hostlogic import rhs_plan
from oracle import cpu_frame, fers_oracle


def _dense_from_bsr(cf):
    n = cf.model.n_nodes
    K = np.zeros((6 * n, 6 * n))
    vals = cf.vals.reshape(-1, 6, 6)
    for i in range(n):
        for e in range(cf.rowptr[i], cf.rowptr[i + 1]):
            j = cf.colidx[e]
            K[6 * i:6 * i + 6, 6 * j:6 * j + 6] = vals[e]
    return K


@pytest.fixture(scope="module")
def frame():
    model = synth.grid_frame(4, 3, 3, n_combos=5)
    rng = np.random.default_rng(5)
    model.m_roll = rng.uniform(-0.6, 0.6, model.n_members)          # rolled sections exercise the full rotation
    model.m_type[rng.choice(model.n_members, 6, replace=False)] = wire.MT_TRUSS
    return model


def test_stiffness_and_loads_match_the_numpy_oracle(frame):
    cf = cpu_frame.CpuFrame(frame)
    cf.assemble()
    asm = fers_oracle.Assembly(frame)
    K = asm.K_elastic().toarray()
    got = _dense_from_bsr(cf)
    assert np.abs(got - K).max() <= 1e-13 * np.abs(K).max()
    rp, ci = fers_oracle.bsr_pattern(frame)
    assert np.array_equal(cf.rowptr, rp) and np.array_equal(cf.colidx, ci)
    F, peq = cf.load_cases()
    Fo, peqo = fers_oracle.load_vectors(asm)
    assert np.abs(F - Fo).max() <= 1e-13 * np.abs(Fo).max()
    assert np.abs(peq.reshape(peqo.shape) - peqo).max() <= 1e-13 * np.abs(peqo).max()
    _, free, dead = fers_oracle.constrain(asm.K_elastic(), frame)
    fixed = np.ones(frame.n_dof, np.uint8); fixed[free] = 0
    assert np.array_equal(cf.fixed, fixed)


@pytest.mark.parametrize("threads", [1, 3])
def test_full_pass_matches_the_numpy_oracle(frame, threads):
    cpu_frame.set_threads(threads)
    try:
        want = fers_oracle.analyze(frame, detail=False)
        plan = rhs_plan(frame)
        cf = cpu_frame.CpuFrame(frame)
        rec = cf.run_linear(plan.factors)
        assert cf.relres < 1e-12
        n = plan.n_rhs
        U = rec["U"].reshape(frame.n_nodes, 6, n).transpose(2, 0, 1)
        R = rec["R"].reshape(frame.n_nodes, 6, n).transpose(2, 0, 1)
        fl = rec["f_local"].transpose(2, 0, 1); fg = rec["f_global"].transpose(2, 0, 1); ul = rec["u_local"].transpose(2, 0, 1)
        for a, b, scale in ((U, want.U, want.U), (R, want.R, want.R), (fl, want.f_local, want.f_local),
                            (fg, want.f_global, want.f_local), (ul, want.u_local, want.U)):
            for c in range(n):
                assert np.abs(a[c] - b[c]).max() <= 1e-9 * np.abs(scale[c]).max()
    finally:
        cpu_frame.set_threads(cpu_frame.os.cpu_count() or 1)


def test_larger_frame_with_a_deep_front_tree():
    model = synth.grid_frame(12, 10, 8, n_combos=8, solve_loadcases=False)
    cf = cpu_frame.CpuFrame(model)
    rec = cf.run_linear(model.combo_factors)
    assert cf.relres < 1e-11 and cf._cplan.info.n_levels >= 3
    want = fers_oracle.analyze(model, detail=False)
    U = rec["U"].reshape(model.n_nodes, 6, -1).transpose(2, 0, 1)
    assert np.abs(U - want.U).max() <= 1e-9 * np.abs(want.U).max()


def test_rejects_models_outside_the_subset(frame):
    import dataclasses
    m = dataclasses.replace(frame, m_rel=np.where(np.arange(frame.n_members) == 0, 0, -1).astype(np.int32),
                            rel=np.full((1, wire.REL_STRIDE), -1.0))
    with pytest.raises(ValueError):
        cpu_frame.CpuFrame(m)
