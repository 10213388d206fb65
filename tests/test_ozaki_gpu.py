"""fp64-equivalent Schur update on the int8 tensor cores (csrc/ozaki_syrk.cu, fers_b200_oz_syrk) against a plain torch
fp64 reference of the same update, C (lower) -= A A^T: ragged shapes (rows / K not multiples of the tile), rows of very
different magnitude, and the error model of the digit-plane split (7 planes: fp64 grade; one plane = 8 bits)."""
import numpy as np
import pytest
import torch

from fers_core_b200 import _lib, chol, engine, synth, wire

pytestmark = pytest.mark.gpu


def _update(m2, k1, S, seed, spread=3.0):
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(seed)
    lda = m2 + 8
    A = torch.randn((k1, m2), generator=g, dtype=torch.float64, device=dev)
    A *= torch.exp(spread * torch.randn((1, m2), generator=g, dtype=torch.float64, device=dev))
    Abuf = torch.zeros((k1, lda), dtype=torch.float64, device=dev)
    Abuf[:, :m2] = A                                           # column-major m2 x k1, leading dimension lda
    wb = lib.fers_b200_oz_syrk_work_bytes(m2, k1, S)
    work = torch.empty(wb, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    # the update on its own (C = 0): error of the digit-plane products; then on a random C: the update is subtracted from it
    Z = torch.zeros((m2, m2), dtype=torch.float64, device=dev)
    _lib.check(lib.fers_b200_oz_syrk(Abuf.data_ptr(), lda, Z.data_ptr(), m2, m2, k1, S, work.data_ptr(), wb, st), "oz_syrk")
    C0 = torch.randn((m2, m2), generator=g, dtype=torch.float64, device=dev)
    C = C0.clone()
    _lib.check(lib.fers_b200_oz_syrk(Abuf.data_ptr(), lda, C.data_ptr(), m2, m2, k1, S, work.data_ptr(), wb, st), "oz_syrk")
    torch.cuda.synchronize()
    ref = -(A.T @ A)                                           # symmetric: the transposed (column-major) view is the same
    amax = A.abs().amax(dim=0)
    bound = k1 * amax[:, None] * amax[None, :]                 # scale of the digit-plane truncation: K max|a_i| max|a_j|
    low = torch.triu(torch.ones_like(C, dtype=torch.bool))     # view[col, row] with row >= col
    err = (((Z - ref).abs() / bound)[low]).max().item()
    # same products; the wide kernel subtracts the low-order diagonals and the high-order ones in two steps (two roundings
    # of C, each half an ulp), the 128 x 64 kernel in one
    assert bool(((C - (C0 + Z)).abs()[low] <= 2.0 ** -51 * torch.maximum(C0.abs(), (C0 + Z).abs())[low]).all())
    upper_same = bool((torch.tril(C, -1) == torch.tril(C0, -1)).all()) and bool((torch.tril(Z, -1) == 0).all())
    return err, upper_same


@pytest.mark.parametrize("m2,k1", [(128, 64), (200, 100), (1000, 700), (1537, 1031), (4096, 320)])
def test_seven_planes_are_fp64_grade(m2, k1):
    err, upper_same = _update(m2, k1, 7, seed=m2 + k1)
    assert upper_same                      # strictly upper part of the column-major matrix is never written
    # 7 planes carry 6 + 6*8 = 54 bits below 2^E_i; the six dropped products of diagonal d = 7 weigh 6 * 2^-68 * 2^14 relative to
    # 2^(E_i + E_j) <= 4 max|a_i| max|a_j|: worst case 2^-49.4 of the bound, random signs make it smaller for long rows
    assert err < 2.0 ** -48, err


@pytest.mark.parametrize("S,tol", [(8, 2.0 ** -50), (6, 2.0 ** -40), (5, 2.0 ** -32), (4, 2.0 ** -24)])
def test_a_plane_is_worth_eight_bits(S, tol):
    err, _ = _update(777, 513, S, seed=S)
    assert (S == 8 or tol / 2 ** 12 < err) and err < tol, (S, err)      # 8 planes: at the rounding level of the fp64 reference


def test_identical_rows_and_zero_rows():
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    m2, k1 = 300, 96
    A = torch.zeros((k1, m2), dtype=torch.float64, device=dev)
    A[:, 5] = 1.0; A[:, 7] = -2.0 ** -30; A[:, 250] = torch.arange(k1, dtype=torch.float64, device=dev)
    C = torch.zeros((m2, m2), dtype=torch.float64, device=dev)
    wb = lib.fers_b200_oz_syrk_work_bytes(m2, k1, 7)
    work = torch.empty(wb, dtype=torch.uint8, device=dev)
    _lib.check(lib.fers_b200_oz_syrk(A.data_ptr(), m2, C.data_ptr(), m2, m2, k1, 7, work.data_ptr(), wb,
                                     torch.cuda.current_stream(dev).cuda_stream), "oz_syrk")
    ref = -(A.T @ A)
    low = torch.triu(torch.ones_like(C, dtype=torch.bool))
    assert torch.equal(C[low], ref[low])           # powers of two and small integers are represented exactly


def test_factorisation_with_int8_schur_updates_matches_the_fp64_pipe():
    """Whole sparse solve with the upper levels' Schur updates on the int8 path (threshold lowered so that a mid-size
    frame uses it) against the same solve on the fp64 tensor pipe: same residual class, displacements to 1e-10."""
    lib = _lib.load()
    model = synth.grid_frame(24, 24, 12, n_combos=8)
    out = {}
    try:
        for S in (0, 7):
            _lib.check(lib.fers_b200_chol_set_ozaki(S, 256), "set_ozaki")
            eng = engine.Engine(model, "cuda:0", solver="cholesky")
            rec = eng.run_linear(model.combo_factors)
            out[S] = (rec["U"].cpu().numpy(), eng.stats.relres, eng.stats.iterations)
    finally:
        _lib.check(lib.fers_b200_chol_set_ozaki(7, 1024), "set_ozaki")
    (u0, r0, _), (u8, r8, it8) = out[0], out[7]
    assert r8 < 1e-11 and it8 <= 1
    assert np.abs(u8 - u0).max() <= 1e-10 * np.abs(u0).max()
