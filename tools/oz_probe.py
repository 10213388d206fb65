#!/usr/bin/env python
"""Probe of the int8 tensor-core Schur update (fers_b200_oz_syrk): accuracy against torch fp64 and device time.
    python tools/oz_probe.py [--shapes m2xk1,...] [--slices 8,7] [--reps 3]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fers_core_b200 import _lib  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--shapes", default="200x100,1000x700,4096x1024")
p.add_argument("--slices", default="8")
p.add_argument("--reps", type=int, default=3)
p.add_argument("--no-check", action="store_true")
a = p.parse_args()
lib = _lib.load()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
st = torch.cuda.current_stream(dev).cuda_stream
out = []
for shape in a.shapes.split(","):
    m2, k1 = (int(v) for v in shape.split("x"))
    g = torch.Generator(device=dev).manual_seed(m2 * 31 + k1)
    lda = m2 + 24
    Abuf = torch.zeros((k1, lda), dtype=torch.float64, device=dev)          # column-major m2 x k1 with leading dimension lda
    A = torch.randn((k1, m2), generator=g, dtype=torch.float64, device=dev)
    A *= torch.exp(4.0 * torch.randn((1, m2), generator=g, dtype=torch.float64, device=dev))      # rows of very different size
    A *= torch.exp(2.0 * torch.randn((k1, m2), generator=g, dtype=torch.float64, device=dev))     # and a spread inside the rows
    Abuf[:, :m2] = A
    C0 = torch.randn((m2, m2), generator=g, dtype=torch.float64, device=dev)
    for S in (int(v) for v in a.slices.split(",")):
        wb = lib.fers_b200_oz_syrk_work_bytes(m2, k1, S)
        work = torch.empty(wb, dtype=torch.uint8, device=dev)
        C = C0.clone()
        _lib.check(lib.fers_b200_oz_syrk(Abuf.data_ptr(), lda, C.data_ptr(), m2, m2, k1, S, work.data_ptr(), wb, st), "oz_syrk")
        torch.cuda.synchronize()
        row = {"m2": m2, "k1": k1, "slices": S}
        if not a.no_check:
            # column-major C: C[col, row] in torch's row-major view; the lower triangle (row >= col) is view[col, row] with row >= col
            ref = C0 - (A.T @ A).T                      # (A^T)^T... symmetric product, stored transposed consistently
            got_low = torch.triu(C)                      # view[col, row], row >= col
            ref_low = torch.triu(ref)
            # norm-wise error bound of the update: |a_i| |a_j|
            nrm = torch.linalg.norm(A, dim=0)
            bound = torch.triu(nrm[:, None] * nrm[None, :]) + 1e-300
            err = ((got_low - ref_low).abs() / bound).max().item()
            untouched = (torch.tril(C, -1) - torch.tril(C0, -1)).abs().max().item()
            row.update(max_err_over_norms=err, upper_untouched=untouched == 0.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(a.reps):
            e0.record()
            _lib.check(lib.fers_b200_oz_syrk(Abuf.data_ptr(), lda, C.data_ptr(), m2, m2, k1, S, work.data_ptr(), wb, st), "oz_syrk")
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = min(ts)
        row.update(ms=ms, tflops_fp64_equiv=k1 * m2 * (m2 + 1.0) / (ms * 1e-3) / 1e12)
        out.append(row)
        print(json.dumps(row), flush=True)
        del work, C
    del Abuf, A, C0
    torch.cuda.empty_cache()
