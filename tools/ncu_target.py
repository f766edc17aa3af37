#!/usr/bin/env python
"""Tiny launcher for ncu captures: run `reps` launches of one kernel selector on one shape.
    python tools/ncu_target.py float64 8192 8192 8192 <selector|auto|dmma|simt|tf32x3> [reps] [sets]
`sets` > 1 rotates that many operand sets so that no launch finds its operands in L2 (cold captures of small shapes)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jblas.jl_b200 as jb  # noqa: E402
from jblas.jl_b200 import api  # noqa: E402

if sys.argv[1] == "batched":  # python tools/ncu_target.py batched M N P batch [reps]   (jBLAS names: D MxP = A MxN * X NxP)
    M, N, P, batch = (int(v) for v in sys.argv[2:6])
    reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
    bdt = sys.argv[7] if len(sys.argv) > 7 else "float64"
    jb.init(0)
    A = jb.mrandn_batch(batch, M, N, bdt, seed=1)
    X = jb.mrandn_batch(batch, N, P, bdt, seed=2)
    D = jb.empty_colmajor_batch(batch, M, P, bdt)
    for _ in range(reps):
        jb.fastmul_batched_(D, A, X)
    torch.cuda.synchronize()
    print("done batched")
    sys.exit(0)
dtype, M, N, K, sel = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
sel = {"auto": None, "dmma": jb.F64_DMMA, "simt": jb.F64_SIMT, "tf32x3": jb.F32_3XTF32}.get(sel, None) if not sel.isdigit() else int(sel)
sets = int(sys.argv[7]) if len(sys.argv) > 7 else 1
jb.init(0)
bufs = [(jb.empty_colmajor(M, N, dtype), jb.mrandn(M, K, dtype, seed=2 * i + 1), jb.mrandn(K, N, dtype, seed=2 * i + 2)) for i in range(sets)]
for r in range(reps):
    D, A, X = bufs[r % sets]
    api._gemm(D, A, X, False, sel)
torch.cuda.synchronize()
print("done", jb.plan(M, K, N, dtype, kernel=sel)["kernel"])
