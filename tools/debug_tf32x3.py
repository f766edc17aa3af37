#!/usr/bin/env python
"""Structured probes of the 3xTF32 tcgen05 kernel (development aid): which (m, n, k) the hardware thinks it multiplied.
All probe values are small integers, exact in TF32, so any wrong entry is a layout/descriptor problem, not rounding."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jblas.jl_b200 as jb  # noqa: E402
from jblas.jl_b200 import api  # noqa: E402


def run(A, X, sel):
    M, K = A.shape
    N = X.shape[1]
    dA = torch.from_numpy(np.ascontiguousarray(A.T)).cuda().t()
    dX = torch.from_numpy(np.ascontiguousarray(X.T)).cuda().t()
    dD = jb.empty_colmajor(M, N, "float32", fill=float("nan"))
    api._gemm(dD, dA, dX, False, sel)
    torch.cuda.synchronize()
    return dD.cpu().numpy()


def main():
    jb.init(0)
    names = jb.kernel_names()
    for i, n in enumerate(names):
        if not n.startswith("tf32x3"):
            continue
        sel = jb.EXPLICIT_BASE + i
        print("=====", n)
        M, N, K = 128, 256, 32
        # T2: row mapping
        A = np.zeros((M, K), np.float32); X = np.zeros((K, N), np.float32)
        A[:, 0] = np.arange(1, M + 1); X[0, :] = 1
        D = run(A, X, sel)
        want = A @ X
        print("T2 rows ok:", np.array_equal(D, want), "col0 head", D[:8, 0], "col0[32:36]", D[32:36, 0], "any nan", np.isnan(D).any())
        # T3: column mapping
        A[:] = 0; X[:] = 0; A[:, 0] = 1; X[0, :] = np.arange(1, N + 1)
        D = run(A, X, sel)
        print("T3 cols ok:", np.array_equal(D, A @ X), "row0 head", D[0, :8], "row0[128:132]", D[0, 128:132])
        # T1: k pairing
        bad = []
        for ka in range(K):
            A[:] = 0; X[:] = 0; A[:, ka] = 1; X[ka, :] = 1
            D = run(A, X, sel)
            if not np.array_equal(D, np.ones((M, N), np.float32)):
                bad.append((ka, float(np.nanmean(D))))
        print("T1 k pairing bad:", bad[:10], "of", len(bad))
        # T4: second k tile / accumulate across stages
        K2 = 96
        A = np.zeros((M, K2), np.float32); X = np.zeros((K2, N), np.float32)
        A[:, 40] = 2; X[40, :] = 3; A[:, 70] = 1; X[70, :] = 5
        D = run(A, X, sel)
        print("T4 multi-stage ok:", np.array_equal(D, np.full((M, N), 11, np.float32)), D[0, :4])
        # T5: random, error stats
        rng = np.random.default_rng(0)
        A = rng.standard_normal((300, 200)).astype(np.float32); X = rng.standard_normal((200, 270)).astype(np.float32)
        D = run(A, X, sel)
        ref = A.astype(np.float64) @ X.astype(np.float64)
        print("T5 random max rel err:", float(np.abs(D - ref).max() / np.abs(ref).max()))


if __name__ == "__main__":
    main()
