#!/usr/bin/env python
"""End-to-end timing of the single-process multi-GPU entry (jblas_b200_mgpu_gemm_f64) on pinned host matrices, the leg
bench.py reports as `e2e` at N > 1, in isolation:   python tools/mgpu_e2e.py [gpu counts ...]   (JBLAS_B200_TRACE=1: stage timeline)
Weak workload of bench.py: M = K = 8192, 8192 columns of X and D per GPU."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import jblas.jl_b200 as jb  # noqa: E402

counts = [int(a) for a in sys.argv[1:]] or [n for n in (1, 2, 4, 8) if n <= torch.cuda.device_count()]
M = K = 8192
per_gpu = int(os.environ.get("COLS_PER_GPU", "8192"))
jb.init(0)
nmax = max(counts)
A = np.asfortranarray(jb.mrandn(M, K, seed=1).cpu().numpy())
X = np.empty((K, per_gpu * nmax), order="F")
for c0 in range(0, X.shape[1], 4096):
    X[:, c0:c0 + 4096] = jb.mrandn(K, 4096, seed=2, first_col=c0).cpu().numpy()
D = np.full((M, per_gpu * nmax), np.nan, order="F")
torch.cuda.empty_cache()
with jb.pinned(A, X, D):
    base = None
    for n in counts:
        Xn, Dn = X[:, :per_gpu * n], D[:, :per_gpu * n]
        jb.jmul_(Dn, A, Xn, gpus=n)
        steps = 3
        t = time.perf_counter()
        for _ in range(steps):
            jb.jmul_(Dn, A, Xn, gpus=n)
        ms = (time.perf_counter() - t) / steps * 1e3
        tf = 2.0 * M * K * per_gpu * n / (ms * 1e-3) / 1e12
        base = base or tf
        print(f"gpus {n}: {ms:7.2f} ms  {tf:7.1f} TFLOP/s  efficiency vs the first count {tf / (base * n / counts[0]):.3f}  "
              f"(H2D {(A.nbytes + Xn.nbytes) / 2**30:.2f} GiB, D2H {Dn.nbytes / 2**30:.2f} GiB per call)", flush=True)
        assert not np.isnan(Dn).any()
