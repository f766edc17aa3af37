#!/usr/bin/env python
"""Tall-skinny Float64 shapes: the dedicated kernel (AUTO) against the tile kernels and cuBLAS, cold operands (rotating sets
larger than L2), CUDA events over back-to-back launches."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jblas.jl_b200 as jb  # noqa: E402
from jblas.jl_b200 import api  # noqa: E402

jb.init(0)
names = jb.kernel_names()
tile = {n: jb.EXPLICIT_BASE + i for i, n in enumerate(names) if n in ("dmma_tma_f64_64x32x32_s4_x2", "dmma_tma_f64_64x64x32_s3_x2", "dmma_tma_f64_64x64x64_s3")}
skinny = jb.EXPLICIT_BASE + names.index("dmma_skinny_f64_16x64_xres_w12")
xreg = jb.EXPLICIT_BASE + names.index("dmma_skinny_f64_16x16_xreg_w8")
team = jb.EXPLICIT_BASE + names.index("dmma_skinny_f64_16x16_xreg_team_w16")


def timeit(fn, bufs, reps=200):
    for b in bufs:
        fn(*b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for i in range(reps):
            fn(*bufs[i % len(bufs)])
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


for (M, N, K) in [(65536, 64, 64), (16384, 64, 64), (32768, 64, 64), (131072, 64, 64), (300000, 64, 64), (1000000, 64, 64), (65536, 32, 64), (65536, 64, 128),
                  (65536, 48, 72), (65536, 48, 64), (262144, 16, 32), (131072, 64, 32), (300000, 32, 64), (262144, 16, 64), (65536, 8, 64)]:
    nbytes = (M * K + K * N + M * N) * 8
    sets = max(2, min(12, -(-400_000_000 // nbytes)))
    bufs = [(jb.empty_colmajor(M, N, "float64"), jb.mrandn(M, K, "float64", seed=2 * i + 1), jb.mrandn(K, N, "float64", seed=2 * i + 2)) for i in range(sets)]
    row = {"skinny": timeit(lambda D, A, X: api._gemm(D, A, X, False, skinny), bufs) if K % 8 == 0 and K <= 128 else float("nan")}
    if K in (32, 64):
        row["xreg"] = timeit(lambda D, A, X: api._gemm(D, A, X, False, xreg), bufs)
        row["team"] = timeit(lambda D, A, X: api._gemm(D, A, X, False, team), bufs)
    for n, sel in tile.items():
        row[n[13:]] = timeit(lambda D, A, X, sel=sel: api._gemm(D, A, X, False, sel), bufs)
    row["cuBLAS"] = timeit(lambda D, A, X: torch.matmul(A, X, out=D.t().contiguous().t() if False else None), bufs)
    flops = 2.0 * M * N * K
    ideal = max(flops / 36.8e12, nbytes / 6.4559e12) * 1e6
    print(f"{M}x{N}x{K}: ideal max(FP64, HBM) {ideal:6.1f} us | " + " | ".join(f"{k} {v:7.2f} us ({flops / v / 1e6:5.1f} TF/s)" for k, v in row.items()) +
          f" | AUTO -> {jb.plan(M, K, N)['kernel']}", flush=True)
    del bufs
    torch.cuda.empty_cache()
