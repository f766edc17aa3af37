#!/usr/bin/env python
"""The cta_group::2 3xTF32 kernel against the single-CTA one (same MMA sequence per element -> identical bits) and the oracle
bound, then both timed at 8192^3 and 16384^3 (kernel + split pre-pass, CUDA events)."""
import os
import sys
import subprocess

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jblas.jl_b200 as jb  # noqa: E402
from jblas.jl_b200 import api  # noqa: E402
import oracle  # noqa: E402

jb.init(0)
names = jb.kernel_names()
one = jb.EXPLICIT_BASE + names.index("tf32x3_tcgen05_f32_128x256x32_s2")
pair = jb.EXPLICIT_BASE + names.index("tf32x3_tcgen05_2cta_f32_256x256x32_s3")
rng = np.random.Generator(np.random.PCG64(7))
ok_all = True
for (M, N, K) in [(256, 256, 32), (256, 256, 256), (512, 768, 96), (300, 100, 270), (129, 513, 40), (1024, 2048, 512), (2048, 1024, 4096)]:
    A = np.asfortranarray(rng.standard_normal((K, M)).T.astype(np.float32))
    X = np.asfortranarray(rng.standard_normal((N, K)).T.astype(np.float32))
    dA = torch.from_numpy(np.ascontiguousarray(A.T)).cuda().t()
    dX = torch.from_numpy(np.ascontiguousarray(X.T)).cuda().t()
    out = {}
    for name, sel in (("one", one), ("pair", pair)):
        dD = jb.empty_colmajor(M, N, "float32", fill=float("nan"))
        api._gemm(dD, dA, dX, False, sel)
        torch.cuda.synchronize()
        out[name] = np.asfortranarray(dD.cpu().numpy())
    want = oracle.oracle_gemm(A, X)
    okb, worst = oracle.error_bound_ok(out["pair"], want, A, X, extra_rel=2.0 ** -18)
    same = out["one"].tobytes() == out["pair"].tobytes()
    nbad = int(np.sum(out["one"] != out["pair"]))
    ok_all &= okb
    print(f"{M}x{N}x{K}: pair within bound {okb} (err/bound {worst:.3g}), identical to the single-CTA kernel: {same} ({nbad} elements differ)", flush=True)
    # accumulate form
    C = np.asfortranarray(rng.standard_normal((N, M)).T.astype(np.float32))
    dC = torch.from_numpy(np.ascontiguousarray(C.T)).cuda().t().clone()
    dC2 = dC.clone()
    api._gemm(dC, dA, dX, True, one); api._gemm(dC2, dA, dX, True, pair)
    torch.cuda.synchronize()
    print("   accumulate identical:", bool((dC == dC2).all().item()), flush=True)
for n in (8192, 16384):
    A = jb.mrandn(n, n, "float32", seed=1); X = jb.mrandn(n, n, "float32", seed=2); D = jb.empty_colmajor(n, n, "float32")
    for name, sel in (("single-CTA 128x256", one), ("pair 256x256", pair), ("single-CTA 128x256", one), ("pair 256x256", pair)):
        for _ in range(2): api._gemm(D, A, X, False, sel)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 6 if n == 8192 else 3
        e0.record()
        for _ in range(reps): api._gemm(D, A, X, False, sel)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{n}^3 {name}: {ms:.3f} ms incl. split pre-pass = {2.0 * n ** 3 / ms / 1e9:.1f} TFLOP/s", flush=True)
print("ALL WITHIN BOUND" if ok_all else "BOUND VIOLATED")
