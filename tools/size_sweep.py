#!/usr/bin/env python
"""AUTO kernel vs cuBLAS over a range of square and rectangular sizes (development aid). Writes gpurun_out/size_sweep.json."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jblas.jl_b200 as jb
from jblas.jl_b200 import api
from tools.sweep import time_call

jb.init(0)
out = {}
shapes = [(n, n, n) for n in (128, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096, 6144)] + [(8192, 512, 8192), (512, 8192, 512), (4096, 4096, 256), (256, 4096, 4096), (10000, 3000, 7000)]
for dtype in ("float64", "float32"):
    for M, N, K in shapes:
        A = jb.mrandn(M, K, dtype, seed=1); X = jb.mrandn(K, N, dtype, seed=2); D = jb.empty_colmajor(M, N, dtype)
        reps = 20 if M * N * K < 2 ** 32 else 5
        ms = time_call(lambda: api._gemm(D, A, X, False, None), reps)
        mc = time_call(lambda: torch.matmul(X.t(), A.t()), reps)
        fl = 2.0 * M * N * K
        k = jb.plan(M, K, N, dtype)["kernel"]
        out[f"{dtype}_{M}x{N}x{K}"] = {"kernel": k, "ms": ms, "tflops": fl / ms / 1e9, "cublas_ms": mc, "cublas_tflops": fl / mc / 1e9}
        print(f"{dtype} {M}x{N}x{K:<6d} {k:34s} {ms:9.4f} ms {fl/ms/1e9:7.2f} TF | cuBLAS {mc:9.4f} ms {fl/mc/1e9:7.2f} TF | ratio {mc/ms:5.2f}", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "size_sweep.json"), "w"), indent=1)
