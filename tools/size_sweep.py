#!/usr/bin/env python
"""AUTO kernel vs cuBLAS over a range of square and rectangular sizes (development aid). Writes gpurun_out/size_sweep.json.

`--all` additionally times every registered kernel of the dtype's AUTO family on shapes below 2^33 flops, so the
planner's efficiency table (capi.cu: g_kernels[].eff) can be checked against what actually wins."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jblas.jl_b200 as jb
from jblas.jl_b200 import api
from tools.sweep import time_call

jb.init(0)
out = {}
ALL = "--all" in sys.argv
dtypes = [a for a in sys.argv[1:] if a.startswith("float")] or ["float64", "float32"]
names = jb.kernel_names()
shapes = [(n, n, n) for n in (128, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096, 6144)] + [(8192, 512, 8192), (512, 8192, 512), (4096, 4096, 256), (256, 4096, 4096), (10000, 3000, 7000)]
if "--c4" in sys.argv:  # BASELINE configs[3]: the ragged and the tall-skinny shape
    shapes = [(1023, 777, 4097), (65536, 64, 64), (1024, 776, 4096)]
for dtype in dtypes:
    for M, N, K in shapes:
        A = jb.mrandn(M, K, dtype, seed=1); X = jb.mrandn(K, N, dtype, seed=2); D = jb.empty_colmajor(M, N, dtype)
        reps = 20 if M * N * K < 2 ** 32 else 5
        ms = time_call(lambda: api._gemm(D, A, X, False, None), reps)
        mc = time_call(lambda: torch.matmul(X.t(), A.t()), reps)
        fl = 2.0 * M * N * K
        k = jb.plan(M, K, N, dtype)["kernel"]
        row = {"kernel": k, "ms": ms, "tflops": fl / ms / 1e9, "cublas_ms": mc, "cublas_tflops": fl / mc / 1e9}
        line = f"{dtype} {M}x{N}x{K:<6d} {k:34s} {ms:9.4f} ms {fl/ms/1e9:7.2f} TF | cuBLAS {mc:9.4f} ms {fl/mc/1e9:7.2f} TF | ratio {mc/ms:5.2f}"
        if ALL and M * N * K <= 2 ** 33:
            per = {}
            for i, n in enumerate(names):
                fam = ("dmma_tma_f64" in n) if dtype == "float64" else ("simt_f32x2" in n or "simt_f32_tma" in n)
                if not fam:
                    continue
                try:
                    per[n] = time_call(lambda: api._gemm(D, A, X, False, jb.EXPLICIT_BASE + i), reps)
                except Exception as e:  # noqa: BLE001
                    per[n] = None
            row["per_kernel_ms"] = per
            bestn = min((n for n in per if per[n]), key=lambda n: per[n])
            line += f" | best {bestn} {per[bestn]:.4f} ms"
        out[f"{dtype}_{M}x{N}x{K}"] = row
        print(line, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "size_sweep.json"), "w"), indent=1)
