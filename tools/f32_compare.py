#!/usr/bin/env python
"""Exact Float32 kernels side by side (device-resident, CUDA events): python tools/f32_compare.py [sizes...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import jblas.jl_b200 as jb
from jblas.jl_b200 import api

jb.init(0)
names = jb.kernel_names()
sels = [(n, jb.EXPLICIT_BASE + i) for i, n in enumerate(names) if n in ("simt_f32x2_128x128x32", "simt_f32_tma_ffma2_128x256x32_s4")]
for n in [int(a) for a in sys.argv[1:]] or [4096, 8192, 16384]:
    A, X = jb.mrandn(n, n, "float32", seed=1), jb.mrandn(n, n, "float32", seed=2)
    D = jb.empty_colmajor(n, n, "float32", fill=float("nan"))
    ref = None
    for name, sel in sels + [("AUTO -> " + jb.plan(n, n, n, "float32")["kernel"], None)]:
        for _ in range(2):
            api._gemm(D, A, X, False, sel)
        torch.cuda.synchronize()
        reps = 3 if n >= 16384 else 8
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            api._gemm(D, A, X, False, sel)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        same = "" if ref is None else (" bit-identical to the first" if torch.equal(D, ref) else " DIFFERS from the first")
        if ref is None:
            ref = D.clone()
        print(f"{n}^3 {name:44s} {ms:9.3f} ms {2.0 * n**3 / ms / 1e9:7.2f} TFLOP/s{same}", flush=True)
    del A, X, D, ref
    torch.cuda.empty_cache()
