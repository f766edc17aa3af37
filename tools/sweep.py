#!/usr/bin/env python
"""GPU-side tuning sweep (development tool, not the benchmark): pipe probes, every registered kernel on the
BASELINE shapes, cuBLAS (torch.matmul) as the 'what is achievable' yardstick.  Writes gpurun_out/sweep.json."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jblas.jl_b200 as jb  # noqa: E402
from jblas.jl_b200 import api  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def time_call(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    jb.init(0)
    res = {"probes": {}, "shapes": {}}
    for kind in ("dfma", "dmma", "ffma", "dmma_tile", "dfma_tile", "ffma_tile", "ffma2_tile"):
        vals = [jb.probe_pipe(kind, 4000 if "dmma" in kind else (10000 if "tile" in kind else 40000))[0] for _ in range(2)]
        res["probes"][kind] = vals
        print("probe", kind, vals, flush=True)
    names = jb.kernel_names()
    shapes = [("float64", 8192, 8192, 8192, 3), ("float64", 4096, 4096, 4096, 5), ("float64", 1023, 777, 4097, 20), ("float64", 65536, 64, 64, 50),
              ("float64", 256, 256, 256, 50), ("float32", 8192, 8192, 8192, 3), ("float32", 16384, 16384, 16384, 2)]
    only = sys.argv[1:]
    for dtype, M, N, K, reps in shapes:
        key = f"{dtype}_{M}x{N}x{K}"
        if only and not any(o in key for o in only):
            continue
        A = jb.mrandn(M, K, dtype, seed=1)
        X = jb.mrandn(K, N, dtype, seed=2)
        D = jb.empty_colmajor(M, N, dtype)
        flops = 2.0 * M * N * K
        row = {}
        for i, n in enumerate(names):
            if ("f64" in n) != (dtype == "float64"):
                continue
            try:
                ms = time_call(lambda: api._gemm(D, A, X, False, jb.EXPLICIT_BASE + i), reps)
                row[n] = {"ms": ms, "tflops": flops / ms / 1e9}
            except Exception as e:  # noqa: BLE001
                row[n] = {"error": str(e)}
        ms = time_call(lambda: api._gemm(D, A, X, False, None), reps)
        row["AUTO:" + jb.plan(M, K, N, dtype)["kernel"]] = {"ms": ms, "tflops": flops / ms / 1e9}
        # cuBLAS yardstick (torch.matmul on the same column-major operands)
        Dt = torch.empty_like(D.t().contiguous()).t()
        ms = time_call(lambda: torch.matmul(A, X, out=Dt) if False else torch.matmul(X.t(), A.t()), reps)
        row["cublas(torch.matmul)"] = {"ms": ms, "tflops": flops / ms / 1e9}
        res["shapes"][key] = row
        print(key, json.dumps(row), flush=True)
        del A, X, D, Dt
        torch.cuda.empty_cache()
    # batched fastmul! (HBM-bound): algorithmic GB/s against the measured HBM peak, torch.bmm (cuBLAS batched) beside it
    res["batched"] = {}
    for dtype, M, N, P, batch in [("float64", 16, 32, 14, 1_000_000), ("float32", 16, 32, 14, 2_000_000), ("float64", 32, 32, 28, 400_000),
                                  ("float64", 8, 8, 8, 4_000_000), ("float64", 64, 64, 64, 60_000)]:
        key = f"batched_{dtype}_{M}x{N}x{P}_b{batch}"
        if only and not any(o in key for o in only):
            continue
        es = 8 if dtype == "float64" else 4
        A = jb.mrandn_batch(batch, M, N, dtype, seed=3)
        X = jb.mrandn_batch(batch, N, P, dtype, seed=4)
        D = jb.empty_colmajor_batch(batch, M, P, dtype)
        nbytes = batch * (M * N + N * P + M * P) * es
        flops = 2.0 * batch * M * N * P
        ms = time_call(lambda: jb.fastmul_batched_(D, A, X), 5)
        row = {"fastmul_batched": {"ms": ms, "GBps": nbytes / ms / 1e6, "tflops": flops / ms / 1e9}}
        ms = time_call(lambda: torch.bmm(A, X), 5)
        row["torch.bmm"] = {"ms": ms, "GBps": nbytes / ms / 1e6, "tflops": flops / ms / 1e9}
        res["batched"][key] = row
        print(key, json.dumps(row), flush=True)
        del A, X, D
        torch.cuda.empty_cache()
    with open(os.path.join(OUT, "sweep.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
