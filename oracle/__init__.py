"""CPU oracle for the jBLAS.jl `jmul!` path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
leg may import this package.  The product (``jblas.jl_b200``) never does; it fails loudly when its CUDA
library is missing instead of falling back to anything here.

PARITY UNPINNED: the reference ships no golden vectors and cannot run here (no Julia); see
``oracle_gemm.c`` and DESIGN.md.
"""
from .cpu import (  # noqa: F401
    build,
    lib,
    oracle_gemm,
    oracle_absgemm,
    oracle_gemm_sampled,
    error_bound_ok,
    jmul_baseline,
    jmul_baseline_tile,
    fastmul_baseline_batched,
    pick_kernel_size,
    num_threads,
)
