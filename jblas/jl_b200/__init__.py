"""jblas.jl_b200 -- B200-native (sm_100a) drop-in for ONE hot path of JuliaBLAS/jBLAS.jl: `jmul!` (D = A*X).

Layout:  csrc/   hand-written CUDA kernels + the C ABI (include/jblas_b200.h -> libjblas_b200.so)
         api.py  host-side mirror of the reference interface (jmul_, gemm_, fastmul_, kernel_, initkernel_, mrandn)
         multigpu.py  column-block sharding of X/D across ranks + K-panel broadcast of A (torch.distributed / NCCL);
                      the single-process form of the same mode is jmul_(..., gpus=n) -> jblas_b200_mgpu_gemm_*
The CUDA library is the product: importing this package never falls back to a CPU implementation.
"""
from .api import (  # noqa: F401
    F32_3XTF32,
    F32_EXACT,
    F64_AUTO,
    F64_DMMA,
    F64_SIMT,
    EXPLICIT_BASE,
    JblasB200Error,
    Kernel,
    empty_colmajor,
    empty_colmajor_batch,
    fastmul_batched_,
    gemm_plus_c_,
    gemm_x_plus_c_,
    mrandn_batch,
    fastmul_,
    gemm_,
    init,
    initkernel_,
    jmul_,
    kernel_,
    kernel_names,
    launch_count,
    mgpu_init,
    pinned,
    mrandn,
    plan,
    probe_pipe,
    shutdown,
)
