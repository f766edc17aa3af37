"""ctypes binding of libjblas_b200.so -- the same C ABI a Julia `ccall` binds (include/jblas_b200.h).

There is NO fallback: if the library is missing, or no CUDA device is present when a compute entry point
is called, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libjblas_b200.so")

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_vp = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/jblas_b200.h one to one
_GEMM = [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_int]
_JMUL = [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64]
_KERN = [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64]
_FUSED = [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_int]  # D, A, X, C, M, K, N, ldd, lda, ldx, ldc, sel
PROTOTYPES = {
    "jblas_b200_init": (c_int, [c_int]),
    "jblas_b200_shutdown": (c_int, []),
    "jblas_b200_version": (c_int, []),
    "jblas_b200_last_error": (ctypes.c_char_p, []),
    "jblas_b200_device_count": (c_int, []),
    "jblas_b200_gemm_f64": (c_int, _GEMM),
    "jblas_b200_gemm_f32": (c_int, _GEMM),
    "jblas_b200_mgpu_init": (c_int, [c_int]),
    "jblas_b200_mgpu_gemm_f64": (c_int, _GEMM + [c_int]),
    "jblas_b200_mgpu_gemm_f32": (c_int, _GEMM + [c_int]),
    "jblas_b200_jmul_f64": (c_int, _JMUL),
    "jblas_b200_jmul_f32": (c_int, _JMUL),
    "jblas_b200_fastmul_f64": (c_int, _JMUL),
    "jblas_b200_fastmul_f32": (c_int, _JMUL),
    "jblas_b200_kernel_f64": (c_int, _KERN),
    "jblas_b200_initkernel_f64": (c_int, _KERN),
    "jblas_b200_kernel_f32": (c_int, _KERN),
    "jblas_b200_initkernel_f32": (c_int, _KERN),
    "jblas_b200_gemm_f64_dev": (c_int, _GEMM + [c_vp]),
    "jblas_b200_gemm_f32_dev": (c_int, _GEMM + [c_vp]),
    "jblas_b200_gemm_plus_c_f64_dev": (c_int, _FUSED + [c_vp]),
    "jblas_b200_gemm_plus_c_f32_dev": (c_int, _FUSED + [c_vp]),
    "jblas_b200_gemm_x_plus_c_f64_dev": (c_int, _FUSED + [c_vp]),
    "jblas_b200_gemm_x_plus_c_f32_dev": (c_int, _FUSED + [c_vp]),
    "jblas_b200_gemm_plus_c_f64": (c_int, _FUSED),
    "jblas_b200_gemm_plus_c_f32": (c_int, _FUSED),
    "jblas_b200_gemm_x_plus_c_f64": (c_int, _FUSED),
    "jblas_b200_gemm_x_plus_c_f32": (c_int, _FUSED),
    "jblas_b200_fastmul_batched_f64": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64]),
    "jblas_b200_fastmul_batched_f32": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64]),
    "jblas_b200_fastmul_batched_f64_dev": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp]),
    "jblas_b200_fastmul_batched_f32_dev": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp]),
    "jblas_b200_alloc": (c_int, [ctypes.POINTER(c_vp), ctypes.c_size_t]),
    "jblas_b200_free": (c_int, [c_vp]),
    "jblas_b200_h2d": (c_int, [c_vp, c_vp, ctypes.c_size_t]),
    "jblas_b200_d2h": (c_int, [c_vp, c_vp, ctypes.c_size_t]),
    "jblas_b200_host_register": (c_int, [c_vp, ctypes.c_size_t]),
    "jblas_b200_host_unregister": (c_int, [c_vp]),
    "jblas_b200_stream_sync": (c_int, [c_vp]),
    "jblas_b200_ipc_export": (c_int, [c_vp, c_vp, ctypes.POINTER(c_i64)]),
    "jblas_b200_ipc_open": (c_int, [c_vp, c_i64, ctypes.POINTER(c_vp)]),
    "jblas_b200_ipc_close": (c_int, [c_vp]),
    "jblas_b200_copy_async": (c_int, [c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "jblas_b200_randn_fill": (c_int, [c_vp, c_i64, c_i64, ctypes.c_uint64, c_int, c_vp]),
    "jblas_b200_plan": (c_int, [c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, ctypes.POINTER(c_i64)]),
    "jblas_b200_num_kernels": (c_int, []),
    "jblas_b200_kernel_name": (ctypes.c_char_p, [c_int]),
    "jblas_b200_launch_count": (c_i64, []),
    "jblas_b200_time_last_ms": (ctypes.c_float, []),
    "jblas_b200_probe_pipe": (c_int, [c_int, c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)]),
}

_lib = None


class JblasB200Error(RuntimeError):
    """A libjblas_b200 call returned a negative status (message from jblas_b200_last_error)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"jblas_b200 error {code}: {msg}")
        self.code = code


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: build it with `python -m jblas.jl_b200.build` "
                "(the CUDA library is the product; there is no CPU fallback)"
            )
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError here = header/library mismatch, which must be loud
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> int:
    if rc < 0:
        raise JblasB200Error(rc, lib().jblas_b200_last_error().decode(errors="replace"))
    return rc
