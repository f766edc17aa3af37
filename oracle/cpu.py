"""ctypes loader for oracle/_build/liboracle.so (oracle_gemm.c + jmul_baseline.c).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  All matrices are numpy arrays in Fortran
(column-major) order, the layout of the reference's MMatrix storage (src/gemm.jl:309-311).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_i64 = ctypes.c_int64
_p = ctypes.c_void_p


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only; no reference sources involved)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_gemm.c", "jmul_baseline.c", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.oracle_gemm_f64.argtypes = [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, ctypes.c_int]
        L.oracle_gemm_f32.argtypes = L.oracle_gemm_f64.argtypes
        L.oracle_absgemm_f64.argtypes = [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p]
        L.oracle_gemm_f64_sampled.argtypes = [_p, _p, _p, _i64, _i64, _i64, _p, _p, _i64]
        L.oracle_gemm_f32_sampled.argtypes = L.oracle_gemm_f64_sampled.argtypes
        L.jmul_baseline_f64.argtypes = [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, ctypes.c_int, ctypes.c_int, _p]
        L.jmul_baseline_f32.argtypes = L.jmul_baseline_f64.argtypes
        L.jmul_baseline_tile.argtypes = [ctypes.c_int, _p, _p, _p]
        L.jmul_pick_kernel_size.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _p, _p, _p]
        L.jmul_pick_kernel_size.restype = None
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def _check_f(a: np.ndarray, name: str):
    if a.ndim != 2 or (a.size > 0 and a.shape[0] > 1 and a.strides[0] != a.itemsize):
        raise ValueError(f"{name} must be a 2-D column-major array with unit row stride")
    return a.strides[1] // a.itemsize if (a.shape[1] > 1 and a.size > 0) else max(a.shape[0], 1)


def oracle_gemm(A: np.ndarray, X: np.ndarray, D: np.ndarray | None = None, accumulate: bool = False) -> np.ndarray:
    """D = A@X (or D += A@X) with the reference's per-element chain: product, then ascending-k fma."""
    M, K = A.shape
    K2, N = X.shape
    assert K == K2 and A.dtype == X.dtype and A.dtype in (np.float64, np.float32)
    if D is None:
        assert not accumulate
        D = np.full((M, N), np.nan, dtype=A.dtype, order="F")
    assert D.shape == (M, N) and D.dtype == A.dtype
    lda, ldx, ldd = _check_f(A, "A"), _check_f(X, "X"), _check_f(D, "D")
    fn = lib().oracle_gemm_f64 if A.dtype == np.float64 else lib().oracle_gemm_f32
    rc = fn(_ptr(D), _ptr(A), _ptr(X), M, K, N, ldd, lda, max(ldx, 1), int(accumulate))
    if rc:
        raise RuntimeError(f"oracle_gemm failed: {rc}")
    return D


def oracle_absgemm(A: np.ndarray, X: np.ndarray) -> np.ndarray:
    """(|A||X|)_ij in float64 -- the factor of the 2*K*eps*(|A||B|) acceptance bound."""
    A = np.asfortranarray(A, dtype=np.float64)
    X = np.asfortranarray(X, dtype=np.float64)
    M, K = A.shape
    _, N = X.shape
    B = np.zeros((M, N), dtype=np.float64, order="F")
    if K:
        sa, sx = np.empty(M * K), np.empty(K * N)
        rc = lib().oracle_absgemm_f64(_ptr(B), _ptr(A), _ptr(X), M, K, N, max(M, 1), max(M, 1), max(K, 1), _ptr(sa), _ptr(sx))
        if rc:
            raise RuntimeError(f"oracle_absgemm failed: {rc}")
    return B


def oracle_gemm_sampled(A: np.ndarray, X: np.ndarray, rows, cols) -> np.ndarray:
    """Chain values for the listed (row, col) pairs only (for sizes a CPU cannot finish in full)."""
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    cols = np.ascontiguousarray(cols, dtype=np.int64)
    out = np.empty(rows.shape[0], dtype=A.dtype)
    lda, ldx = _check_f(A, "A"), _check_f(X, "X")
    fn = lib().oracle_gemm_f64_sampled if A.dtype == np.float64 else lib().oracle_gemm_f32_sampled
    fn(_ptr(out), _ptr(A), _ptr(X), A.shape[1], lda, ldx, _ptr(rows), _ptr(cols), rows.shape[0])
    return out


def error_bound_ok(D, D_ref, A, X, extra_rel: float = 0.0):
    """Check |D - D_ref|_ij <= (2*K*eps(T) + extra_rel) * (|A||X|)_ij for every element.

    Returns (ok, worst_ratio) where worst_ratio = max_ij err_ij / bound_ij (<= 1 passes)."""
    K = A.shape[1]
    eps = 2.0 ** -52 if D_ref.dtype == np.float64 else 2.0 ** -23
    bound = (2.0 * K * eps + extra_rel) * oracle_absgemm(A, X)
    err = np.abs(D.astype(np.float64) - D_ref.astype(np.float64))
    if not np.all(np.isfinite(D)):
        return False, float("inf")
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(bound > 0, err / bound, np.where(err == 0, 0.0, np.inf))
    worst = float(ratio.max()) if ratio.size else 0.0
    return worst <= 1.0, worst


def pick_kernel_size(t_size: int, register_size: int, register_count: int):
    """src/kernel_structure.jl:76-99 restated: (vector_length, rows, cols)."""
    v, r, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib().jmul_pick_kernel_size(t_size, register_size, register_count, ctypes.byref(v), ctypes.byref(r), ctypes.byref(c))
    return v.value, r.value, c.value


def jmul_baseline_tile(t_size: int):
    v, r, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib().jmul_baseline_tile(t_size, ctypes.byref(v), ctypes.byref(r), ctypes.byref(c))
    if rc:
        raise RuntimeError("host CPU has neither AVX-512 nor AVX2+FMA (deps/build.jl would throw too)")
    return v.value, r.value, c.value


def jmul_baseline(D, A, X, col_tiles=None, nthreads: int = 1, fill_edges: bool = False):
    """The reference's jmul! loop nest (tile from pick_kernel_size, cc outer / rc inner, edges skipped).

    Dense column-major only (leading dimension = rows), as MMatrix.  Returns (rows_covered, cols_covered)."""
    M, K = A.shape
    _, N = X.shape
    for a, nm in ((A, "A"), (X, "X"), (D, "D")):
        if not a.flags.f_contiguous:
            raise ValueError(f"{nm} must be dense column-major")
    lo, hi = (0, -1) if col_tiles is None else col_tiles
    cov = (ctypes.c_int64 * 2)()
    fn = lib().jmul_baseline_f64 if A.dtype == np.float64 else lib().jmul_baseline_f32
    rc = fn(_ptr(D), _ptr(A), _ptr(X), M, K, N, lo, hi, nthreads, int(fill_edges), cov)
    if rc:
        raise RuntimeError(f"jmul_baseline failed: {rc}")
    return int(cov[0]), int(cov[1])


def fastmul_baseline_batched(D, A, X):
    """fastmul! (src/kernels.jl:43-130, 202-208) restated and applied to a batch: D[b] = A[b] @ X[b], float64.

    Arrays are (batch, rows, cols) views of column-major matrices, i.e. strides (stride_b, 1, rows) in elements, as the
    GPU batched entry takes them.  One thread (the reference is single-threaded)."""
    batch, M, N = A.shape
    _, _, P = X.shape
    assert D.shape == (batch, M, P) and X.shape[1] == N and A.dtype == np.float64 == X.dtype == D.dtype
    it = A.itemsize
    for a, r, nm in ((A, M, "A"), (X, N, "X"), (D, M, "D")):
        if a.size and (a.strides[1] != it or (a.shape[2] > 1 and a.strides[2] != r * it)):
            raise ValueError(f"{nm}: every matrix must be dense column-major")
    fn = lib().fastmul_baseline_batched_f64
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64] * 7
    rc = fn(D.ctypes.data, A.ctypes.data, X.ctypes.data, M, N, P, batch, D.strides[0] // it, A.strides[0] // it, X.strides[0] // it)
    if rc:
        raise RuntimeError(f"fastmul_baseline_batched failed: {rc}")
    return D


def num_threads() -> int:
    return lib().oracle_num_threads()
