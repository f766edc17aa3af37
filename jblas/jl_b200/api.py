"""Host-side mirror of the reference's interface for the `jmul!` path, above the C ABI.

Julia is not installed in this environment, so this module is the executable host side; the Julia wrapper
that binds the same C ABI with `ccall` ships in julia/jBLASB200.jl (un-executed).
Names and argument meaning follow the reference (Python cannot spell `!`, so `jmul!` is `jmul_`):

    jmul_(D, A, X)        <- jmul!(D, A, X)                        src/gemm.jl:244     D = A*X
    gemm_(D, A, X)        <- BASELINE.json's name for the same call
    fastmul_(D, A, X)     <- fastmul!(D, A, X)                     src/kernels.jl:202  small matrices, any M
    kernel_(pD,pA,pX,K)   <- kernel!(pD, pA, pX, ::Kernel)         src/kernels.jl:239  D += A*X
    initkernel_(...)      <- initkernel!(pD, pA, pX, ::Kernel)     src/kernels.jl:273  D  = A*X
    Kernel(Mk,Pk,stride_AD,stride_X,N)  <- Kernel{...}             src/kernel_structure.jl:8-9
    mrandn(M, N)          <- mrandn(M, N)                          src/randmat.jl:11-14
    plan(...)             <- pick_kernel_size / blocking_structure src/kernel_structure.jl:76, memory_management.jl:78

Matrices are column-major with unit row stride (MMatrix storage): numpy arrays in Fortran order take the
host-pointer entry points (H2D/D2H inside the call, like a Julia caller); torch CUDA tensors with strides
(1, ld) take the device-pointer entry points on torch's current stream.  Output first, D returned, A and X
read-only, D must not alias A or X.  Size/type mismatches raise (the reference raises MethodError by dispatch).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import JblasB200Error, check  # noqa: F401

F64_AUTO, F64_DMMA, F64_SIMT = 0, 1, 2
F32_EXACT, F32_3XTF32 = 0, 1
EXPLICIT_BASE = 100
DT_F64, DT_F32 = 0, 1

_initialised_device = None


def init(device: int | None = None) -> int:
    """Bind this process to one GPU (one process per GPU).  Raises if no CUDA device exists."""
    global _initialised_device
    if device is None:
        device = 0
        try:
            import torch

            if torch.cuda.is_available():
                device = torch.cuda.current_device()
        except Exception:
            pass
    if _initialised_device == device:
        return device
    check(_lib.lib().jblas_b200_init(int(device)))
    _initialised_device = device
    return device


def shutdown() -> None:
    global _initialised_device
    check(_lib.lib().jblas_b200_shutdown())
    _initialised_device = None


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _describe(x, name: str):
    """-> (ptr, rows, cols, ld, dtype_tag, on_device)"""
    if _is_torch(x):
        import torch

        if x.dim() != 2:
            raise ValueError(f"{name} must be 2-D")
        if not x.is_cuda:
            raise ValueError(f"{name}: torch tensors must live on the GPU (use numpy arrays for host data)")
        if x.dtype not in (torch.float64, torch.float32):
            raise TypeError(f"{name}: dtype {x.dtype} not supported (Float64/Float32 only)")
        r, c = x.shape
        sr, sc = x.stride()
        if r > 1 and c > 0 and sr != 1:
            raise ValueError(f"{name} must be column-major with unit row stride (got strides {x.stride()})")
        ld = sc if (c > 1 and r > 0) else max(r, 1)
        if ld < max(r, 1):
            raise ValueError(f"{name}: leading dimension {ld} < rows {r}")
        return x.data_ptr(), r, c, ld, (DT_F64 if x.dtype == torch.float64 else DT_F32), True
    if not isinstance(x, np.ndarray):
        raise TypeError(f"{name} must be a numpy array or a torch CUDA tensor")
    if x.ndim != 2:
        raise ValueError(f"{name} must be 2-D")
    if x.dtype not in (np.float64, np.float32):
        raise TypeError(f"{name}: dtype {x.dtype} not supported (Float64/Float32 only)")
    r, c = x.shape
    if r > 1 and c > 0 and x.strides[0] != x.itemsize:
        raise ValueError(f"{name} must be column-major with unit row stride (use order='F')")
    ld = x.strides[1] // x.itemsize if (c > 1 and r > 0) else max(r, 1)
    if ld < max(r, 1):
        raise ValueError(f"{name}: leading dimension {ld} < rows {r}")
    return x.ctypes.data, r, c, ld, (DT_F64 if x.dtype == np.float64 else DT_F32), False


def mgpu_init(ngpus: int = 0) -> int:
    """Create the per-GPU contexts of the single-process multi-GPU mode on devices 0..ngpus-1 and enable peer access
    (ngpus <= 0: every visible GPU).  Returns the number of GPUs.  `jmul_(..., gpus=n)` calls it on demand."""
    global _initialised_device
    n = check(_lib.lib().jblas_b200_mgpu_init(int(ngpus)))
    if _initialised_device is None:
        _initialised_device = 0  # the library bound the process to device 0
    return n


class pinned:
    """Context manager: page-lock caller-owned numpy arrays for the duration of a block (jblas_b200_host_register), so the
    host-pointer entries DMA at full PCIe rate.  A Julia caller does the same once per buffer (julia/jBLASB200.jl: pin!)."""

    def __init__(self, *arrays):
        self.arrays = [a.base if (a.base is not None and isinstance(a.base, np.ndarray)) else a for a in arrays]
        self.done = []

    def __enter__(self):
        init()
        L = _lib.lib()
        try:
            for a in self.arrays:
                check(L.jblas_b200_host_register(a.ctypes.data, a.nbytes))
                self.done.append(a)
        except Exception:
            self.__exit__(None, None, None)
            raise
        return self

    def __exit__(self, *exc):
        L = _lib.lib()
        for a in self.done:
            L.jblas_b200_host_unregister(a.ctypes.data)
        self.done = []
        return False


def _gemm(D, A, X, accumulate: bool, selector: int | None, gpus: int | None = None):
    pD, M, N, ldd, tD, devD = _describe(D, "D")
    pA, M2, K, lda, tA, devA = _describe(A, "A")
    pX, K2, N2, ldx, tX, devX = _describe(X, "X")
    if not (tD == tA == tX):
        raise TypeError("D, A and X must share one element type")  # MethodError in the reference
    if M2 != M or K2 != K or N2 != N:
        raise ValueError(f"dimension mismatch: D is {M}x{N}, A is {M2}x{K}, X is {K2}x{N2}")
    if not (devD == devA == devX):
        raise ValueError("D, A and X must all be host arrays or all be GPU tensors")
    if not devD and not D.flags.writeable:
        raise ValueError("D must be writeable")
    L = _lib.lib()
    if selector is None:
        selector = F64_AUTO if tD == DT_F64 else F32_EXACT
    if gpus is not None and gpus != 1:
        # single-process multi-GPU mode: host matrices only (device-resident shards are the torch.distributed mode, multigpu.py)
        if devD:
            raise ValueError("gpus=N takes host (numpy) matrices; for GPU-resident shards use multigpu.ShardedGemm")
        mgpu_init(gpus)
        fn = L.jblas_b200_mgpu_gemm_f64 if tD == DT_F64 else L.jblas_b200_mgpu_gemm_f32
        check(fn(pD, pA, pX, M, K, N, ldd, lda, max(ldx, 1), int(accumulate), int(selector), int(gpus)))
        return D
    dev = init()
    if devD:
        import torch

        for name, t in (("D", D), ("A", A), ("X", X)):
            if t.device.index != dev:
                raise ValueError(f"{name} lives on cuda:{t.device.index} but this process is bound to cuda:{dev} (one process per GPU)")
        stream = torch.cuda.current_stream(D.device).cuda_stream
        fn = L.jblas_b200_gemm_f64_dev if tD == DT_F64 else L.jblas_b200_gemm_f32_dev
        check(fn(pD, pA, pX, M, K, N, ldd, lda, max(ldx, 1), int(accumulate), int(selector), stream))
    else:
        fn = L.jblas_b200_gemm_f64 if tD == DT_F64 else L.jblas_b200_gemm_f32
        check(fn(pD, pA, pX, M, K, N, ldd, lda, max(ldx, 1), int(accumulate), int(selector)))
    return D


def jmul_(D, A, X, Aprefetch=7, Xprefetch=7, A_loc=3, X_loc=3, D_loc=3, *, kernel: int | None = None, gpus: int | None = None):
    """D = A*X into the preallocated column-major D; returns D.  (jmul!, src/gemm.jl:244-348)

    The five prefetch arguments of the reference (`Val{Aprefetch_freq}` ..., src/gemm.jl:245) are accepted and
    ignored: software prefetch is replaced by the asynchronous shared-memory pipeline.  Unlike the reference,
    remainder rows/columns are computed.  `kernel` selects F64_AUTO/F64_DMMA/F64_SIMT (or F32_EXACT/F32_3XTF32).
    `gpus=n` (host matrices) runs the single-node multi-GPU mode: column blocks of X and D per GPU -- jmul!'s outer
    column-tile loop, src/gemm.jl:313 -- through jblas_b200_mgpu_gemm_*; the result equals the one-GPU result bit for bit."""
    del Aprefetch, Xprefetch, A_loc, X_loc, D_loc
    return _gemm(D, A, X, False, kernel, gpus)


gemm_ = jmul_  # BASELINE.json calls the same entry `gemm!`


def _fused(D, A, X, C, x_plus_c: bool, selector: int | None):
    pD, M, N, ldd, tD, devD = _describe(D, "D")
    pA, M2, K, lda, tA, devA = _describe(A, "A")
    pX, K2, N2, ldx, tX, devX = _describe(X, "X")
    pC, Cr, Cc, ldc, tC, devC = _describe(C, "C")
    if not (tD == tA == tX == tC):
        raise TypeError("D, A, X and C must share one element type")
    want_c = (K, N) if x_plus_c else (M, N)
    if M2 != M or K2 != K or N2 != N or (Cr, Cc) != want_c:
        raise ValueError(f"dimension mismatch: D is {M}x{N}, A is {M2}x{K}, X is {K2}x{N2}, C is {Cr}x{Cc} (expected {want_c[0]}x{want_c[1]})")
    if not (devD == devA == devX == devC):
        raise ValueError("D, A, X and C must all be host arrays or all be GPU tensors")
    if not devD and not D.flags.writeable:
        raise ValueError("D must be writeable")
    init()
    L = _lib.lib()
    if selector is None:
        selector = F64_AUTO if tD == DT_F64 else F32_EXACT
    name = "jblas_b200_gemm_" + ("x_plus_c" if x_plus_c else "plus_c") + ("_f64" if tD == DT_F64 else "_f32")
    if devD:
        import torch

        stream = torch.cuda.current_stream(D.device).cuda_stream
        check(getattr(L, name + "_dev")(pD, pA, pX, pC, M, K, N, ldd, lda, max(ldx, 1), max(ldc, 1), int(selector), stream))
    else:
        check(getattr(L, name)(pD, pA, pX, pC, M, K, N, ldd, lda, max(ldx, 1), max(ldc, 1), int(selector)))
    return D


def gemm_plus_c_(D, A, X, C, *, kernel: int | None = None):
    """D = A*X + C, the first fused form the reference planned (src/memory_management.jl:72-76).  Each element's fma
    chain starts from C[i,j] -- kernel!'s accumulate (src/kernels.jl:226) reading its start value from C; C may be D."""
    return _fused(D, A, X, C, False, kernel)


def gemm_x_plus_c_(D, A, X, C, *, kernel: int | None = None):
    """D = A*(X + C), the second planned fused form (src/memory_management.jl:72-76); C is K x N like X.  X + C is
    rounded once per element, then the jmul! chain."""
    return _fused(D, A, X, C, True, kernel)


def fastmul_(D, A, X):
    """fastmul!(D, A, X) (src/kernels.jl:202-208): D = A*X for small matrices, any row count (the reference masks
    the row remainder, src/kernels.jl:59-75).  Runs the exact (bit-identical chain) kernels."""
    t = _describe(D, "D")[4]
    return _gemm(D, A, X, False, F64_SIMT if t == DT_F64 else F32_EXACT)


@dataclass(frozen=True)
class Kernel:
    """Kernel{Mk,Pk,stride_AD,stride_X,N} (src/kernel_structure.jl:8-9): a tile product on raw storage.

    D tile is Mk x Pk, contraction length N; `stride_AD` is the column stride of BOTH A and D, `stride_X` the
    column stride of X, in elements (src/kernels.jl:213-215)."""

    Mk: int
    Pk: int
    stride_AD: int
    stride_X: int
    N: int


def _tile_views(pD, pA, pX, k: Kernel):
    """Column-major strided views of flat 1-D storage, the Python stand-in for Ptr{T} arithmetic."""

    def view(buf, rows, cols, ld, name):
        need = (cols - 1) * ld + rows if cols > 0 else 0
        if _is_torch(buf):
            if buf.dim() != 1 or buf.numel() < need:
                raise ValueError(f"{name}: need a 1-D buffer of at least {need} elements")
            return buf.as_strided((rows, cols), (1, ld))
        if buf.ndim != 1 or buf.size < need:
            raise ValueError(f"{name}: need a 1-D buffer of at least {need} elements")
        return np.lib.stride_tricks.as_strided(buf, shape=(rows, cols), strides=(buf.itemsize, ld * buf.itemsize))

    if k.stride_AD < k.Mk or k.stride_X < k.N:
        raise ValueError("stride_AD must be >= Mk and stride_X >= N")
    return (view(pD, k.Mk, k.Pk, k.stride_AD, "pD"), view(pA, k.Mk, k.N, k.stride_AD, "pA"),
            view(pX, k.N, k.Pk, k.stride_X, "pX"))


def kernel_(pD, pA, pX, k: Kernel):
    """kernel!(pD, pA, pX, K) (src/kernels.jl:212-241): D += A*X on raw storage (loads D first, :226).
    The reference throws unless Mk is a multiple of the vector width (:219); here any Mk is accepted."""
    D, A, X = _tile_views(pD, pA, pX, k)
    t = _describe(D, "D")[4]
    _gemm(D, A, X, True, F64_SIMT if t == DT_F64 else F32_EXACT)
    return None


def initkernel_(pD, pA, pX, k: Kernel):
    """initkernel!(pD, pA, pX, K) (src/kernels.jl:242-275): D = A*X on raw storage (first step is a plain product)."""
    D, A, X = _tile_views(pD, pA, pX, k)
    t = _describe(D, "D")[4]
    _gemm(D, A, X, False, F64_SIMT if t == DT_F64 else F32_EXACT)
    return None


def mrandn(M: int, N: int, dtype="float64", seed: int = 0x6A424C41, device=None, first_col: int = 0):
    """mrandn(M, N) (src/randmat.jl:11-14): an M x N column-major matrix of iid N(0,1) draws, generated on the GPU.

    The reference uses Julia's unseeded global RNG; here the stream is a counter-based Philox keyed by `seed`
    (default "jBLA"), so element i depends only on (seed, i).  Float32 matrices receive the Float64 draw rounded
    to Float32, as `x[i] = randn()` does.  `first_col` generates columns [first_col, first_col+N) of a wider matrix
    (a column shard holds exactly the values of the whole).  Returns a torch CUDA tensor with strides (1, M)."""
    import torch

    dev = init(device if isinstance(device, int) else None)
    tdt = {"float64": torch.float64, "float32": torch.float32}[str(dtype).replace("torch.", "")]
    store = torch.empty((N, M), dtype=tdt, device=f"cuda:{dev}")
    stream = torch.cuda.current_stream(store.device).cuda_stream
    check(_lib.lib().jblas_b200_randn_fill(store.data_ptr(), first_col * M, M * N, seed & (2**64 - 1),
                                           DT_F64 if tdt == torch.float64 else DT_F32, stream))
    return store.t()


def empty_colmajor(M: int, N: int, dtype="float64", device=None, fill=None):
    """Preallocated column-major M x N GPU matrix (strides (1, M)), optionally filled (NaN sentinel in tests)."""
    import torch

    dev = init(device if isinstance(device, int) else None)
    tdt = {"float64": torch.float64, "float32": torch.float32}[str(dtype).replace("torch.", "")]
    store = torch.empty((N, M), dtype=tdt, device=f"cuda:{dev}")
    if fill is not None:
        store.fill_(fill)
    return store.t()


def mrandn_batch(batch: int, M: int, N: int, dtype="float64", seed: int = 0x6A424C41, device=None):
    """`batch` column-major M x N matrices of iid N(0,1) draws: shape (batch, M, N), strides (M*N, 1, M)."""
    flat = mrandn(M * N, batch, dtype, seed, device)  # column b of the flat matrix is matrix b
    return flat.t().reshape(batch, N, M).transpose(1, 2)


def empty_colmajor_batch(batch: int, M: int, N: int, dtype="float64", device=None, fill=None):
    import torch

    dev = init(device if isinstance(device, int) else None)
    tdt = {"float64": torch.float64, "float32": torch.float32}[str(dtype).replace("torch.", "")]
    store = torch.empty((batch, N, M), dtype=tdt, device=f"cuda:{dev}")
    if fill is not None:
        store.fill_(fill)
    return store.transpose(1, 2)


def fastmul_batched_(D, A, X):
    """Batched fastmul! (src/kernels.jl:202-208 is the single-product form): D[b] = A[b] * X[b] for every b in one launch.

    D, A, X have shape (batch, M, P), (batch, M, N), (batch, N, P) and every matrix is dense column-major (strides
    (s, 1, rows) with s >= rows*cols).  torch CUDA tensors take the device entry on the current stream; numpy arrays take
    the host-pointer entry (chunked H2D / kernel / D2H pipeline, synchronous).  Exact chain per element (bit-identical to
    the oracle)."""
    dev = _is_torch(D)

    def desc(t, name):
        if dev:
            import torch

            if not _is_torch(t) or not t.is_cuda or t.dim() != 3:
                raise ValueError(f"{name} must be a 3-D torch CUDA tensor (batch, rows, cols), like D")
            if t.dtype not in (torch.float64, torch.float32):
                raise TypeError(f"{name}: dtype {t.dtype} not supported")
            strides, tag = t.stride(), (DT_F64 if t.dtype == torch.float64 else DT_F32)
        else:
            if not isinstance(t, np.ndarray) or t.ndim != 3:
                raise ValueError(f"{name} must be a 3-D numpy array (batch, rows, cols), like D")
            if t.dtype not in (np.float64, np.float32):
                raise TypeError(f"{name}: dtype {t.dtype} not supported")
            strides, tag = tuple(st // t.itemsize for st in t.strides), (DT_F64 if t.dtype == np.float64 else DT_F32)
        b, r, c = t.shape
        sb, sr, sc = strides
        if (r > 1 and sr != 1) or (c > 1 and sc != r):
            raise ValueError(f"{name}: every matrix must be dense column-major (strides (s, 1, rows)), got {strides}")
        return b, r, c, (sb if b > 1 else r * c), tag

    bD, M, P, sD, tD = desc(D, "D")
    bA, M2, N, sA, tA = desc(A, "A")
    bX, N2, P2, sX, tX = desc(X, "X")
    if not (tD == tA == tX):
        raise TypeError("D, A and X must share one element type")
    if not (bD == bA == bX) or M2 != M or N2 != N or P2 != P:
        raise ValueError(f"shape mismatch: D {tuple(D.shape)}, A {tuple(A.shape)}, X {tuple(X.shape)}")
    init()
    L = _lib.lib()
    if dev:
        import torch

        fn = L.jblas_b200_fastmul_batched_f64_dev if tD == DT_F64 else L.jblas_b200_fastmul_batched_f32_dev
        check(fn(D.data_ptr(), A.data_ptr(), X.data_ptr(), M, N, P, bD, sD, sA, sX, torch.cuda.current_stream(D.device).cuda_stream))
    else:
        if not D.flags.writeable:
            raise ValueError("D must be writeable")
        fn = L.jblas_b200_fastmul_batched_f64 if tD == DT_F64 else L.jblas_b200_fastmul_batched_f32
        check(fn(D.ctypes.data, A.ctypes.data, X.ctypes.data, M, N, P, bD, sD, sA, sX))
    return D


def plan(M: int, K: int, N: int, dtype="float64", kernel: int | None = None, ldd=None, lda=None, ldx=None) -> dict:
    """What the planner would launch for D(MxN) = A(MxK)*X(KxN): the B200 analogue of pick_kernel_size
    (src/kernel_structure.jl:76-99) and blocking_structure (src/memory_management.jl:78-140).  Pure host logic."""
    dt = DT_F64 if "64" in str(dtype) else DT_F32
    if kernel is None:
        kernel = F64_AUTO if dt == DT_F64 else F32_EXACT
    out = (ctypes.c_int64 * 10)()
    L = _lib.lib()
    check(L.jblas_b200_plan(dt, M, K, N, ldd or M, lda or M, ldx or K, kernel, out))
    name = L.jblas_b200_kernel_name(int(out[0])).decode()
    if name.startswith("tf32x3"):
        staging = "hi/lo TF32 split pre-pass into K-major scratch, then TMA cp.async.bulk.tensor, 128B swizzle"
    elif "tma" in name and out[9] == 0:  # misaligned operands on a persistent TMA-layout kernel: its ragged producers
        staging = "cp.async element-wise into the 128B-swizzled TMA layout (ragged producer warpgroup, no scratch copy)"
    else:
        staging = "TMA cp.async.bulk.tensor, 128B swizzle" if "tma" in name else ("cp.async 16B" if out[9] else "cp.async element-wise")
    if out[9] == 2 and not name.startswith("tf32x3"):
        staging += " (after re-aligning the ragged operand into scratch)"
    return {
        "kernel": name,
        "kernel_index": int(out[0]),
        "tile_m": int(out[1]), "tile_n": int(out[2]), "tile_k": int(out[3]),
        "stages": int(out[4]), "threads": int(out[5]), "grid": int(out[6]), "raster_group": int(out[7]),
        "smem_bytes": int(out[8]), "staging": staging,
    }


def kernel_names() -> list[str]:
    L = _lib.lib()
    return [L.jblas_b200_kernel_name(i).decode() for i in range(L.jblas_b200_num_kernels())]


def launch_count() -> int:
    return int(_lib.lib().jblas_b200_launch_count())


def probe_pipe(kind: str, iters: int = 20000):
    """Measured pipe rate in TFLOP/s for 'dfma' | 'dmma' | 'ffma' (register-only loop)."""
    init()
    tf, ms = ctypes.c_double(), ctypes.c_float()
    kinds = {"dfma": 0, "dmma": 1, "ffma": 2, "dmma_tile": 3, "dfma_tile": 4, "ffma_tile": 5, "ffma2_tile": 6}
    # 'ffma2_lds<mode>': the exact-FP32 kernel's inner loop with its shared-memory loads (mode bits: aux_kernels.cuh)
    if kind.startswith("ffma2_wide"):  # 8 x 16 thread tile: 0 = loads, 1 = loads + barrier, 2 = no loads, 3 = no loads + barrier
        k = 100 + int(kind[len("ffma2_wide"):] or 0)
    else:
        k = 7 + int(kind[len("ffma2_lds"):] or 0) if kind.startswith("ffma2_lds") else kinds[kind]
    check(_lib.lib().jblas_b200_probe_pipe(k, iters, ctypes.byref(tf), ctypes.byref(ms)))
    return tf.value, ms.value
