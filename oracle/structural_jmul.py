"""Line-by-line emulation of the code jBLAS.jl's `@generated jmul!` emits -- the oracle's pin.

TEST INFRASTRUCTURE ONLY.  Small cases only (pure-Python loops, exact rational arithmetic).

The reference cannot be executed here (no Julia) and has no golden vectors, so the fast C oracle
(oracle_gemm.c: "product, then ascending-k fma chain per element") is anchored on this slower,
more literal restatement, which follows the GENERATED STRUCTURE of the reference rather than its
mathematical meaning:

  * `pick_kernel_size`                                    src/kernel_structure.jl:76-99
  * `jmul!` generator body: divrem, row_loads, k split    src/gemm.jl:258-304
  * the emitted loops `for cc, rc { init; 2:Nr; Nd x cache_length; store }`   src/gemm.jl:313-336
  * `initialize_block`: vload of A at byte offset (r-1)*L*st + rc*rows*st, X[1, c+cc*cols],
    product                                               src/gemm.jl:71-91
  * `fma_increment_block`: vload at (r-1)*L*st - M*st + rc*rows*st + n*M*st, X[n, c+cc*cols],
    fma                                                   src/gemm.jl:149-170
  * `store_block`: vstore! at (r-1)*L*st + M*(c-1)*st + rc*rows*st + cc*cols*M*st
                                                          src/gemm.jl:3-11
  * remainder rows/cols never touched                     src/gemm.jl:266-267,340-345

Memory is a flat byte-addressed buffer exactly like `pointer(A)`; vectors are Python lists of L
lanes; `*` is a correctly-rounded product and `fma` a correctly-rounded fused multiply-add, both
evaluated exactly with `fractions.Fraction` and rounded once to the target binary format
(what SIMDPirates' vmul / llvm.fmuladd do on FMA hardware, src/kernels.jl:18,95).
"""
from __future__ import annotations

from fractions import Fraction
import math
import numpy as np

_FMT = {8: (53, -1022, 1023, np.float64), 4: (24, -126, 127, np.float32)}


def round_to_format(x: Fraction, t_size: int):
    """Round an exact rational to binary64/binary32, round-to-nearest-even, with subnormals."""
    p, emin, emax, ty = _FMT[t_size]
    if x == 0:
        return ty(0.0)
    sign = -1 if x < 0 else 1
    ax = abs(x)
    e = ax.numerator.bit_length() - ax.denominator.bit_length()
    if Fraction(2) ** e > ax:
        e -= 1
    if Fraction(2) ** (e + 1) <= ax:
        e += 1
    e = max(e, emin)
    q = ax / (Fraction(2) ** (e - p + 1))  # significand as a rational, target integer in [0, 2^p]
    n = q.numerator // q.denominator
    rem = q - n
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (n & 1)):
        n += 1
    val = sign * n * (Fraction(2) ** (e - p + 1))
    if abs(val) >= Fraction(2) ** (emax + 1):
        return ty(sign * math.inf)
    return ty(float(val))  # exact: val is representable in the target format


def _neg(x) -> bool:
    return bool(np.signbit(x))


def _mul(a, b, st):
    """Correctly rounded a*b (finite inputs), with IEEE signed-zero semantics."""
    ty = _FMT[st][3]
    fa, fb = Fraction(float(a)), Fraction(float(b))
    if fa == 0 or fb == 0:
        return ty(-0.0) if (_neg(a) != _neg(b)) else ty(0.0)
    r = round_to_format(fa * fb, st)
    if r == 0:  # underflow to zero keeps the sign of the exact product
        return ty(-0.0) if (_neg(a) != _neg(b)) else ty(0.0)
    return r


def _fma(a, b, c, st):
    """Correctly rounded a*b + c with one rounding (finite inputs), IEEE signed zeros, round-to-nearest-even."""
    ty = _FMT[st][3]
    fa, fb, fc = Fraction(float(a)), Fraction(float(b)), Fraction(float(c))
    exact = fa * fb + fc
    pneg = _neg(a) != _neg(b)
    if exact == 0:
        if fa * fb == 0 and fc == 0:  # (+-0) + (+-0): same sign keeps it, otherwise +0
            return ty(-0.0) if (pneg and _neg(c)) else ty(0.0)
        return ty(0.0)  # exact cancellation of non-zeros gives +0 under round-to-nearest
    r = round_to_format(exact, st)
    if r == 0:
        return ty(-0.0) if exact < 0 else ty(0.0)
    return r


def pick_kernel_size(t_size: int, register_size: int, register_count: int):
    """src/kernel_structure.jl:76-99."""
    elements_per_register = register_size // t_size
    cache_line = elements_per_register
    max_total = elements_per_register * register_count
    num_cache_lines = -(-max_total // cache_line)
    prev_num_rows, prev_num_cols = 0, 0
    prev_ratio = -math.inf
    for a_loads in range(1, num_cache_lines + 1):
        num_rows = a_loads * elements_per_register
        num_cols = (register_count - a_loads - 1) // a_loads
        length_D = num_rows * num_cols
        num_loads = num_cols + a_loads
        next_ratio = length_D / num_loads
        if next_ratio < prev_ratio:
            break
        prev_ratio = next_ratio
        prev_num_rows, prev_num_cols = num_rows, num_cols
    return elements_per_register, prev_num_rows, prev_num_cols


def jmul_structural(D: np.ndarray, A: np.ndarray, X: np.ndarray, register_size: int = 64, register_count: int = 32,
                    cacheline_size: int = 64):
    """Emulate jmul!(D, A, X) on dense column-major arrays; D's untouched remainder keeps its old contents.

    jBLAS naming inside: D is M x P, A is M x N, X is N x P (src/gemm.jl:244)."""
    M, N = A.shape
    N2, P = X.shape
    assert N == N2 and D.shape == (M, P) and A.dtype == X.dtype == D.dtype
    st = A.dtype.itemsize
    memA = np.asfortranarray(A).reshape(-1, order="F")  # pointer(A): flat, element index = byte offset / st
    memD = D.reshape(-1, order="F").copy()

    def vload(mem, byte_off, L):
        assert byte_off % st == 0
        i = byte_off // st
        assert 0 <= i and i + L <= mem.shape[0], "generated code would read out of bounds"
        return [mem[i + l] for l in range(L)]

    def vstore(mem, byte_off, v):
        i = byte_off // st
        assert 0 <= i and i + len(v) <= mem.shape[0]
        for l, x in enumerate(v):
            mem[i + l] = x

    L, rows, cols = pick_kernel_size(st, register_size, register_count)  # gemm.jl:258
    row_chunks, _row_remainder = divmod(M, rows)  # gemm.jl:266 (remainder unused, as in the reference)
    col_chunks, _col_remainder = divmod(P, cols)  # gemm.jl:267
    row_loads = rows // L  # gemm.jl:269
    cache_length = cacheline_size // st  # gemm.jl:282
    Nd, Nr = divmod(N - 1, cache_length)  # gemm.jl:299
    Nr += 1
    if Nr < cache_length // 3 and Nd > 0:  # gemm.jl:301-304
        Nr += cache_length
        Nd -= 1

    def fma_block(acc, rc, cc, n):  # gemm.jl:149-170 ; n is 1-based
        pA = [vload(memA, (r - 1) * L * st - M * st + rc * rows * st + n * M * st, L) for r in range(1, row_loads + 1)]
        for c in range(1, cols + 1):
            x = X[n - 1, c + cc * cols - 1]  # X[n, c + cc*cols]
            for r in range(1, row_loads + 1):
                acc[r, c] = [_fma(a, x, d, st) for a, d in zip(pA[r - 1], acc[r, c])]

    for cc in range(col_chunks):  # gemm.jl:313
        for rc in range(row_chunks):
            acc = {}
            # initialize_block, gemm.jl:71-91
            pA = [vload(memA, (r - 1) * L * st + rc * rows * st, L) for r in range(1, row_loads + 1)]
            for c in range(1, cols + 1):
                x = X[0, c + cc * cols - 1]
                for r in range(1, row_loads + 1):
                    acc[r, c] = [_mul(a, x, st) for a in pA[r - 1]]
            for n in range(2, Nr + 1):  # gemm.jl:319
                fma_block(acc, rc, cc, n)
            for nd in range(1, Nd + 1):  # gemm.jl:324
                pxn = (nd - 1) * cache_length + Nr
                for n in range(pxn + 1, pxn + cache_length + 1):
                    fma_block(acc, rc, cc, n)
            # store_block, gemm.jl:3-11
            for c in range(1, cols + 1):
                for r in range(1, row_loads + 1):
                    vstore(memD, ((r - 1) * L * st + M * (c - 1) * st) + rc * (rows * st) + cc * (cols * M * st), acc[r, c])
    D[...] = memD.reshape(M, P, order="F")
    return D, (row_chunks * rows, col_chunks * cols)
