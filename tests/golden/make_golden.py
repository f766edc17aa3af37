"""Generate tests/golden/*.npz -- small known-answer vectors for the jmul! chain.

    python tests/golden/make_golden.py

The reference ships no golden vectors and cannot be run (no Julia), so these are NOT outputs of the
reference: they are outputs of oracle/structural_jmul.py, the literal emulation of the code `jmul!`
generates, evaluated in exact rational arithmetic (see its docstring for the file:line map).  Two kinds:
  * "tiled_*": jmul_structural with the AVX-512 machine constants (tile 40x5 / 80x5): interior computed by
    the emulated tile loops, remainder untouched (NaN) exactly as the reference leaves it;
  * "chain_*": every element by the scalar exact chain  d = A[i,1]*X[1,j]; d = fma(A[i,n], X[n,j], d)
    (src/gemm.jl:86,165), for ragged shapes incl. the reference's own script shapes (test/runtests.jl:103-179).
They pin the C oracle, the CPU baseline and (on the GPU box) the CUDA kernels to the same bits.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.structural_jmul import jmul_structural, _mul, _fma  # noqa: E402
from tests.helpers import randn_f, SEED_A, SEED_X  # noqa: E402


def exact_chain(A, X):
    M, K = A.shape
    _, N = X.shape
    st = A.dtype.itemsize
    D = np.empty((M, N), dtype=A.dtype, order="F")
    for j in range(N):
        for i in range(M):
            d = _mul(A[i, 0], X[0, j], st)
            for k in range(1, K):
                d = _fma(A[i, k], X[k, j], d, st)
            D[i, j] = d
    return D


TILED = [  # (M, K, N, dtype)  BLAS naming
    (80, 19, 10, np.float64),
    (43, 9, 7, np.float64),   # ragged: only 40x5 is touched
    (160, 24, 10, np.float32),
    (83, 11, 6, np.float32),  # ragged: 80x5 touched
]
CHAIN = [
    (16, 32, 14, np.float64),  # test/runtests.jl:103-107
    (32, 32, 6, np.float64),   # test/runtests.jl:166-170
    (23, 17, 7, np.float64),
    (16, 32, 14, np.float32),
    (9, 40, 5, np.float32),
]


def main():
    for idx, (M, K, N, dt) in enumerate(TILED):
        A = randn_f((M, K), dt, SEED_A + 10 * idx)
        X = randn_f((K, N), dt, SEED_X + 10 * idx)
        D = np.full((M, N), np.nan, dtype=dt, order="F")
        D, cov = jmul_structural(D, A, X)
        name = f"tiled_{np.dtype(dt).name}_{M}x{K}x{N}.npz"
        np.savez_compressed(os.path.join(HERE, name), A=A, X=X, D=D, covered=np.array(cov))
        print(name, cov)
    for idx, (M, K, N, dt) in enumerate(CHAIN):
        A = randn_f((M, K), dt, SEED_A + 1000 + 10 * idx)
        X = randn_f((K, N), dt, SEED_X + 1000 + 10 * idx)
        if idx == 2:  # exercise signed zeros / exact cancellation / subnormals in one vector
            A[0, :] = 0.0
            X[:, 0] = -np.abs(X[:, 0])
            A[1, :] = np.array([1.0, -1.0] * (K // 2) + [0.0] * (K % 2), dtype=dt)
            X[:, 1] = 1.0
            A[2, :] *= dt(1e-160) if dt == np.float64 else dt(1e-20)
            X[:, 2] *= dt(1e-160) if dt == np.float64 else dt(1e-20)
        D = exact_chain(A, X)
        name = f"chain_{np.dtype(dt).name}_{M}x{K}x{N}.npz"
        np.savez_compressed(os.path.join(HERE, name), A=A, X=X, D=D)
        print(name)


if __name__ == "__main__":
    main()
