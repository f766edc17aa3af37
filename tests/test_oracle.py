"""CPU tests (-m "not gpu"): the oracle against the golden vectors, the structural pin, the CPU baseline."""
import glob
import os

import numpy as np
import pytest

import oracle
from oracle.structural_jmul import jmul_structural, pick_kernel_size as py_pick
from tests.helpers import bits_equal, nan_f, randn_f, SEED_A, SEED_X

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def test_golden_files_present():
    assert len(GOLDEN) >= 9


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_golden(path):
    g = np.load(path)
    A, X, D = np.asfortranarray(g["A"]), np.asfortranarray(g["X"]), g["D"]
    Do = oracle.oracle_gemm(A, X)
    if "covered" in g:  # tiled_*: only the interior the reference's tile loops touch is defined
        r, c = (int(v) for v in g["covered"])
        assert bits_equal(Do[:r, :c], D[:r, :c])
        assert np.isnan(D[r:, :]).all() and np.isnan(D[:, c:]).all()
    else:
        assert bits_equal(Do, D)


@pytest.mark.parametrize("path", [p for p in GOLDEN if "tiled_" in p], ids=lambda p: os.path.basename(p))
def test_baseline_loop_nest_matches_golden_and_skips_edges(path):
    g = np.load(path)
    A, X, D = np.asfortranarray(g["A"]), np.asfortranarray(g["X"]), g["D"]
    Db = nan_f(D.shape, D.dtype)
    cov = oracle.jmul_baseline(Db, A, X)
    if oracle.jmul_baseline_tile(A.dtype.itemsize)[1:] == (40, 5) or oracle.jmul_baseline_tile(A.dtype.itemsize)[1:] == (80, 5):
        assert cov == tuple(int(v) for v in g["covered"])
        assert bits_equal(Db, D)  # including the untouched NaN remainder (src/gemm.jl:266-267,313)
    r, c = cov
    assert bits_equal(Db[:r, :c], oracle.oracle_gemm(A, X)[:r, :c])


def test_pick_kernel_size_table():
    # AVX-512 -> (8,40,5)/(16,80,5); AVX2 -> (4,12,4)/(8,24,4)
    for args, want in [((8, 64, 32), (8, 40, 5)), ((4, 64, 32), (16, 80, 5)), ((8, 32, 16), (4, 12, 4)), ((4, 32, 16), (8, 24, 4))]:
        assert oracle.pick_kernel_size(*args) == want
        assert py_pick(*args) == want


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("regs", [(64, 32), (32, 16)])
def test_structural_emulation_equals_oracle_interior(dt, regs):
    M, K, N = 97, 21, 11
    A, X = randn_f((M, K), dt, SEED_A), randn_f((K, N), dt, SEED_X)
    Ds, (r, c) = jmul_structural(nan_f((M, N), dt), A, X, register_size=regs[0], register_count=regs[1])
    assert r > 0 and c > 0
    Do = oracle.oracle_gemm(A, X)
    assert bits_equal(Ds[:r, :c], Do[:r, :c])  # the chain is tile-independent
    assert np.isnan(Ds[r:, :]).all() and np.isnan(Ds[:, c:]).all()


@pytest.mark.parametrize("shape", [(16, 32, 14), (32, 32, 28), (128, 128, 126), (800, 900, 840), (1, 1, 1), (5, 1, 3), (257, 513, 129)])
def test_oracle_within_bound_of_numpy(shape):
    M, K, N = shape  # first four: the reference script's shapes, test/runtests.jl:103-150 (jBLAS M x N x P)
    A, X = randn_f((M, K)), randn_f((K, N), seed=SEED_X)
    ok, worst = oracle.error_bound_ok(A @ X, oracle.oracle_gemm(A, X), A, X)
    assert ok, worst


def test_oracle_accumulate_is_kernel_semantics():
    M, K, N = 40, 13, 5
    A, X = randn_f((M, K)), randn_f((K, N), seed=SEED_X)
    D0 = randn_f((M, N), seed=7)
    D = D0.copy(order="F")
    oracle.oracle_gemm(A, X, D, accumulate=True)
    # kernel!: D loaded first, then fma for n = 0..N-1 (src/kernels.jl:225-233), i.e. the chain continued:
    # [A A] * [X; X] in one chain == A*X followed by an accumulate pass of A*X
    Dfull = oracle.oracle_gemm(np.asfortranarray(np.hstack([A, A])), np.asfortranarray(np.vstack([X, X])))
    D2 = oracle.oracle_gemm(A, X)
    oracle.oracle_gemm(A, X, D2, accumulate=True)
    assert bits_equal(D2, Dfull)
    assert not bits_equal(D, D0)


def test_oracle_strided_and_empty():
    A, X = randn_f((33, 7), ld=40), randn_f((7, 9), seed=SEED_X, ld=16)
    D = nan_f((33, 9), ld=50)
    oracle.oracle_gemm(A, X, D)
    assert bits_equal(D, oracle.oracle_gemm(np.asfortranarray(A), np.asfortranarray(X)))
    assert np.isnan(D.base[33:, :]).all()
    Z = oracle.oracle_gemm(np.zeros((4, 0), order="F"), np.zeros((0, 3), order="F"))
    assert (Z == 0).all()
    assert oracle.oracle_gemm(np.zeros((0, 5), order="F"), np.zeros((5, 3), order="F")).shape == (0, 3)


def test_sampled_oracle_equals_full():
    A, X = randn_f((64, 100)), randn_f((100, 48), seed=SEED_X)
    rows, cols = np.array([0, 5, 63, 17]), np.array([0, 47, 3, 3])
    full = oracle.oracle_gemm(A, X)
    assert bits_equal(oracle.oracle_gemm_sampled(A, X, rows, cols), full[rows, cols])


def test_baseline_threads_and_edges():
    M, K, N = 173, 64, 23
    for dt in (np.float64, np.float32):
        A, X = randn_f((M, K), dt), randn_f((K, N), dt, seed=SEED_X)
        Do = oracle.oracle_gemm(A, X)
        D1, D2 = nan_f((M, N), dt), nan_f((M, N), dt)
        oracle.jmul_baseline(D1, A, X, nthreads=1, fill_edges=True)
        oracle.jmul_baseline(D2, A, X, nthreads=4, fill_edges=True)
        assert bits_equal(D1, Do) and bits_equal(D2, Do)
        D3 = nan_f((M, N), dt)
        r, c = oracle.jmul_baseline(D3, A, X, col_tiles=(1, 3))
        _, _, cols = oracle.jmul_baseline_tile(np.dtype(dt).itemsize)
        assert c == 2 * cols and bits_equal(D3[:r, cols:3 * cols], Do[:r, cols:3 * cols]) and np.isnan(D3[:, :cols]).all()


@pytest.mark.parametrize("shape", [(16, 32, 14), (8, 5, 3), (16, 1, 14), (24, 9, 4), (5, 3, 7), (16, 32, 15), (32, 4, 2)], ids=str)
def test_fastmul_baseline_batched_equals_chain_oracle(shape):
    """The fastmul! restatement (register-resident D, src/kernels.jl:43-130) computes the same chain as oracle_gemm,
    for the specialised 16x32x14 kernel, the generic SIMD kernel and the scalar fallback."""
    M, N, P = shape
    batch = 5
    rng = np.random.Generator(np.random.PCG64(41))
    A = np.ascontiguousarray(rng.standard_normal((batch, N, M))).transpose(0, 2, 1)  # (batch, M, N), column-major matrices
    X = np.ascontiguousarray(rng.standard_normal((batch, P, N))).transpose(0, 2, 1)
    D = np.full((batch, P, M), np.nan).transpose(0, 2, 1)
    oracle.fastmul_baseline_batched(D, A, X)
    for b in range(batch):
        want = oracle.oracle_gemm(np.asfortranarray(A[b]), np.asfortranarray(X[b]))
        assert np.ascontiguousarray(D[b]).tobytes() == np.ascontiguousarray(want).tobytes()
