"""GPU parity tests (-m gpu) of the fused forms the reference planned (src/memory_management.jl:72-76):
D = A*X + C and D = A*(X + C), through the C ABI, against the CPU oracle.

Oracle restatement: D = A*X + C is the accumulate chain started from C (oracle_gemm(..., accumulate=True) on a copy of
C); D = A*(X + C) is the overwrite chain on the elementwise IEEE sum X + C.  Exact kernels: bit-identical."""
import numpy as np
import pytest

import oracle
from tests.helpers import SEED_A, SEED_X, bits_equal, nan_f, randn_f, to_dev, to_host

pytestmark = pytest.mark.gpu

SHAPES = [(1, 1, 1), (16, 32, 14), (129, 17, 127), (300, 260, 200), (257, 64, 1000), (1000, 3, 5)]


def _selectors(jb, dt):
    names = jb.kernel_names()
    if dt == np.float64:
        return [jb.F64_SIMT, jb.F64_AUTO, jb.F64_DMMA] + [jb.EXPLICIT_BASE + i for i, n in enumerate(names) if n.startswith("dmma_f64")][-1:]
    return [jb.F32_EXACT] + [jb.EXPLICIT_BASE + i for i, n in enumerate(names) if n.startswith("simt_f32x2")][-2:]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_gemm_plus_c_device(jb, shape, dt):
    import torch

    M, K, N = shape
    A, X, C = randn_f((M, K), dt, SEED_A), randn_f((K, N), dt, SEED_X), randn_f((M, N), dt, 11, ld=M + 3)
    want = oracle.oracle_gemm(A, X, np.asfortranarray(C).copy(order="F"), accumulate=True)
    dA, dX, dC = to_dev(A), to_dev(X), to_dev(C)
    for sel in _selectors(jb, dt):
        dD = to_dev(nan_f((M, N), dt, ld=M + 1))
        assert jb.gemm_plus_c_(dD, dA, dX, dC, kernel=sel) is dD
        torch.cuda.synchronize()
        got = to_host(dD)
        if sel in (jb.F64_SIMT, jb.F32_EXACT) or dt == np.float32:
            assert bits_equal(got, want), (sel, shape)
        else:  # DMMA contract: tolerance (measured: identical bits)
            assert np.abs(got - want).max() <= 2 * K * 2.0 ** -52 * (np.abs(A) @ np.abs(X) + np.abs(C)).max()
        assert bits_equal(to_host(dC), np.asfortranarray(C))  # C is read-only
    # C aliasing D is kernel!'s D += A*X
    dD = to_dev(np.asfortranarray(C).copy(order="F"))
    jb.gemm_plus_c_(dD, dA, dX, dD, kernel=jb.F64_SIMT if dt == np.float64 else jb.F32_EXACT)
    torch.cuda.synchronize()
    assert bits_equal(to_host(dD), want)


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_gemm_x_plus_c_device(jb, shape, dt):
    import torch

    M, K, N = shape
    A, X, C = randn_f((M, K), dt, SEED_A), randn_f((K, N), dt, SEED_X, ld=K + 1), randn_f((K, N), dt, 12)
    Xs = np.asfortranarray(np.asfortranarray(X) + C)  # one IEEE rounding per element
    want = oracle.oracle_gemm(A, Xs)
    dA, dX, dC = to_dev(A), to_dev(X), to_dev(C)
    for sel in _selectors(jb, dt):
        dD = to_dev(nan_f((M, N), dt))
        jb.gemm_x_plus_c_(dD, dA, dX, dC, kernel=sel)
        torch.cuda.synchronize()
        got = to_host(dD)
        if sel in (jb.F64_SIMT, jb.F32_EXACT) or dt == np.float32:
            assert bits_equal(got, want), (sel, shape)
        else:
            ok, worst = oracle.error_bound_ok(got, want, A, Xs)
            assert ok, worst
        assert bits_equal(to_host(dX), np.asfortranarray(X))  # X is not modified


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_fused_forms_host_pointers_and_errors(jb, dt):
    M, K, N = 130, 77, 45
    A, X = randn_f((M, K), dt, SEED_A, ld=M + 2), randn_f((K, N), dt, SEED_X)
    C, Cx = randn_f((M, N), dt, 13), randn_f((K, N), dt, 14, ld=K + 5)
    exact = jb.F64_SIMT if dt == np.float64 else jb.F32_EXACT
    D = nan_f((M, N), dt, ld=M + 7)
    jb.gemm_plus_c_(D, A, X, C, kernel=exact)
    assert bits_equal(D, oracle.oracle_gemm(np.asfortranarray(A), X, C.copy(order="F"), accumulate=True))
    assert np.isnan(D.base[M:, :]).all()  # padding rows of the caller's D untouched
    D2 = nan_f((M, N), dt)
    jb.gemm_x_plus_c_(D2, A, X, Cx, kernel=exact)
    assert bits_equal(D2, oracle.oracle_gemm(np.asfortranarray(A), np.asfortranarray(X + np.asfortranarray(Cx))))
    with pytest.raises(ValueError):
        jb.gemm_plus_c_(D2, A, X, Cx)  # C must be M x N
    with pytest.raises(ValueError):
        jb.gemm_x_plus_c_(D2, A, X, C)  # C must be K x N
    with pytest.raises(TypeError):
        jb.gemm_plus_c_(D2, A, X, C.astype(np.float32 if dt == np.float64 else np.float64))


def test_fused_forms_empty_contraction(jb):
    """K = 0: D = C (plus_c) and D = 0 (x_plus_c), like the plain product's D = 0."""
    import torch

    M, N = 33, 9
    C = randn_f((M, N), np.float64, 15)
    dD = to_dev(nan_f((M, N)))
    jb.gemm_plus_c_(dD, to_dev(np.zeros((M, 0), order="F")), to_dev(np.zeros((0, N), order="F")), to_dev(C))
    torch.cuda.synchronize()
    assert bits_equal(to_host(dD), C)
    jb.gemm_x_plus_c_(dD, to_dev(np.zeros((M, 0), order="F")), to_dev(np.zeros((0, N), order="F")), to_dev(np.zeros((0, N), order="F")))
    torch.cuda.synchronize()
    assert (to_host(dD) == 0).all()


@pytest.mark.parametrize("shape", [(20000, 64, 64), (17001, 64, 40), (30000, 72, 48)], ids=lambda s: "x".join(map(str, s)))
def test_fused_forms_on_the_tall_skinny_kernels(jb, shape):
    """D = A*X + C and D = A*(X + C) on shapes AUTO gives to the tall-skinny kernels (team kernel for K = 64, the shared-memory
    variant for K = 72): C is a separate matrix with its own leading dimension, read as the start of every chain; bit-identical."""
    import torch

    M, K, N = shape
    assert "skinny" in jb.plan(M + (M & 1), K, N)["kernel"]
    A, X = randn_f((M, K), np.float64, SEED_A, ld=M + (M & 1)), randn_f((K, N), np.float64, SEED_X)
    Ad = np.asfortranarray(A)
    C = randn_f((M, N), np.float64, 11, ld=M + 4 - (M & 1) * 0 + (M & 1))
    dA, dX, dC = to_dev(A), to_dev(X), to_dev(C)
    dD = to_dev(nan_f((M, N), np.float64, ld=M + 2 + (M & 1)))
    jb.gemm_plus_c_(dD, dA, dX, dC)
    torch.cuda.synchronize()
    assert bits_equal(to_host(dD), oracle.oracle_gemm(Ad, X, np.asfortranarray(C).copy(order="F"), accumulate=True))
    assert bits_equal(to_host(dC), np.asfortranarray(C))
    C2 = randn_f((K, N), np.float64, 12)
    dC2 = to_dev(C2)
    dD2 = to_dev(nan_f((M, N), np.float64))
    jb.gemm_x_plus_c_(dD2, dA, dX, dC2)
    torch.cuda.synchronize()
    assert bits_equal(to_host(dD2), oracle.oracle_gemm(Ad, np.asfortranarray(X + C2)))
