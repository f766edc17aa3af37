"""GPU tests (-m gpu): batched fastmul! -- many small independent products in one launch, each element the
reference chain, bit-identical to the oracle."""
import numpy as np
import pytest

import oracle
from tests.helpers import bits_equal

pytestmark = pytest.mark.gpu

# jBLAS names (M, N, P): D is MxP, A is MxN, X is NxP.  First entries: the reference script's small shapes
#   Float64 with M, P <= 32 runs the one-warp-per-product DMMA kernel (MI x NI fragment grids 1..4, k chunks of 32 or 16, k tails),
#   everything else the shared-memory SIMT kernel
SHAPES = [(16, 32, 14), (32, 32, 6), (32, 32, 24), (1, 1, 1), (5, 3, 7), (13, 40, 5), (40, 8, 5), (64, 64, 64), (3, 100, 2), (97, 11, 33),
          (25, 7, 31), (32, 70, 32), (8, 4, 8), (24, 33, 9), (9, 17, 24), (8, 16, 8), (20, 10, 130), (128, 5, 100),
          (96, 34, 70), (130, 6, 40), (66, 130, 34),  # Float64, even M and N, M or P > 32: the TMA/DMMA GEMM kernel over 3-D tensor maps
          # Float32 with even M <= 32, P <= 16, N % 4 == 0: the warp-private FFMA2 kernel (8 or 16 lanes per product, idle row lanes,
          # odd P, the largest product that still fits its shared-memory stages, and one that does not)
          (2, 4, 1), (16, 64, 16), (30, 8, 3), (32, 128, 16), (32, 256, 16), (18, 12, 15)]


def _to_np(t):
    return np.asfortranarray(t.cpu().numpy())


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("dt", ["float64", "float32"])
def test_fastmul_batched_bit_identical(jb, shape, dt):
    import torch

    M, N, P = shape
    for batch in (1, 7, 300):
        A = jb.mrandn_batch(batch, M, N, dt, seed=21)
        X = jb.mrandn_batch(batch, N, P, dt, seed=22)
        D = jb.empty_colmajor_batch(batch, M, P, dt, fill=float("nan"))
        assert jb.fastmul_batched_(D, A, X) is D
        torch.cuda.synchronize()
        assert not torch.isnan(D).any()
        for b in sorted({0, batch // 2, batch - 1}):
            want = oracle.oracle_gemm(_to_np(A[b]), _to_np(X[b]))
            assert bits_equal(_to_np(D[b]), want), (shape, dt, batch, b)


def test_fastmul_batched_strided_batch_and_special_values(jb):
    import torch

    batch, M, N, P = 10, 6, 9, 4
    # matrices embedded in a longer per-matrix stride (gaps between consecutive matrices must stay untouched)
    gap = 5
    storeD = torch.full((batch, M * P + gap), float("nan"), dtype=torch.float64, device="cuda")
    storeA = torch.randn((batch, M * N + gap), dtype=torch.float64, device="cuda")
    storeX = torch.randn((batch, N * P + gap), dtype=torch.float64, device="cuda")
    A = storeA[:, : M * N].unflatten(1, (N, M)).transpose(1, 2)
    X = storeX[:, : N * P].unflatten(1, (P, N)).transpose(1, 2)
    D = storeD[:, : M * P].unflatten(1, (P, M)).transpose(1, 2)
    A[0, 0, :] = 0.0
    X[0, :, 0] = -X[0, :, 0].abs()  # product of zeros with negatives: -0.0 must survive
    jb.fastmul_batched_(D, A, X)
    torch.cuda.synchronize()
    assert torch.isnan(storeD[:, M * P:]).all()
    for b in range(batch):
        assert bits_equal(_to_np(D[b]), oracle.oracle_gemm(_to_np(A[b]), _to_np(X[b])))
    d00 = D[0, 0, 0].item()
    assert d00 == 0.0 and np.signbit(d00)


def test_fastmul_batched_argument_errors(jb):
    import torch

    A = jb.mrandn_batch(4, 8, 8)
    X = jb.mrandn_batch(4, 8, 8)
    D = jb.empty_colmajor_batch(4, 8, 8)
    with pytest.raises(ValueError):
        jb.fastmul_batched_(D, A, X[:3])
    with pytest.raises(ValueError):
        jb.fastmul_batched_(D, A.transpose(1, 2), X)  # row-major matrices
    with pytest.raises(TypeError):
        jb.fastmul_batched_(D, A.float(), X)
    big = jb.empty_colmajor_batch(1, 257, 257)  # odd: no TMA path, and too big for the shared-memory kernel
    with pytest.raises(jb.JblasB200Error):
        jb.fastmul_batched_(big, big, big)  # too large for the small-matrix kernel: use gemm


def test_fastmul_batched_large_batch_throughput_shape(jb):
    """The reference's published shape (16x32x14, test/runtests.jl:110-122) at a batch that needs many waves."""
    import torch

    batch, M, N, P = 200_000, 16, 32, 14
    A = jb.mrandn_batch(batch, M, N, seed=31)
    X = jb.mrandn_batch(batch, N, P, seed=32)
    D = jb.empty_colmajor_batch(batch, M, P, fill=float("nan"))
    jb.fastmul_batched_(D, A, X)
    torch.cuda.synchronize()
    assert not torch.isnan(D).any()
    for b in (0, 1, 12345, batch - 1):
        assert bits_equal(_to_np(D[b]), oracle.oracle_gemm(_to_np(A[b]), _to_np(X[b])))
    # cross-check everything against torch.bmm within the reference tolerance
    ref = torch.bmm(A, X)
    bound = 2 * N * 2.0 ** -52 * torch.bmm(A.abs(), X.abs())
    assert ((D - ref).abs() <= bound).all()


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_fastmul_batched_host_pointer_pipeline(jb, dt):
    """Host-pointer entry (what a Julia ccall on an Array{T,3} hits): chunked H2D / kernel / D2H over three device slots;
    several chunks, a ragged last chunk, and gaps between matrices that must be neither read into D nor overwritten."""
    M, N, P = 16, 32, 14
    es = np.dtype(dt).itemsize
    batch = int(3.4 * (32 << 20) / ((M * N + N * P + M * P) * es))  # 3.4 chunks of 32 MiB
    rng = np.random.Generator(np.random.PCG64(77))
    gap = 6
    sA = np.ascontiguousarray(rng.standard_normal((batch, M * N + gap)).astype(dt))
    sX = np.ascontiguousarray(rng.standard_normal((batch, N * P)).astype(dt))
    sD = np.full((batch, M * P + gap), np.nan, dtype=dt)
    A = sA[:, : M * N].reshape(batch, N, M).transpose(0, 2, 1)
    X = sX.reshape(batch, P, N).transpose(0, 2, 1)
    D = sD[:, : M * P].reshape(batch, P, M).transpose(0, 2, 1)
    assert jb.fastmul_batched_(D, A, X) is D
    assert np.isnan(sD[:, M * P:]).all() and not np.isnan(sD[:, : M * P]).any()
    for b in sorted({0, 1, batch // 3, batch // 2, batch - 2, batch - 1}):
        want = oracle.oracle_gemm(np.asfortranarray(A[b]), np.asfortranarray(X[b]))
        assert bits_equal(np.asfortranarray(D[b]), want), b
    with pytest.raises(ValueError):
        jb.fastmul_batched_(D, A, X[:-1])
