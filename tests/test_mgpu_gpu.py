"""GPU tests (-m gpu) of the SINGLE-PROCESS multi-GPU entry, jblas_b200_mgpu_gemm_* (SURVEY 8b/8e): what a Julia
`jmul!(D, A, X; gpus = n)` hits.  Host matrices; GPU g owns a column block of X and D (jmul!'s outer column-tile loop,
src/gemm.jl:313), every K panel of A crosses PCIe once (one slice per GPU) and reaches the other GPUs over NVLink.

The pipeline with ONE GPU is the same code (no peer pulls), so its cases run on a 1-GPU box too; the n-GPU cases need n
visible devices and are skipped otherwise.  Bar: bit-identical to the oracle chain (exact kernels AND the DMMA path, which
is measured bit-identical on B200) and therefore to the one-GPU result."""
import numpy as np
import pytest

import oracle
from tests.helpers import SEED_A, SEED_X, bits_equal, nan_f, randn_f

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch

    return torch.cuda.device_count()


def _need(n):
    if _ngpu() < n:
        pytest.skip(f"needs {n} GPUs, {_ngpu()} visible")


GPUS = [1, 2, 4, 8]


@pytest.mark.parametrize("gpus", GPUS)
def test_mgpu_pipeline_shape_bit_identical(jb, gpus):
    """Big enough for several K panels, both phases and several column blocks per GPU; host matrices with ld > rows."""
    _need(gpus)
    M, K, N = 4096, 4608, 2304 * min(gpus, 2)
    A, X = randn_f((M, K), ld=M + 2), randn_f((K, N), seed=SEED_X)
    want = oracle.oracle_gemm(np.asfortranarray(A), X)
    for sel in (jb.F64_SIMT, jb.F64_AUTO):
        D = nan_f((M, N), ld=M + 6)
        out = jb.jmul_(D, A, X, kernel=sel, gpus=gpus)
        assert out is D and bits_equal(D, want), (gpus, sel)
        assert np.isnan(D.base[M:, :]).all()  # padding rows of the caller's D untouched


@pytest.mark.parametrize("gpus", GPUS)
def test_mgpu_ragged_config_and_accumulate(jb, gpus):
    """BASELINE configs[3] (1023 x 777 x 4097: odd everything, 777 columns over n GPUs) and kernel!-style accumulation."""
    _need(gpus)
    from jblas.jl_b200 import api

    M, N, K = 1023, 777, 4097
    A, X = randn_f((M, K)), randn_f((K, N), seed=SEED_X)
    want = oracle.oracle_gemm(A, X)
    for sel in (jb.F64_SIMT, jb.F64_AUTO):
        D = nan_f((M, N))
        jb.jmul_(D, A, X, kernel=sel, gpus=gpus)
        assert bits_equal(D, want), (gpus, sel)
    D0 = randn_f((M, N), seed=9)
    D1 = D0.copy(order="F")
    api._gemm(D1, A, X, True, jb.F64_AUTO, gpus)
    assert bits_equal(D1, oracle.oracle_gemm(A, X, D0.copy(order="F"), accumulate=True))


@pytest.mark.parametrize("gpus", GPUS)
def test_mgpu_float32_and_fewer_columns_than_gpus(jb, gpus):
    _need(gpus)
    M, K, N = 515, 1300, 3  # with 4 or 8 GPUs some own no column at all
    A, X = randn_f((M, K), np.float32), randn_f((K, N), np.float32, SEED_X)
    D = nan_f((M, N), np.float32)
    jb.jmul_(D, A, X, gpus=gpus)
    assert bits_equal(D, oracle.oracle_gemm(A, X))
    z = jb.jmul_(np.full((4, 3), np.nan, order="F"), np.zeros((4, 0), order="F"), np.zeros((0, 3), order="F"), gpus=gpus)
    assert (z == 0).all()  # empty contraction: zeros, as on one GPU


@pytest.mark.parametrize("gpus", [2, 4, 8])
def test_mgpu_equals_one_gpu_on_a_large_product(jb, gpus):
    """8192 x 4096 x (2048 per GPU), pinned host buffers: the n-GPU result equals the one-GPU host-pointer result bit for
    bit (and a sampled block equals the oracle).  This is the path bench.py's N > 1 `e2e` figure times."""
    _need(gpus)
    import torch

    M, K, N = 8192, 4096, 2048 * gpus
    A = np.asfortranarray(jb.mrandn(M, K, seed=SEED_A).cpu().numpy())
    X = np.asfortranarray(jb.mrandn(K, N, seed=SEED_X).cpu().numpy())
    torch.cuda.empty_cache()
    D1, Dn = nan_f((M, N)), nan_f((M, N))
    with jb.pinned(A, X, D1, Dn):
        jb.jmul_(D1, A, X)
        jb.jmul_(Dn, A, X, gpus=gpus)
    assert bits_equal(D1, Dn)
    rows = np.array([0, 1, 127, 128, 4095, 4096, M - 2, M - 1])
    cols = np.unique(np.concatenate([[0, 1, N - 1], np.arange(2048 - 2, N, 2048), np.arange(2048, N, 2048)]))  # shard seams
    want = oracle.oracle_gemm(np.asfortranarray(A[rows, :]), np.asfortranarray(X[:, cols]))
    assert bits_equal(np.asfortranarray(Dn[np.ix_(rows, cols)]), want)


def test_mgpu_rejects_more_gpus_than_visible(jb):
    D, A, X = nan_f((4, 5)), randn_f((4, 3)), randn_f((3, 5), seed=SEED_X)
    with pytest.raises(jb.JblasB200Error):
        jb.jmul_(D, A, X, gpus=_ngpu() + 1)
    assert np.isnan(D).all()
    import torch

    with pytest.raises(ValueError):  # device-resident shards belong to the torch.distributed mode
        jb.jmul_(jb.empty_colmajor(4, 5), torch.zeros(3, 4, device="cuda", dtype=torch.float64).t(), torch.zeros(5, 3, device="cuda", dtype=torch.float64).t(), gpus=2)


def test_tile_counters_are_per_stream_and_per_thread(jb):
    """two LIVE persistent TMA kernels must never share a tile-counter pair.  Pairs belong to streams (kernels of
    one stream never overlap): many launches on many streams and from several host threads (each with its own
    cudaStreamPerThread-like torch stream), all overlapping, must each give the single-launch result."""
    import threading

    import torch
    from jblas.jl_b200 import api
    from tests.helpers import to_dev, to_host

    M, K, N = 2080, 96, 1576
    A, X = randn_f((M, K)), randn_f((K, N), seed=SEED_X)
    dA, dX = to_dev(A), to_dev(X)
    want = oracle.oracle_gemm(A, X)
    sel = jb.EXPLICIT_BASE + jb.kernel_names().index("dmma_tma_f64_32x32x64_s3_x2")
    streams = [torch.cuda.Stream() for _ in range(24)]
    outs = [to_dev(nan_f((M, N))) for _ in range(48)]
    torch.cuda.synchronize()
    errors = []

    def worker(t):
        try:
            torch.cuda.set_device(0)
            for i in range(t, len(outs), 4):
                with torch.cuda.stream(streams[i % len(streams)]):
                    api._gemm(outs[i], dA, dX, False, sel)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    torch.cuda.synchronize()
    assert not errors, errors
    for dD in outs:
        assert bits_equal(to_host(dD), want)
