"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle.

Bar (north_star / SURVEY 8c):
  * exact kernels (SIMT DFMA / FFMA)  : BIT-IDENTICAL to the oracle chain (product, then ascending-k fma);
  * DMMA tensor path                  : |D - D_oracle|_ij <= 2*K*eps*(|A||X|)_ij, eps = 2^-52;
  * edges are computed (the reference skips them), D prefilled with a NaN sentinel catches unwritten elements.
"""
import glob
import json
import os

import numpy as np
import pytest

import oracle
from tests.helpers import SEED_A, SEED_X, bits_equal, nan_f, randn_f, to_dev, to_host

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _exact_selectors(jb, dt, A=None, X=None):
    """Selectors of every exact (chain) kernel of the dtype.  The TMA-fed Float32 kernel needs 16-byte aligned operands
    (leading dimensions multiples of 4): it is left out when the given host matrices do not qualify -- forcing it then
    is an error by contract (test_tma_kernels_refuse_unaligned_operands), AUTO never picks it there."""
    names = jb.kernel_names()
    tag = "simt_f64" if dt == np.float64 else "simt_f32"
    out = []
    for i, n in enumerate(names):
        if not n.startswith(tag):
            continue
        if "_tma_" in n and A is not None:
            lda = A.strides[1] // A.itemsize if A.shape[1] > 1 else A.shape[0]
            ldx = X.strides[1] // X.itemsize if X.shape[1] > 1 else X.shape[0]
            if lda % 4 or ldx % 4:
                continue
        out.append(jb.EXPLICIT_BASE + i)
    return out


def _dmma_selectors(jb, tma=True):
    """cp.async DMMA kernels and (tma=True) the TMA warp-specialised ones (those need 16-byte aligned operands)."""
    return [jb.EXPLICIT_BASE + i for i, n in enumerate(jb.kernel_names())
            if n.startswith("dmma_f64") or (tma and n.startswith("dmma_tma_f64"))]


# persistent TMA-layout kernels that also carry element-wise (ragged) producers: misaligned operands are legal for them
RAGGED_CAPABLE = {"dmma_tma_f64_96x64x32_s4_w8", "dmma_tma_f64_64x64x32_s3_x2", "dmma_tma_f64_128x64x32_s4_w8", "dmma_tma_f64_32x32x64_s3_x2",
                  "dmma_tma_f64_64x32x32_s4_x2"}


def _run_dev(jb, A, X, kernel, accumulate_into=None, ldd=None):
    import torch

    M, N = A.shape[0], X.shape[1]
    Dh = nan_f((M, N), A.dtype, ld=ldd) if accumulate_into is None else accumulate_into.copy(order="F")
    dD, dA, dX = to_dev(Dh), to_dev(A), to_dev(X)
    if accumulate_into is None:
        jb.jmul_(dD, dA, dX, kernel=kernel)
    else:
        from jblas.jl_b200 import api

        api._gemm(dD, dA, dX, True, kernel)
    torch.cuda.synchronize()
    return to_host(dD)


# ------------------------------------------------------------------------------------------------------
# golden vectors
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_golden_vectors(jb, path):
    g = np.load(path)
    A, X, D = np.asfortranarray(g["A"]), np.asfortranarray(g["X"]), g["D"]
    full = oracle.oracle_gemm(A, X)
    for sel in _exact_selectors(jb, A.dtype.type, A, X):
        got = _run_dev(jb, A, X, sel)
        if "covered" in g:
            r, c = (int(v) for v in g["covered"])
            assert bits_equal(got[:r, :c], D[:r, :c])
        else:
            assert bits_equal(got, D)
        assert bits_equal(got, full)  # remainder rows/cols are computed too (the reference leaves them untouched)
    Dh = nan_f(D.shape, D.dtype)
    jb.fastmul_(Dh, A, X)  # host-pointer entry (what a Julia ccall hits)
    assert bits_equal(Dh, full)


# ------------------------------------------------------------------------------------------------------
# exact kernels: bit-identical
# ------------------------------------------------------------------------------------------------------
SHAPES = [
    (1, 1, 1), (1, 7, 1), (3, 2, 5), (16, 32, 14), (32, 32, 28), (128, 128, 126), (800, 900, 840),  # ref script shapes
    (64, 16, 64), (128, 16, 128), (129, 17, 127), (257, 33, 65), (255, 15, 129), (256, 256, 256), (130, 1000, 70),
    (40, 3, 5), (1000, 1, 1000), (2, 3000, 3),
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_exact_kernels_bit_identical(jb, shape, dt):
    M, K, N = shape
    A, X = randn_f((M, K), dt, SEED_A), randn_f((K, N), dt, SEED_X)
    want = oracle.oracle_gemm(A, X)
    for sel in _exact_selectors(jb, dt, A, X):
        got = _run_dev(jb, A, X, sel)
        assert not np.isnan(got).any(), "unwritten element (NaN sentinel survived)"
        assert bits_equal(got, want), (jb.kernel_names()[sel - jb.EXPLICIT_BASE], shape)


@pytest.mark.parametrize("lds", [(130, 131, 40), (129, 136, 33), (200, 129, 64), (144, 160, 48)], ids=str)
@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_exact_kernels_strided_leading_dimensions(jb, lds, dt):
    """Sub-matrix views: ld > rows, odd and even (both staging paths), untouched padding stays NaN."""
    M, K, N = 129, 33, 37
    lda, ldd, ldx = lds
    A, X = randn_f((M, K), dt, SEED_A, ld=lda), randn_f((K, N), dt, SEED_X, ld=ldx)
    want = oracle.oracle_gemm(np.asfortranarray(A), np.asfortranarray(X))
    for sel in _exact_selectors(jb, dt, A, X):
        import torch

        Dh = nan_f((M, N), dt, ld=ldd)
        dD, dA, dX = to_dev(Dh), to_dev(A), to_dev(X)
        jb.jmul_(dD, dA, dX, kernel=sel)
        torch.cuda.synchronize()
        assert bits_equal(to_host(dD), want)
        full = dD._base.cpu().numpy()  # the (N, ldd) storage: padding rows M..ldd-1 of every column must be untouched
        assert full.shape == (N, ldd) and np.isnan(full[:, M:]).all()


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_accumulate_is_kernel_bang_semantics(jb, dt):
    """kernel!: D is loaded first and the chain continues (src/kernels.jl:226) -- bit-identical to the oracle."""
    M, K, N = 200, 77, 90
    A, X = randn_f((M, K), dt, SEED_A), randn_f((K, N), dt, SEED_X)
    D0 = randn_f((M, N), dt, 99)
    want = oracle.oracle_gemm(A, X, D0.copy(order="F"), accumulate=True)
    for sel in _exact_selectors(jb, dt, A, X):
        assert bits_equal(_run_dev(jb, A, X, sel, accumulate_into=D0), want)
    # split-K by accumulate passes == one pass (the property the multi-GPU K-panel pipeline relies on)
    one = oracle.oracle_gemm(A, X)
    part = _run_dev(jb, np.asfortranarray(A[:, :40]), np.asfortranarray(X[:40, :]), _exact_selectors(jb, dt)[0])
    two = _run_dev(jb, np.asfortranarray(A[:, 40:]), np.asfortranarray(X[40:, :]), _exact_selectors(jb, dt)[0], accumulate_into=part)
    assert bits_equal(two, one)


def test_special_values_follow_the_chain(jb):
    M, K, N = 70, 37, 9
    A, X = randn_f((M, K)), randn_f((K, N), seed=SEED_X)
    A[0, :] = 0.0
    X[:, 0] = -np.abs(X[:, 0])          # row 0 x col 0: every product is -0.0 -> result -0.0
    A[1, :] = np.tile([1.0, -1.0], K)[:K]
    X[:, 1] = 1.0                       # exact cancellation
    A[2, 5] = np.inf
    A[3, 6] = np.nan
    A[4, :] *= 1e-160
    X[:, 2] *= 1e-160                   # subnormal / underflowing products
    want = oracle.oracle_gemm(A, X)
    for sel in _exact_selectors(jb, np.float64, A, X):
        got = _run_dev(jb, A, X, sel)
        fin = np.isfinite(want)
        assert bits_equal(got[fin], want[fin])
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))
    assert np.signbit(want[0, 0]) and want[0, 0] == 0.0


# ------------------------------------------------------------------------------------------------------
# DMMA tensor path: tolerance contract, and a measurement of how close to the chain it is
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(16, 32, 14), (128, 128, 126), (129, 17, 127), (257, 33, 65), (256, 256, 256), (800, 900, 840), (130, 1000, 70)],
                         ids=lambda s: "x".join(map(str, s)))
def test_dmma_within_reference_tolerance(jb, shape):
    M, K, N = shape
    A, X = randn_f((M, K)), randn_f((K, N), seed=SEED_X)
    want = oracle.oracle_gemm(A, X)
    report = {}
    for sel in _dmma_selectors(jb):
        name = jb.kernel_names()[sel - jb.EXPLICIT_BASE]
        if "tma" in name and (M % 2 or K % 2) and name not in RAGGED_CAPABLE:  # TMA needs 16-byte aligned leading dimensions
            with pytest.raises(jb.JblasB200Error):
                _run_dev(jb, A, X, sel)
            continue
        got = _run_dev(jb, A, X, sel)
        assert not np.isnan(got).any()
        ok, worst = oracle.error_bound_ok(got, want, A, X)  # 2*K*2^-52*(|A||X|)
        assert ok, worst
        report[jb.kernel_names()[sel - jb.EXPLICIT_BASE]] = {
            "worst_err_over_bound": worst,
            "bit_identical_fraction": float((got.view(np.uint64) == want.view(np.uint64)).mean()),
        }
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, f"dmma_vs_chain_{M}x{K}x{N}.json"), "w") as f:
        json.dump(report, f, indent=1)


@pytest.mark.parametrize("shape", [(70, 37, 9), (72, 36, 10), (256, 130, 64), (128, 2, 128)], ids=lambda s: "x".join(map(str, s)))
def test_dmma_is_bit_identical_to_the_chain_on_b200(jb, shape):
    """MEASURED hardware property, asserted so a regression is loud: B200's DMMA.8x8x4 accumulates its 4 products as
    sequential single-rounded FMAs in ascending k, so the tensor path reproduces the reference chain bit for bit --
    including signed zeros, exact cancellation, subnormals and the zero-padded K tail (padded X side is -0.0)."""
    M, K, N = shape
    A, X = randn_f((M, K)), randn_f((K, N), seed=SEED_X)
    A[0, :] = 0.0
    X[:, 0] = -np.abs(X[:, 0])          # every product -0.0 -> result must stay -0.0
    A[1, :] = np.tile([1.0, -1.0], K)[:K]
    X[:, 1] = 1.0                       # exact cancellation (K even) -> +0.0
    A[4, :] *= 1e-160
    X[:, 2] *= 1e-160                   # subnormal / underflowing products
    want = oracle.oracle_gemm(A, X)
    assert np.signbit(want[0, 0]) and want[0, 0] == 0.0
    for sel in _dmma_selectors(jb):
        name = jb.kernel_names()[sel - jb.EXPLICIT_BASE]
        if "tma" in name and (M % 2 or K % 2) and name not in RAGGED_CAPABLE:
            continue
        got = _run_dev(jb, A, X, sel)
        assert bits_equal(got, want), name
    D0 = randn_f((M, N), seed=5)
    want_acc = oracle.oracle_gemm(A, X, D0.copy(order="F"), accumulate=True)
    for sel in _dmma_selectors(jb):
        name = jb.kernel_names()[sel - jb.EXPLICIT_BASE]
        if "tma" in name and (M % 2 or K % 2) and name not in RAGGED_CAPABLE:
            continue
        assert bits_equal(_run_dev(jb, A, X, sel, accumulate_into=D0), want_acc), name


def test_dmma_accumulate_and_strided(jb):
    M, K, N = 150, 50, 66
    A, X = randn_f((M, K), ld=151), randn_f((K, N), seed=SEED_X, ld=50)
    D0 = randn_f((M, N), seed=5)
    want = oracle.oracle_gemm(np.asfortranarray(A), np.asfortranarray(X), D0.copy(order="F"), accumulate=True)
    for sel in _dmma_selectors(jb, tma=False):
        got = _run_dev(jb, A, X, sel, accumulate_into=D0)
        assert np.abs(got - want).max() <= 2 * K * 2.0 ** -52 * (np.abs(A) @ np.abs(X) + np.abs(D0)).max()


@pytest.mark.parametrize("shape", [(150, 50, 66), (128, 16, 128), (1, 2, 1), (300, 1000, 260), (129, 36, 255)], ids=str)
def test_dmma_tma_kernels_even_strides_accumulate_edges(jb, shape):
    """The TMA warp-specialised kernels: ragged M/N/K (zero-filled boxes), ld > rows (even), accumulate."""
    M, K, N = shape
    lda, ldx, ldd = M + (M % 2) + 2, K + (K % 2) + 4, M + 3
    A, X = randn_f((M, K), ld=lda), randn_f((K, N), seed=SEED_X, ld=ldx)
    Ad, Xd = np.asfortranarray(A), np.asfortranarray(X)
    want = oracle.oracle_gemm(Ad, Xd)
    D0 = randn_f((M, N), seed=5)
    want_acc = oracle.oracle_gemm(Ad, Xd, D0.copy(order="F"), accumulate=True)
    tma = [s for s in _dmma_selectors(jb) if "tma" in jb.kernel_names()[s - jb.EXPLICIT_BASE]]
    assert tma
    for sel in tma:
        got = _run_dev(jb, A, X, sel, ldd=ldd)
        assert not np.isnan(got).any()
        ok, worst = oracle.error_bound_ok(got, want, Ad, Xd)
        assert ok, worst
        got = _run_dev(jb, A, X, sel, accumulate_into=D0)
        assert np.abs(got - want_acc).max() <= 2 * K * 2.0 ** -52 * (np.abs(Ad) @ np.abs(Xd) + np.abs(D0)).max()


@pytest.mark.parametrize("shape", [(1023, 4097, 777), (129, 37, 255), (333, 64, 191), (2050, 129, 1031), (1, 3, 1)], ids=str)
def test_ragged_producers_bit_identical(jb, shape):
    """The element-wise (RAGGED) producers of the persistent kernels: odd row counts and leading dimensions, an odd base
    offset (a sub-matrix view starting at row 1), ragged M / N / K edges, overwrite and accumulate -- every bit of the chain
    (they write the same swizzled layout the TMA boxes would, so the consumers are the TMA kernels' own)."""
    import torch
    from jblas.jl_b200 import api

    M, K, N = shape
    A, X = randn_f((M, K), ld=M + 3), randn_f((K, N), seed=SEED_X, ld=K + 1 + (K % 2))  # lda odd or even+odd mix, ldx odd
    Ad, Xd = np.asfortranarray(A), np.asfortranarray(X)
    want = oracle.oracle_gemm(Ad, Xd)
    D0 = randn_f((M, N), seed=5)
    want_acc = oracle.oracle_gemm(Ad, Xd, D0.copy(order="F"), accumulate=True)
    for i, name in enumerate(jb.kernel_names()):
        if name not in RAGGED_CAPABLE:
            continue
        sel = jb.EXPLICIT_BASE + i
        assert bits_equal(_run_dev(jb, A, X, sel, ldd=M + 1), want), name
        assert bits_equal(_run_dev(jb, A, X, sel, accumulate_into=D0), want_acc), name
    if M > 2:  # a view that starts one row down: the BASE is off the 16-byte grid even where the leading dimension is even
        Ap = randn_f((M + 1, K), ld=M + 3 + ((M + 3) % 2))
        dA = to_dev(Ap)[1:, :]
        dX, dD = to_dev(X), to_dev(nan_f((M, N)))
        api._gemm(dD, dA, dX, False, jb.F64_AUTO)
        torch.cuda.synchronize()
        assert bits_equal(to_host(dD), oracle.oracle_gemm(np.asfortranarray(Ap[1:, :]), Xd))


def test_dmma_tma_dynamic_tile_scheduler_many_tiles_repeated_and_concurrent(jb):
    """The persistent TMA kernels draw tiles from a self-resetting global counter: many more tiles than resident CTAs,
    back-to-back launches (the counter must be zero again each time) and launches in flight on two streams at once
    (separate counter slots) must all give the single-launch chain result -- bit-identical to the exact SIMT kernel."""
    import torch
    from jblas.jl_b200 import api

    M, K, N = 2080, 96, 1576  # 65 x 50 = 3250 tiles of 32x32, ragged in M and N
    A, X = randn_f((M, K)), randn_f((K, N), seed=SEED_X)
    dA, dX = to_dev(A), to_dev(X)
    want = _run_dev(jb, A, X, jb.F64_SIMT)
    assert bits_equal(want, oracle.oracle_gemm(A, X))
    tma = [s for s in _dmma_selectors(jb) if "tma" in jb.kernel_names()[s - jb.EXPLICIT_BASE]]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for sel in tma:
        outs = [to_dev(nan_f((M, N))) for _ in range(6)]
        torch.cuda.synchronize()
        for i, dD in enumerate(outs):  # alternate streams: consecutive launches overlap on the device
            with torch.cuda.stream(s1 if i % 2 == 0 else s2):
                api._gemm(dD, dA, dX, False, sel)
        torch.cuda.synchronize()
        for dD in outs:
            assert bits_equal(to_host(dD), want), jb.kernel_names()[sel - jb.EXPLICIT_BASE]


# ------------------------------------------------------------------------------------------------------
# the reference-facing API: host pointers, kernel!/initkernel!/fastmul!, degenerate sizes
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_host_pointer_entry_matches_oracle(jb, dt):
    M, K, N = 333, 129, 77
    A, X = randn_f((M, K), dt, ld=340), randn_f((K, N), dt, SEED_X)
    want = oracle.oracle_gemm(np.asfortranarray(A), X)
    D = nan_f((M, N), dt, ld=400)
    out = jb.jmul_(D, A, X, kernel=jb.F64_SIMT if dt == np.float64 else jb.F32_EXACT)
    assert out is D and bits_equal(D, want) and np.isnan(D.base[M:, :]).all()
    D2 = nan_f((M, N), dt)
    jb.gemm_(D2, A, X)  # AUTO selector: tolerance contract
    ok, worst = oracle.error_bound_ok(D2, want, np.asfortranarray(A), X)
    assert ok, worst


def test_host_pointer_k_panel_pipeline_is_bit_identical(jb):
    """Large operands are staged in K panels (accumulate passes) and the last panel is split into column blocks whose
    D2H overlaps the next block's multiply; per-element k order -- and every bit -- is unchanged."""
    M, K, N = 4096, 4608, 2304  # 3 K-panels of <= 2048, D = 72 MiB -> 4 column blocks
    A, X = randn_f((M, K)), randn_f((K, N), seed=SEED_X)
    want = oracle.oracle_gemm(A, X)
    D = nan_f((M, N))
    jb.jmul_(D, A, X, kernel=jb.F64_SIMT)
    assert bits_equal(D, want)
    D2 = nan_f((M, N), ld=M + 6)
    jb.jmul_(D2, randn_f((M, K), ld=M + 2), X)  # AUTO (DMMA/TMA) through the same pipeline, strided host matrices
    assert bits_equal(D2, want) and np.isnan(D2.base[M:, :]).all()
    D0 = randn_f((M, N), seed=9)
    D3 = D0.copy(order="F")
    from jblas.jl_b200 import api

    api._gemm(D3, A, X, True, jb.F64_SIMT)  # accumulate through the pipeline: old D uploaded first
    assert bits_equal(D3, oracle.oracle_gemm(A, X, D0.copy(order="F"), accumulate=True))


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_kernel_initkernel_fastmul(jb, dt):
    Mk, Pk, N, sAD, sX = 40, 5, 23, 48, 32  # Kernel{Mk,Pk,stride_AD,stride_X,N}
    k = jb.Kernel(Mk, Pk, sAD, sX, N)
    rng = np.random.Generator(np.random.PCG64(3))
    pA = rng.standard_normal(sAD * N).astype(dt)
    pX = rng.standard_normal(sX * Pk).astype(dt)
    pD = rng.standard_normal(sAD * Pk).astype(dt)
    A = np.asfortranarray(pA.reshape(N, sAD).T[:Mk, :])
    X = np.asfortranarray(pX.reshape(Pk, sX).T[:N, :])
    D0 = np.asfortranarray(pD.reshape(Pk, sAD).T[:Mk, :])
    pad0 = pD.reshape(Pk, sAD)[:, Mk:].copy()
    # initkernel!: D = A*X
    d1 = pD.copy()
    assert jb.initkernel_(d1, pA, pX, k) is None
    assert bits_equal(np.asfortranarray(d1.reshape(Pk, sAD).T[:Mk, :]), oracle.oracle_gemm(A, X))
    assert bits_equal(d1.reshape(Pk, sAD)[:, Mk:], pad0)  # storage between columns untouched
    # kernel!: D += A*X
    d2 = pD.copy()
    jb.kernel_(d2, pA, pX, k)
    assert bits_equal(np.asfortranarray(d2.reshape(Pk, sAD).T[:Mk, :]), oracle.oracle_gemm(A, X, D0.copy(order="F"), accumulate=True))
    # fastmul!: any M (the reference masks the row remainder)
    for M in (1, 7, 13, 40):
        Af, Xf = randn_f((M, N), dt, 11), randn_f((N, Pk), dt, 12)
        assert bits_equal(jb.fastmul_(nan_f((M, Pk), dt), Af, Xf), oracle.oracle_gemm(Af, Xf))


def test_degenerate_sizes(jb):
    import torch

    z = jb.jmul_(np.full((4, 3), np.nan, order="F"), np.zeros((4, 0), order="F"), np.zeros((0, 3), order="F"))
    assert (z == 0).all()  # empty contraction: defined as zeros
    jb.jmul_(np.zeros((0, 3), order="F"), np.zeros((0, 5), order="F"), np.zeros((5, 3), order="F"))
    jb.jmul_(np.zeros((4, 0), order="F"), np.zeros((4, 5), order="F"), np.zeros((5, 0), order="F"))
    dD = jb.empty_colmajor(4, 3, fill=float("nan"))
    jb.jmul_(dD, torch.zeros((0, 4), device="cuda", dtype=torch.float64).t(), torch.zeros((3, 0), device="cuda", dtype=torch.float64).t())
    torch.cuda.synchronize()
    assert (dD == 0).all()
    with pytest.raises(jb.JblasB200Error):
        from jblas.jl_b200 import _lib

        _lib.check(_lib.lib().jblas_b200_gemm_f64_dev(dD.data_ptr(), dD.data_ptr(), dD.data_ptr(), 4, 3, 3, 2, 4, 3, 0, 0, None))  # ldd < M


# ------------------------------------------------------------------------------------------------------
# BASELINE.json configs
# ------------------------------------------------------------------------------------------------------
def test_config_256_cubed_all_kernels(jb):
    A, X = randn_f((256, 256)), randn_f((256, 256), seed=SEED_X)
    want = oracle.oracle_gemm(A, X)
    for sel in _exact_selectors(jb, np.float64, A, X):
        assert bits_equal(_run_dev(jb, A, X, sel), want)
    for sel in _dmma_selectors(jb) + [jb.F64_AUTO, jb.F64_DMMA]:
        ok, worst = oracle.error_bound_ok(_run_dev(jb, A, X, sel), want, A, X)
        assert ok, worst
    assert bits_equal(_run_dev(jb, A, X, jb.F64_SIMT), want)


def test_config_ragged_1023x777x4097(jb):
    M, N, K = 1023, 777, 4097
    A, X = randn_f((M, K)), randn_f((K, N), seed=SEED_X)
    want = oracle.oracle_gemm(A, X)
    assert bits_equal(_run_dev(jb, A, X, jb.F64_SIMT), want)
    ok, worst = oracle.error_bound_ok(_run_dev(jb, A, X, jb.F64_DMMA), want, A, X)
    assert ok, worst
    D = nan_f((M, N))
    jb.jmul_(D, A, X, kernel=jb.F64_SIMT)  # host-pointer entry on the ragged shape
    assert bits_equal(D, want)


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_ragged_operands_are_realigned_in_one_pass(jb, dt):
    """>= 1 GFLOP with odd leading dimensions / odd row counts: A and X are re-aligned into scratch by one launch
    (realign2_kernel: row groups of 16 bytes, partial last group, 8-column patches with a partial last patch)."""
    M, K, N = 1021, 1027, 515  # 1.08 GFLOP; 1021 and 1027 are odd and not multiples of 4
    A, X = randn_f((M, K), dt, ld=M + 2), randn_f((K, N), dt, SEED_X, ld=K + 6)
    want = oracle.oracle_gemm(np.asfortranarray(A), np.asfortranarray(X))
    got = _run_dev(jb, A, X, jb.F64_SIMT if dt == np.float64 else jb.F32_EXACT)
    assert bits_equal(got, want)
    if dt == np.float64:
        assert "ragged producer" in jb.plan(M, K, N, lda=M + 2, ldx=K + 6)["staging"]  # f64 tensor path: staged inside the kernel
        assert "re-aligning" in jb.plan(M, K, N, lda=M + 2, ldx=K + 6, kernel=jb.F64_SIMT)["staging"]
        ok, worst = oracle.error_bound_ok(_run_dev(jb, A, X, jb.F64_AUTO), want, np.asfortranarray(A), np.asfortranarray(X))
        assert ok, worst


def test_config_tall_skinny_65536x64x64(jb):
    M, N, K = 65536, 64, 64
    A, X = randn_f((M, K)), randn_f((K, N), seed=SEED_X)
    want = oracle.oracle_gemm(A, X)
    assert bits_equal(_run_dev(jb, A, X, jb.F64_SIMT), want)
    assert jb.plan(M, K, N)["kernel"] == SKINNY_XREG  # AUTO: the dedicated tall-skinny kernel (X in registers, A streamed once; A fits in L2: private boxes)
    assert bits_equal(_run_dev(jb, A, X, jb.F64_AUTO), want)  # DMMA chains like the reference on B200: every bit


SKINNY = "dmma_skinny_f64_16x64_xres_w12"
SKINNY_XREG = "dmma_skinny_f64_16x16_xreg_w8"
SKINNY_TEAM = "dmma_skinny_f64_16x16_xreg_team_w16"


@pytest.mark.parametrize("shape", [(65536, 64, 64), (20000, 64, 48), (16385, 32, 33), (100, 64, 64), (5000, 32, 17), (3000, 64, 32), (1, 32, 1),
                                   (40000, 32, 64), (16 * 592 * 3 + 21, 64, 64)], ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("kernel_name", ["dmma_skinny_f64_16x16_xreg_w8", "dmma_skinny_f64_16x16_xreg_team_w16"])
def test_tall_skinny_xreg_kernel_bit_identical(jb, shape, kernel_name):
    """The X-in-registers variants (K = 64 or 32; private boxes per column half, or one box per row block shared by a team of four
    quarter-column warps and launched with programmatic stream serialisation) forced on shapes around their limits: N < 64 (zero X fragments, guarded stores),
    N <= 32 (one column half: every warp on half 0), odd M, fewer blocks than warps, a ragged last round; overwrite,
    accumulate and a D whose leading dimension breaks the 16-byte stores; X needs no alignment at all (odd ldx)."""
    M, K, N = shape
    sel = jb.EXPLICIT_BASE + jb.kernel_names().index(kernel_name)
    A, X = randn_f((M, K), ld=M + (M & 1)), randn_f((K, N), seed=SEED_X, ld=K + 1)
    Ad, Xd = np.asfortranarray(A), np.asfortranarray(X)
    want = oracle.oracle_gemm(Ad, Xd)
    assert bits_equal(_run_dev(jb, A, X, sel), want)
    assert bits_equal(_run_dev(jb, A, X, sel, ldd=M + 3 - (M & 1)), want)  # odd ldd: element-wise stores
    D0 = randn_f((M, N), seed=5)
    assert bits_equal(_run_dev(jb, A, X, sel, accumulate_into=D0), oracle.oracle_gemm(Ad, Xd, D0.copy(order="F"), accumulate=True))


@pytest.mark.parametrize("shape", [(65536, 64, 64), (20000, 72, 48), (16385, 8, 33), (100, 128, 64), (5000, 40, 17), (3000, 64, 32), (1, 8, 1),
                                   (40000, 128, 64), (16 * 1776 + 5, 64, 64)], ids=lambda s: "x".join(map(str, s)))
def test_tall_skinny_kernel_bit_identical(jb, shape):
    """gemm_skinny.cuh forced on shapes around its limits: N < 64 (zero-filled X columns, guarded stores), N <= 32 (the
    4-tile configuration), odd M (TMA zero-fills the missing rows of the last block, element-wise stores), K = 8 .. 128,
    fewer blocks than warps, a block count that leaves a ragged tail of half-block items; overwrite, accumulate
    (kernel! semantics) and a D whose leading dimension breaks the 16-byte stores."""
    M, K, N = shape
    sel = jb.EXPLICIT_BASE + jb.kernel_names().index(SKINNY)
    ldA = M + (M & 1)  # the TMA path needs an even leading dimension
    A, X = randn_f((M, K), ld=ldA), randn_f((K, N), seed=SEED_X)
    Ad = np.asfortranarray(A)
    want = oracle.oracle_gemm(Ad, X)
    assert bits_equal(_run_dev(jb, A, X, sel), want)
    assert bits_equal(_run_dev(jb, A, X, sel, ldd=M + 3 - (M & 1)), want)  # odd ldd: element-wise stores
    D0 = randn_f((M, N), seed=5)
    assert bits_equal(_run_dev(jb, A, X, sel, accumulate_into=D0), oracle.oracle_gemm(Ad, X, D0.copy(order="F"), accumulate=True))


def test_tall_skinny_kernel_limits(jb):
    sel = jb.EXPLICIT_BASE + jb.kernel_names().index(SKINNY)
    A, X = randn_f((4096, 36)), randn_f((36, 64), seed=SEED_X)  # K not a multiple of 8
    with pytest.raises(jb.JblasB200Error):
        _run_dev(jb, A, X, sel)
    A, X = randn_f((4096, 64)), randn_f((64, 72), seed=SEED_X)  # N > 64
    with pytest.raises(jb.JblasB200Error):
        _run_dev(jb, A, X, sel)
    # AUTO: only tall shapes with a short contraction take it; everything else stays on the tile kernels
    assert jb.plan(65536, 64, 64)["kernel"] == SKINNY_XREG and jb.plan(300000, 32, 40)["kernel"] == SKINNY_TEAM and jb.plan(1 << 20, 64, 64)["kernel"] == SKINNY_TEAM
    assert jb.plan(300000, 128, 40)["kernel"] == SKINNY and jb.plan(65536, 72, 48)["kernel"] == SKINNY
    assert jb.plan(65536, 64, 32)["kernel"] == SKINNY and jb.plan(262144, 32, 16)["kernel"] == SKINNY_TEAM  # 16 < N <= 32: shared-memory variant
    assert jb.plan(4096, 64, 64)["kernel"] != SKINNY and jb.plan(65536, 256, 64)["kernel"] != SKINNY and jb.plan(65536, 64, 128)["kernel"] != SKINNY
    assert jb.plan(65535, 64, 64)["kernel"] != SKINNY  # odd leading dimension: the ragged producers of the tile kernels


def _sampled_check(jb, n, dtype_name, selector, exact, extra_rel=0.0, tiles=()):
    """Full-size run on device-generated inputs; parity on a sampled sub-grid of rows x cols (each element's chain
    is independent) plus a NaN-sentinel sweep and a checksum property over the whole result."""
    import torch

    A = jb.mrandn(n, n, dtype_name, seed=SEED_A)
    X = jb.mrandn(n, n, dtype_name, seed=SEED_X)
    D = jb.empty_colmajor(n, n, dtype_name, fill=float("nan"))
    jb.jmul_(D, A, X, kernel=selector)
    torch.cuda.synchronize()
    assert not torch.isnan(D).any()
    rng = np.random.Generator(np.random.PCG64(5))
    rows = np.unique(np.concatenate([[0, 1, n - 1, n - 2, 127, 128], rng.integers(0, n, 58)]))
    cols = np.unique(np.concatenate([[0, 1, n - 1, n - 2, 127, 128], rng.integers(0, n, 26)]))
    tr, tc = torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda()
    As = np.asfortranarray(A[tr, :].cpu().numpy())      # the identical bits the GPU used
    Xs = np.asfortranarray(X[:, tc].cpu().numpy())
    got = np.asfortranarray(D[tr][:, tc].cpu().numpy())
    want = oracle.oracle_gemm(As, Xs)
    if exact:
        assert bits_equal(got, want)
    else:
        ok, worst = oracle.error_bound_ok(got, want, As, Xs, extra_rel=extra_rel)
        assert ok, worst
    # whole tiles, not only scattered elements: the 128 x 128 corner tile (matrix edge) and a 128 x 128 window that
    # straddles a tile seam in both directions (rows/columns 64..191 past a 128-multiple)
    for r0, c0 in tiles:
        As = np.asfortranarray(A[r0:r0 + 128, :].cpu().numpy())
        Xs = np.asfortranarray(X[:, c0:c0 + 128].cpu().numpy())
        got = np.asfortranarray(D[r0:r0 + 128, c0:c0 + 128].cpu().numpy())
        want = oracle.oracle_gemm(As, Xs)
        if exact:
            assert bits_equal(got, want), (r0, c0)
        else:
            ok, worst = oracle.error_bound_ok(got, want, As, Xs, extra_rel=extra_rel)
            assert ok, (worst, r0, c0)
    # size-independent property: D*1 == A*(X*1) within the reference bound scaled for the extra sum
    ones = torch.ones(n, 1, dtype=torch.float64, device="cuda")
    lhs = D.double() @ ones
    rhs = A.double() @ (X.double() @ ones)
    scale = (A.double().abs() @ (X.double().abs() @ ones))
    eps = 2.0 ** -52 if dtype_name == "float64" else 2.0 ** -23
    assert ((lhs - rhs).abs() <= (4 * n * eps + extra_rel) * scale).all()
    del A, X, D
    torch.cuda.empty_cache()


def test_config_8192_cubed_f64_simt_sampled_bit_identical(jb):
    _sampled_check(jb, 8192, "float64", jb.F64_SIMT, exact=True)


def test_config_8192_cubed_f64_dmma_sampled_bit_identical(jb):
    """The headline kernel (AUTO = dmma_tma_f64_128x128x32_s3).  Its contract is the tolerance; DMMA is MEASURED
    bit-identical to the chain on B200 (test_dmma_is_bit_identical_to_the_chain_on_b200), so the headline config is
    held to every bit: scattered sample + the corner tile + a window across a tile seam."""
    assert jb.plan(8192, 8192, 8192)["kernel"] == "dmma_tma_f64_128x128x32_s3"
    _sampled_check(jb, 8192, "float64", jb.F64_AUTO, exact=True, tiles=[(8192 - 128, 8192 - 128), (4032, 1984)])


def test_config_32768_cubed_f64_sampled_bit_identical(jb):
    """BASELINE configs[4] on ONE GPU (3 x 8 GiB resident): sampled parity against the oracle chain, corner tile, seam
    window, NaN-sentinel sweep and the checksum property.  The multi-GPU form of the same config is checked against
    this single-launch result in tests/test_multigpu_gpu.py and by bench.py's parity_check."""
    _sampled_check(jb, 32768, "float64", jb.F64_AUTO, exact=True, tiles=[(32768 - 128, 32768 - 128), (16320, 8128)])


def test_config_16384_cubed_f32_3xtf32_sampled_within_stated_bound(jb):
    """BASELINE configs[2], opt-in 3xTF32 path at FULL size: (2*K*2^-23 + 2^-18)*(|A||X|) on the sample and two tiles."""
    _sampled_check(jb, 16384, "float32", jb.F32_3XTF32, exact=False, extra_rel=2.0 ** -18,
                   tiles=[(16384 - 128, 16384 - 128), (8128, 4032)])


def test_config_16384_cubed_f32_exact_sampled_bit_identical(jb):
    _sampled_check(jb, 16384, "float32", jb.F32_EXACT, exact=True)


# ------------------------------------------------------------------------------------------------------
# mrandn
# ------------------------------------------------------------------------------------------------------
def test_mrandn_is_seeded_standard_normal(jb):
    import torch

    a = jb.mrandn(1000, 777)
    b = jb.mrandn(1000, 777)
    c = jb.mrandn(1000, 777, seed=SEED_X)
    assert a.shape == (1000, 777) and a.stride() == (1, 1000) and a.dtype == torch.float64
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs(a.mean().item()) < 5e-3 and abs(a.var().item() - 1.0) < 1e-2
    assert abs((a ** 4).mean().item() - 3.0) < 0.1 and a.abs().max().item() > 4.0
    f = jb.mrandn(1000, 777, "float32")
    assert f.dtype == torch.float32 and torch.equal(f, a.float())  # Float64 draw rounded to Float32 (src/randmat.jl:5-10)
    # element i depends only on (seed, i): a prefix of a longer stream is identical
    longer = jb.mrandn(1000, 1500)
    assert torch.equal(longer[:, :777], a)
    odd = jb.mrandn(3, 5)
    assert torch.isfinite(odd).all()


@pytest.mark.parametrize("shape", [(64, 64), (256, 256), (1023, 777), (16384, 64), (20001, 48), (130, 70)], ids=lambda s: "x".join(map(str, s)))
def test_back_to_back_dependent_products_under_programmatic_dependent_launch(jb, shape):
    """Every kernel family is launched with programmatic stream serialisation: the NEXT launch of a stream is scheduled while
    the current one runs and must not touch global memory before griddepcontrol.wait.  A chain of dependent products issued
    back to back with no synchronisation -- D1 = A*X1, D2 = D1*X2, D3 = D2*X3, ... where each launch READS what the previous
    one is still writing if the wait were misplaced -- must reproduce, stage by stage and bit for bit, the oracle applied to
    the GPU's own previous stage.  Repeated so that a race would have many chances."""
    import torch

    M, N = shape
    steps = 5
    A = randn_f((M, N), seed=SEED_A, ld=M + (M & 1))
    Xs = [randn_f((N, N), seed=SEED_X + i) * (1.0 / np.sqrt(N)) for i in range(steps)]
    dXs = [to_dev(np.asfortranarray(x)) for x in Xs]
    for rep in range(6):
        stages = [to_dev(A)] + [to_dev(nan_f((M, N), ld=M + (M & 1))) for _ in range(steps)]
        for i in range(steps):  # no synchronisation in between
            jb.gemm_(stages[i + 1], stages[i], dXs[i])
        torch.cuda.synchronize()
        host = [to_host(t) for t in stages]
        for i in range(steps):
            want = oracle.oracle_gemm(np.asfortranarray(host[i]), np.asfortranarray(Xs[i]))
            assert bits_equal(host[i + 1], want), (shape, rep, i, jb.plan(M, N, N, lda=M + (M & 1))["kernel"])
