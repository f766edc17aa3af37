"""GPU tests (-m gpu): the opt-in 3xTF32 tcgen05/TMEM Float32 path against the Float32 oracle.

Stated bound (north_star "a stated looser bound for 3xTF32"):
    |D - D_oracle32|_ij <= (2*K*2^-23 + 2^-18) * (|A||X|)_ij
The exact path stays the default; this one is only reached with kernel=F32_3XTF32."""
import json
import os

import numpy as np
import pytest

import oracle
from tests.helpers import SEED_A, SEED_X, nan_f, randn_f, to_dev, to_host

pytestmark = pytest.mark.gpu
EXTRA = 2.0 ** -18
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _tf32_selectors(jb):
    return [jb.F32_3XTF32] + [jb.EXPLICIT_BASE + i for i, n in enumerate(jb.kernel_names()) if n.startswith("tf32x3")]


def _run(jb, A, X, sel, acc_into=None, ldd=None):
    import torch

    M, N = A.shape[0], X.shape[1]
    Dh = nan_f((M, N), np.float32, ld=ldd) if acc_into is None else acc_into.copy(order="F")
    dD, dA, dX = to_dev(Dh), to_dev(A), to_dev(X)
    from jblas.jl_b200 import api

    api._gemm(dD, dA, dX, acc_into is not None, sel)
    torch.cuda.synchronize()
    return to_host(dD)


@pytest.mark.parametrize("shape", [(128, 32, 128), (128, 8, 256), (256, 256, 256), (300, 100, 270), (129, 33, 65), (1, 5, 1),
                                   (1024, 2048, 512), (640, 37, 900)], ids=lambda s: "x".join(map(str, s)))
def test_tf32x3_within_stated_bound(jb, shape):
    M, K, N = shape
    A, X = randn_f((M, K), np.float32, SEED_A), randn_f((K, N), np.float32, SEED_X)
    want = oracle.oracle_gemm(A, X)
    report = {}
    for sel in _tf32_selectors(jb):
        got = _run(jb, A, X, sel)
        assert not np.isnan(got).any(), "unwritten element"
        ok, worst = oracle.error_bound_ok(got, want, A, X, extra_rel=EXTRA)
        assert ok, (sel, worst)
        # how much of the bound is used, and the plain relative error against a float64 product
        ref64 = A.astype(np.float64) @ X.astype(np.float64)
        rel = np.abs(got - ref64).max() / np.abs(ref64).max()
        report[str(sel)] = {"worst_err_over_bound": worst, "max_rel_err_vs_f64": float(rel)}
    # a single TF32 pass would be ~2^-11 relative: make sure the split really buys Float32-level accuracy
    assert all(v["max_rel_err_vs_f64"] < 2e-5 for v in report.values()), report
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, f"tf32x3_err_{M}x{K}x{N}.json"), "w") as f:
        json.dump(report, f, indent=1)


def test_tf32x3_strided_and_accumulate(jb):
    M, K, N = 200, 96, 150
    A, X = randn_f((M, K), np.float32, ld=203), randn_f((K, N), np.float32, SEED_X, ld=101)
    Ad, Xd = np.asfortranarray(A), np.asfortranarray(X)
    want = oracle.oracle_gemm(Ad, Xd)
    got = _run(jb, A, X, jb.F32_3XTF32, ldd=211)
    ok, worst = oracle.error_bound_ok(got, want, Ad, Xd, extra_rel=EXTRA)
    assert ok, worst
    D0 = randn_f((M, N), np.float32, 5)
    got = _run(jb, A, X, jb.F32_3XTF32, acc_into=D0)
    want_acc = oracle.oracle_gemm(Ad, Xd, D0.copy(order="F"), accumulate=True)
    bound = (2 * K * 2.0 ** -23 + EXTRA) * (np.abs(Ad).astype(np.float64) @ np.abs(Xd).astype(np.float64) + np.abs(D0))
    assert (np.abs(got.astype(np.float64) - want_acc) <= bound).all()


@pytest.mark.parametrize("shape", [(256, 32, 256), (512, 96, 768), (300, 270, 100), (129, 40, 513), (1024, 4096, 1280)], ids=lambda s: "x".join(map(str, s)))
def test_tf32x3_cta_pair_kernel_equals_single_cta_kernel(jb, shape):
    """cta_group::2 (two CTAs per 256 x 256 tile, X shared across the pair) issues the same MMA sequence per element as the
    single-CTA kernel: lo*hi, hi*lo, hi*hi per 8 k, ascending k, one FP32 TMEM accumulator -- so every bit agrees, for
    overwrite and accumulate, including tiles that hang over the matrix edge in M (second CTA of a pair idle) and N."""
    M, K, N = shape
    names = jb.kernel_names()
    one = jb.EXPLICIT_BASE + names.index("tf32x3_tcgen05_f32_128x256x32_s2")
    pair = jb.EXPLICIT_BASE + names.index("tf32x3_tcgen05_2cta_f32_256x256x32_s3")
    A, X = randn_f((M, K), np.float32, SEED_A), randn_f((K, N), np.float32, SEED_X)
    a, b = _run(jb, A, X, one), _run(jb, A, X, pair)
    assert not np.isnan(b).any(), "unwritten element"
    assert a.tobytes() == b.tobytes()
    D0 = randn_f((M, N), np.float32, 5)
    assert _run(jb, A, X, one, acc_into=D0).tobytes() == _run(jb, A, X, pair, acc_into=D0).tobytes()


def test_tf32x3_auto_picks_the_pair_kernel_for_big_products(jb):
    assert jb.plan(16384, 16384, 16384, "float32", kernel=jb.F32_3XTF32)["kernel"] == "tf32x3_tcgen05_2cta_f32_256x256x32_s3"
    assert jb.plan(128, 512, 4096, "float32", kernel=jb.F32_3XTF32)["kernel"].startswith("tf32x3_tcgen05_f32_128x")  # one tile row: no pair


def test_tf32x3_host_pointer_entry(jb):
    M, K, N = 515, 260, 333
    A, X = randn_f((M, K), np.float32), randn_f((K, N), np.float32, SEED_X)
    D = nan_f((M, N), np.float32)
    jb.jmul_(D, A, X, kernel=jb.F32_3XTF32)
    ok, worst = oracle.error_bound_ok(D, oracle.oracle_gemm(A, X), A, X, extra_rel=EXTRA)
    assert ok, worst


def test_tf32x3_config_4096_cubed_sampled(jb):
    """A many-tile run (both TMEM accumulators and the whole smem ring cycle many times), checked on sampled rows x cols."""
    import torch

    n = 4096
    A = jb.mrandn(n, n, "float32", seed=SEED_A)
    X = jb.mrandn(n, n, "float32", seed=SEED_X)
    D = jb.empty_colmajor(n, n, "float32", fill=float("nan"))
    jb.jmul_(D, A, X, kernel=jb.F32_3XTF32)
    torch.cuda.synchronize()
    assert not torch.isnan(D).any()
    rng = np.random.Generator(np.random.PCG64(5))
    rows = np.unique(np.concatenate([[0, 1, n - 1, 127, 128], rng.integers(0, n, 59)]))
    cols = np.unique(np.concatenate([[0, 1, n - 1, 255, 256], rng.integers(0, n, 27)]))
    tr, tc = torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda()
    As = np.asfortranarray(A[tr, :].cpu().numpy())
    Xs = np.asfortranarray(X[:, tc].cpu().numpy())
    got = np.asfortranarray(D[tr][:, tc].cpu().numpy())
    ok, worst = oracle.error_bound_ok(got, oracle.oracle_gemm(As, Xs), As, Xs, extra_rel=EXTRA)
    assert ok, worst
