"""CPU tests (-m "not gpu") of the N>1 host logic: world_size-2 gloo processes exercise the column partition,
the K-panel broadcast schedule and the accumulate ordering of jblas.jl_b200.multigpu.ShardedGemm.

The local multiply is injected (the CPU oracle) ONLY here: the product's default local_gemm is the CUDA path and
raises without a GPU.  What is under test is the plumbing, not the arithmetic."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from jblas.jl_b200.multigpu import ShardedGemm, column_shard, k_panels  # noqa: E402


def test_column_shard_partitions_exactly():
    for n, w in [(8192, 1), (8192, 8), (777, 2), (777, 4), (777, 8), (5, 8), (0, 4)]:
        spans = [column_shard(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes)  # remainder columns on the last ranks
    with pytest.raises(ValueError):
        column_shard(10, 2, 2)


def test_k_panels():
    assert k_panels(8192, 2048) == [(0, 2048), (2048, 4096), (4096, 6144), (6144, 8192)]
    assert k_panels(100, 64) == [(0, 64), (64, 100)]
    assert k_panels(64, 2048) == [(0, 64)]
    assert k_panels(8192, 2048, 256) == [(0, 256), (256, 2304), (2304, 4352), (4352, 6400), (6400, 8192)]
    assert k_panels(200, 2048, 256) == [(0, 200)]
    with pytest.raises(ValueError):
        k_panels(100, 48)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _colmajor(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a.T)).t()


def _worker(rank, world, port, M, K, N, panel_k, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from tests.helpers import SEED_A, SEED_X, randn_f

        calls = []

        def local_gemm(D, A, X, accumulate, kernel):  # test double for the CUDA kernel (same contract)
            calls.append((tuple(A.shape), tuple(X.shape), bool(accumulate)))
            Dn = np.asfortranarray(D.numpy())
            oracle.oracle_gemm(np.asfortranarray(A.numpy()), np.asfortranarray(X.numpy()), Dn, accumulate=accumulate)
            D.copy_(torch.from_numpy(Dn))

        sg = ShardedGemm(M, K, N, panel_k=panel_k, local_gemm=local_gemm)
        Xfull = randn_f((K, N), seed=SEED_X)
        Afull = randn_f((M, K), seed=SEED_A)
        A = _colmajor(Afull if rank == 0 else np.full((M, K), np.nan))  # only the root holds A
        Xs = _colmajor(Xfull[:, sg.c0:sg.c1])
        Ds = _colmajor(np.full((M, sg.shard_cols), np.nan))
        sg(Ds, A, Xs)
        assert np.array_equal(A.numpy(), Afull)  # the broadcast delivered every panel
        want = oracle.oracle_gemm(Afull, np.asfortranarray(Xfull[:, sg.c0:sg.c1]))
        assert np.asfortranarray(Ds.numpy()).tobytes() == want.tobytes()  # panel accumulate == single chain, bit for bit
        assert [c[2] for c in calls] == [False] + [True] * (len(sg.panels) - 1)
        assert sum(c[0][1] for c in calls) == K
        np.save(os.path.join(out_dir, f"d{rank}.npy"), Ds.numpy())
        np.save(os.path.join(out_dir, f"span{rank}.npy"), np.array([sg.c0, sg.c1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(96, 200, 37, 64), (64, 128, 64, 128)], ids=str)
def test_sharded_gemm_world2_gloo(tmp_path, shape):
    M, K, N, panel_k = shape
    mp.spawn(_worker, args=(2, _free_port(), M, K, N, panel_k, str(tmp_path)), nprocs=2, join=True)
    import oracle
    from tests.helpers import SEED_A, SEED_X, randn_f

    want = oracle.oracle_gemm(randn_f((M, K), seed=SEED_A), randn_f((K, N), seed=SEED_X))
    got = np.full((M, N), np.nan)
    for r in range(2):
        c0, c1 = np.load(tmp_path / f"span{r}.npy")
        got[:, c0:c1] = np.load(tmp_path / f"d{r}.npy")
    assert np.asfortranarray(got).tobytes() == want.tobytes()  # the shards tile D exactly; 2-rank == 1-rank bits


def test_default_local_gemm_is_the_cuda_path_and_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from jblas.jl_b200 import JblasB200Error
    from jblas.jl_b200 import build

    build.build()
    sg = ShardedGemm(4, 4, 4)
    D, A, X = (_colmajor(np.zeros((4, 4))) for _ in range(3))
    with pytest.raises((JblasB200Error, ValueError)):
        sg(D, A, X)
