"""GPU test (-m gpu, needs >= 2 GPUs, skipped otherwise): the column-sharded multi-GPU mode over NCCL.

Each rank owns a column block of X and D; A lives on rank 0 and is broadcast in K panels overlapped with the local
GEMM (accumulate passes).  With the exact kernels the concatenated shards must equal the single-chain oracle bit for
bit; with AUTO (DMMA) they must meet the reference tolerance."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, M, K, N, panel_k, out_dir):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import jblas.jl_b200 as jb
        from jblas.jl_b200.multigpu import ShardedGemm

        jb.init(rank)
        res = {}
        for tag, sel in (("simt", jb.F64_SIMT), ("auto", None)):
            sg = ShardedGemm(M, K, N, panel_k=panel_k, kernel=sel)
            A = jb.mrandn(M, K, seed=11) if rank == 0 else jb.empty_colmajor(M, K, fill=float("nan"))
            X = jb.mrandn(K, sg.shard_cols, seed=12, first_col=sg.c0)  # a column shard of the one global X
            D = jb.empty_colmajor(M, sg.shard_cols, fill=float("nan"))
            for _ in range(2):  # second call reuses the receive buffer (WAR ordering across calls)
                sg(D, A, X)
            torch.cuda.synchronize()
            assert not torch.isnan(A).any() and not torch.isnan(D).any()
            res[tag] = D.cpu().numpy()
            # the copy-engine transport: the owner's A is mapped over CUDA IPC and pulled in K panels, no NCCL data path
            sgp = ShardedGemm(M, K, N, panel_k=panel_k, kernel=sel, bcast="p2p")
            Ap = jb.mrandn(M, K, seed=11) if rank == 0 else jb.empty_colmajor(M, K, fill=float("nan"))
            Dp = jb.empty_colmajor(M, sg.shard_cols, fill=float("nan"))
            for _ in range(2):
                sgp(Dp, Ap, X)
            torch.cuda.synchronize()
            assert not torch.isnan(Ap).any() and np.array_equal(Dp.cpu().numpy(), res[tag]), f"p2p transport differs ({tag})"
            if rank == 0:
                Ap.mul_(2.0)  # the owner rewrites A in place between calls: the fences must order it against the peers' pulls
            sgp(Dp, Ap, X)
            torch.cuda.synchronize()
            assert np.array_equal(Dp.cpu().numpy(), 2.0 * res[tag]), f"p2p transport: stale A after an in-place update ({tag})"
            sgp.close()
            sga = ShardedGemm(M, K, N, panel_k=panel_k, kernel=sel, bcast="auto")  # collective choice, made at the first call
            if rank == 0:
                Ap.mul_(0.5)
            sga(Dp, Ap, X)
            torch.cuda.synchronize()
            assert sga.bcast in ("p2p", "nccl") and np.array_equal(Dp.cpu().numpy(), res[tag]), f"auto transport ({sga.bcast}, {tag})"
            sga.close()
            # host-facing pipelined form: pinned host shards in, pinned host shard out, same bits
            Xh = torch.empty((sg.shard_cols, K), dtype=torch.float64).pin_memory()
            Xh.copy_(X.t())
            Dh = torch.full((sg.shard_cols, M), float("nan"), dtype=torch.float64).pin_memory()
            Ah = None
            if rank == 0:
                Ah = torch.empty((K, M), dtype=torch.float64).pin_memory()
                Ah.copy_(A.t())
            A2 = jb.empty_colmajor(M, K, fill=float("nan"))
            D2 = jb.empty_colmajor(M, sg.shard_cols, fill=float("nan"))
            X2 = jb.empty_colmajor(K, sg.shard_cols, fill=float("nan"))
            sg.from_host(Dh, Ah, Xh, D2, A2, X2, nblocks=3)
            torch.cuda.synchronize()
            assert np.array_equal(Dh.numpy().T, res[tag]), f"from_host differs from the device-resident path ({tag})"
            if tag == "simt":
                np.save(os.path.join(out_dir, f"A{rank}.npy"), A.cpu().numpy())
                np.save(os.path.join(out_dir, f"X{rank}.npy"), X.cpu().numpy())
            np.save(os.path.join(out_dir, f"D_{tag}{rank}.npy"), res[tag])
        np.save(os.path.join(out_dir, f"span{rank}.npy"), np.array([sg.c0, sg.c1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(512, 1024, 771, 256), (256, 320, 256, 64)], ids=str)
def test_sharded_gemm_nccl(tmp_path, shape):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import oracle

    M, K, N, panel_k = shape
    mp.spawn(_worker, args=(world, _free_port(), M, K, N, panel_k, str(tmp_path)), nprocs=world, join=True)
    A = np.asfortranarray(np.load(tmp_path / "A0.npy"))
    for r in range(1, world):  # the broadcast delivered identical A everywhere
        assert np.array_equal(np.load(tmp_path / f"A{r}.npy"), A)
    X = np.asfortranarray(np.concatenate([np.load(tmp_path / f"X{r}.npy") for r in range(world)], axis=1))
    assert X.shape == (K, N)
    want = oracle.oracle_gemm(A, X)
    spans = [tuple(np.load(tmp_path / f"span{r}.npy")) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == N
    simt = np.asfortranarray(np.concatenate([np.load(tmp_path / f"D_simt{r}.npy") for r in range(world)], axis=1))
    assert simt.tobytes() == want.tobytes()  # N-GPU == single chain, bit for bit
    auto = np.asfortranarray(np.concatenate([np.load(tmp_path / f"D_auto{r}.npy") for r in range(world)], axis=1))
    ok, worst = oracle.error_bound_ok(auto, want, A, X)
    assert ok, worst
