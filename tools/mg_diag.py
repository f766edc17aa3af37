#!/usr/bin/env python
"""Diagnostics of the p2p transport (run under torchrun, >= 2 GPUs): fence latency, copy-engine pull bandwidth, and the
timeline of one step on a non-owner rank."""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
import jblas.jl_b200 as jb
from jblas.jl_b200 import _lib, api
from jblas.jl_b200.multigpu import ShardedGemm
jb.init(lr)
L = _lib.lib()
M = K = N = 8192
sg = ShardedGemm(M, K, N * world, panel_k=2048, first_panel_k=256, bcast="p2p")
A = jb.mrandn(M, K, seed=1) if rank == 0 else jb.empty_colmajor(M, K)
X = jb.mrandn(K, sg.shard_cols, seed=2, first_col=sg.c0)
D = jb.empty_colmajor(M, sg.shard_cols)
owner = sg._owner_ptr(A)
flag = torch.zeros(1, dtype=torch.int32, device=dev)
def ev(): return torch.cuda.Event(enable_timing=True)
# (a) fence latency
for _ in range(5): dist.all_reduce(flag)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = ev(), ev(); e0.record()
for _ in range(20): dist.all_reduce(flag)
e1.record(); torch.cuda.synchronize()
out = [f"rank {rank}: fence (tiny all_reduce) {e0.elapsed_time(e1)/20*1e3:.1f} us"]
# (b) pull bandwidth
if rank != 0:
    for mib in (16, 128, 512):
        nbytes = mib << 20
        s = torch.cuda.current_stream().cuda_stream
        _lib.check(L.jblas_b200_copy_async(A.data_ptr(), owner, nbytes, s)); torch.cuda.synchronize()
        e0, e1 = ev(), ev(); e0.record()
        for _ in range(5): _lib.check(L.jblas_b200_copy_async(A.data_ptr(), owner, nbytes, s))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out.append(f"rank {rank}: pull {mib} MiB in {ms:.3f} ms = {nbytes/ms/1e6:.0f} GB/s")
dist.barrier()
# (c) steps
for mode in ("p2p", "nccl"):
    sg2 = ShardedGemm(M, K, N * world, panel_k=2048, first_panel_k=256, bcast=mode)
    if mode == "p2p": sg2._peer = sg._peer
    for _ in range(5): sg2(D, A, X)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = ev(), ev(); e0.record()
    for _ in range(10): sg2(D, A, X)
    e1.record(); torch.cuda.synchronize()
    out.append(f"rank {rank}: {mode} {e0.elapsed_time(e1)/10:.3f} ms/step")
    dist.barrier()
# local GEMM alone: one launch vs 5 accumulate panels
for _ in range(2): api._gemm(D, A, X, False, None)
torch.cuda.synchronize()
e0, e1 = ev(), ev(); e0.record()
for _ in range(5): api._gemm(D, A, X, False, None)
e1.record(); torch.cuda.synchronize()
out.append(f"rank {rank}: single launch {e0.elapsed_time(e1)/5:.3f} ms")
e0, e1 = ev(), ev(); e0.record()
for _ in range(5):
    for p, (k0, k1) in enumerate(sg.panels): api._gemm(D, A[:, k0:k1], X[k0:k1, :], p > 0, None)
e1.record(); torch.cuda.synchronize()
out.append(f"rank {rank}: 5 panel launches {e0.elapsed_time(e1)/5:.3f} ms")
for r in range(world):
    dist.barrier()
    if r == rank: print("\n".join(out), flush=True)
dist.destroy_process_group()
