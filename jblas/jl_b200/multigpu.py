"""Single-node multi-GPU mode of the jmul! path: one process per GPU (torch.distributed, NCCL over NVLink).

Partition.  D = A*X is sharded by COLUMN BLOCKS of D and X -- exactly the outer `cc` loop of the
reference's jmul! (src/gemm.jl:313): column block j of D depends only on column block j of X and on all of A.
Column blocks are contiguous in column-major storage, so shards are born and stay on their GPU; there is no
reduction and no gather.  The one real exchange step is making A visible everywhere: rank `root` owns A and
broadcasts it once per product.

Pipeline.  A is broadcast in K-PANELS (a column panel A[:, k0:k1] of a column-major matrix is one contiguous
M*(k1-k0) block) on the NCCL stream while the local GEMM consumes earlier panels:
    panel p arrives  ->  D_shard (+)= A[:, kp] * X_shard[kp, :]      accumulate = (p > 0)
The accumulate pass has the reference's kernel! semantics (src/kernels.jl:226): each element's chain stays
ascending in k across panels, so an N-GPU result is bit-identical to the 1-GPU result with the exact kernels.

Two transports for the exchange step (`bcast=`):
  "nccl" : torch.distributed.broadcast per panel.  NCCL's kernels need SMs; the persistent GEMM CTAs fill every SM, so the
           two compete (8 GPUs, 32768^3: ~5 % of the step).
  "p2p"  : the owner exports its A allocation over CUDA IPC once (jblas_b200_ipc_export), every other rank maps it and PULLS
           the K panels with the copy engines over NVLink (jblas_b200_copy_async on the communication stream) -- no kernel,
           no SM.  Two tiny all-reduces per call (20 us each, on the compute stream, outside the multiplies) fence the owner's
           buffer: one before the pulls (A is final), one at the end (A may be overwritten once the call returns).  The owner
           multiplies with one full-K launch.
"""
from __future__ import annotations

from typing import Callable, Optional


def column_shard(n_cols: int, world: int, rank: int) -> tuple[int, int]:
    """[c0, c1) of the columns of X and D owned by `rank`; remainder columns go to the last ranks."""
    if world < 1 or not (0 <= rank < world) or n_cols < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(n_cols, world)
    counts = [base + (1 if r >= world - rem else 0) for r in range(world)]
    c0 = sum(counts[:rank])
    return c0, c0 + counts[rank]


def k_panels(K: int, panel_k: int, first_k: int | None = None) -> list[tuple[int, int]]:
    """K split into [k0, k1) panels of `panel_k` (multiple of 64 keeps 16-byte staging and full k-tiles).

    `first_k` (multiple of 64, < panel_k) makes the FIRST panel shorter: the only broadcast that cannot hide behind a
    multiply is the first one, so it should be small."""
    if panel_k <= 0:
        raise ValueError("panel_k must be positive")
    if panel_k % 64 or (first_k is not None and (first_k <= 0 or first_k % 64)):
        raise ValueError("panel sizes must be positive multiples of 64")
    if K <= 0:
        return [(0, 0)]
    edges = [0]
    if first_k is not None and first_k < panel_k and first_k < K:
        edges.append(first_k)
    while edges[-1] < K:
        edges.append(min(edges[-1] + panel_k, K))
    return list(zip(edges[:-1], edges[1:]))


def _default_local_gemm(D, A, X, accumulate: bool, kernel):
    from . import api

    return api._gemm(D, A, X, accumulate, kernel)  # the CUDA path; raises if the library / GPU is missing


class ShardedGemm:
    """D_shard = A * X_shard on every rank, A broadcast from `root` in K panels overlapped with compute.

    All matrices are column-major torch tensors (strides (1, ld)).  `A` must be an M x K buffer on every rank;
    its contents matter on `root` only and are overwritten elsewhere.  `local_gemm(D, A, X, accumulate, kernel)`
    is injectable so the host logic (partition, panel schedule, ordering) can be exercised on CPU with the gloo
    backend; the default is the CUDA kernel path and there is no automatic fallback.
    """

    def __init__(self, M: int, K: int, n_cols_total: int, group=None, root: int = 0, panel_k: int = 2048,
                 kernel: Optional[int] = None, local_gemm: Optional[Callable] = None, first_panel_k: Optional[int] = None,
                 bcast: str = "nccl"):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.root = root
        self.M, self.K, self.n_total = M, K, n_cols_total
        self.c0, self.c1 = column_shard(n_cols_total, self.world, self.rank)
        self.panels = k_panels(K, panel_k, first_panel_k) if self.world > 1 else [(0, K)]
        self.kernel = kernel
        self.local_gemm = local_gemm or _default_local_gemm
        self._comm_stream = None
        if bcast not in ("nccl", "p2p", "auto"):
            raise ValueError("bcast must be 'nccl', 'p2p' or 'auto' (p2p when the CUDA IPC mapping succeeds on every rank)")
        self.bcast = bcast
        self._peer = {}    # local A buffer address -> address of the owner's A as mapped into this process
        self._fence = None

    def _owner_ptr(self, A):
        """Address of the owner's A in THIS process: CUDA IPC mapping, established once per A buffer (collective call).
        The owner's and the peers' A buffers must stay the same allocations for the life of this object."""
        key = A.data_ptr()
        if key in self._peer:
            return self._peer[key]
        import ctypes
        import struct

        import torch

        from . import _lib

        L = _lib.lib()
        wire = torch.zeros(73, dtype=torch.uint8, device=A.device)  # handle (64) | offset (8) | export succeeded (1)
        err = None
        if self.rank == self.root:
            handle, off = (ctypes.c_ubyte * 64)(), ctypes.c_int64()
            try:
                _lib.check(L.jblas_b200_ipc_export(A.data_ptr(), handle, ctypes.byref(off)))
                wire.copy_(torch.frombuffer(bytearray(bytes(handle) + struct.pack("<q", off.value) + b"\x01"), dtype=torch.uint8))
            except Exception as e:  # noqa: BLE001 -- reported to every rank below; the collective sequence must stay aligned
                err = e
        self.dist.broadcast(wire, src=self.root, group=self.group)
        ptr = A.data_ptr()
        if self.rank != self.root:
            raw = bytes(wire.cpu().numpy().tobytes())
            if raw[72] != 1:
                err = RuntimeError("the owner could not export its A allocation over CUDA IPC")
            else:
                handle = (ctypes.c_ubyte * 64).from_buffer_copy(raw[:64])
                mapped = ctypes.c_void_p()
                try:
                    _lib.check(L.jblas_b200_ipc_open(handle, struct.unpack("<q", raw[64:72])[0], ctypes.byref(mapped)))
                    ptr = mapped.value
                except Exception as e:  # noqa: BLE001
                    err = e
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=A.device)
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN, group=self.group)  # every rank learns whether ALL mappings exist
        if int(ok.item()) != 1:
            if err is None and self.rank != self.root and ptr != A.data_ptr():
                L.jblas_b200_ipc_close(ptr)
            raise RuntimeError(f"CUDA IPC mapping of the owner's A failed on at least one rank ({err})")
        self._peer[key] = ptr
        return ptr

    def close(self):
        """Unmap the owner's buffers (peers only)."""
        if self.rank != self.root and self._peer:
            from . import _lib

            for ptr in self._peer.values():
                _lib.lib().jblas_b200_ipc_close(ptr)
        self._peer = {}

    def _call_p2p(self, D_shard, A, X_shard):
        import torch

        from . import _lib

        L = _lib.lib()
        dev = A.device
        owner = self._owner_ptr(A)
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=dev)
        if self._fence is None:
            self._fence = torch.zeros(1, dtype=torch.int32, device=dev)
        compute, comm = torch.cuda.current_stream(dev), self._comm_stream
        es = A.element_size()
        # Both fences run on the COMPUTE stream, i.e. strictly between GEMM launches: a fence kernel that waited for a slow peer
        # while this rank's persistent GEMM CTAs fill every SM would either starve or hold an SM the next launch needs
        # (measured: +3 ms per step when the second fence overlapped the multiplies).
        self.dist.all_reduce(self._fence, group=self.group)  # fence 1: the owner's A is final, every receive buffer is free
        ready = []
        if self.rank != self.root:
            comm.wait_stream(compute)
            for k0, k1 in self.panels:  # a K panel of a dense column-major matrix is one contiguous block
                _lib.check(L.jblas_b200_copy_async(A.data_ptr() + k0 * self.M * es, owner + k0 * self.M * es,
                                                   (k1 - k0) * self.M * es, comm.cuda_stream))
                ev = torch.cuda.Event()
                ev.record(comm)
                ready.append(ev)
            for p, (k0, k1) in enumerate(self.panels):
                compute.wait_event(ready[p])
                self.local_gemm(D_shard, self._panel(A, k0, k1), X_shard[k0:k1, :], p > 0, self.kernel)
        else:
            self.local_gemm(D_shard, A, X_shard, False, self.kernel)  # A is local: one full-K launch
        self.dist.all_reduce(self._fence, group=self.group)  # fence 2: every rank holds its copy; the owner may rewrite A
        return D_shard

    @property
    def shard_cols(self) -> int:
        return self.c1 - self.c0

    def _panel(self, A, k0, k1):
        return A[:, k0:k1]

    def _panel_wire(self, A, k0, k1):
        # A is dense column-major M x K: columns k0..k1 are ONE contiguous block; its transpose is the row-major
        # contiguous (k1-k0, M) tensor torch.distributed wants on the wire (same bytes, no copy)
        return A[:, k0:k1].t()

    def __call__(self, D_shard, A, X_shard):
        import torch

        M, K = self.M, self.K
        if tuple(A.shape) != (M, K) or tuple(X_shard.shape) != (K, self.shard_cols) or tuple(D_shard.shape) != (M, self.shard_cols):
            raise ValueError(f"shape mismatch: A {tuple(A.shape)}, X shard {tuple(X_shard.shape)}, D shard {tuple(D_shard.shape)}; "
                             f"expected ({M},{K}), ({K},{self.shard_cols}), ({M},{self.shard_cols})")
        if self.world == 1:
            self.local_gemm(D_shard, A, X_shard, False, self.kernel)
            return D_shard
        if not A.t().is_contiguous():
            raise ValueError("A must be dense column-major (leading dimension == M) to be broadcast in K panels")
        cuda = A.is_cuda
        if cuda and self.bcast == "auto":  # collective decision, taken once: p2p if every rank can map the owner's A
            try:
                self._owner_ptr(A)
                self.bcast = "p2p"
            except RuntimeError:
                self.bcast = "nccl"
        if cuda and self.bcast == "p2p":
            return self._call_p2p(D_shard, A, X_shard)
        works = []
        if cuda:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=A.device)
            compute = torch.cuda.current_stream(A.device)
            self._comm_stream.wait_stream(compute)  # A (on root) / the receive buffer (elsewhere) must be ready
            with torch.cuda.stream(self._comm_stream):
                for k0, k1 in self.panels:
                    works.append(self.dist.broadcast(self._panel_wire(A, k0, k1), src=self.root, group=self.group, async_op=True))
        else:
            for k0, k1 in self.panels:
                works.append(self.dist.broadcast(self._panel_wire(A, k0, k1), src=self.root, group=self.group, async_op=True))
        for p, (k0, k1) in enumerate(self.panels):
            if cuda:
                with torch.cuda.stream(self._comm_stream):
                    works[p].wait()  # orders the NCCL work before what follows on the comm stream
                    ev = torch.cuda.Event()
                    ev.record(self._comm_stream)
                torch.cuda.current_stream(A.device).wait_event(ev)
            else:
                works[p].wait()
            self.local_gemm(D_shard, self._panel(A, k0, k1), X_shard[k0:k1, :], p > 0, self.kernel)
        return D_shard

    def launches_per_call(self) -> int:
        return len(self.panels)

    def from_host(self, D_host, A_host, X_host, D_shard, A, X_shard, nblocks: int = 4):
        """End-to-end product on PINNED HOST shards (the host-facing form of the sharded mode).

        `A_host` (K x M row-major = A column-major) matters on `root` only, `X_host` (cols x K) and `D_host`
        (cols x M) are this rank's shards; `D_shard`, `A`, `X_shard` are the device buffers.  Three streams:
          in   : root uploads A in K panels (each panel is broadcast as soon as it has landed), every rank uploads
                 its X shard in `nblocks` column blocks (contiguous in column-major storage);
          comm : the K-panel NCCL broadcasts of A;
          main : once A is complete, column block b of D is multiplied (full K, one launch) as soon as X block b is
                 there; `out` copies D block b back to the host while block b+1 is multiplied.
        Every element's chain is a single-launch chain, so results equal __call__'s."""
        import torch

        if self.world == 1:
            raise ValueError("from_host is the multi-rank path; use the C ABI host-pointer entry on one GPU")
        dev = A.device
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=dev)
        if getattr(self, "_in_stream", None) is None:
            self._in_stream, self._out_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        main, comm, s_in, s_out = torch.cuda.current_stream(dev), self._comm_stream, self._in_stream, self._out_stream
        for st in (comm, s_in, s_out):
            st.wait_stream(main)  # previous users of the device buffers are done
        At, Xt, Dt = A.t(), X_shard.t(), D_shard.t()  # row-major views: (K, M), (cols, K), (cols, M)
        works = []
        for k0, k1 in self.panels:
            if self.rank == self.root:
                with torch.cuda.stream(s_in):
                    At[k0:k1].copy_(A_host[k0:k1], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(s_in)
                comm.wait_event(ev)
            with torch.cuda.stream(comm):
                works.append(self.dist.broadcast(At[k0:k1], src=self.root, group=self.group, async_op=True))
        cols = self.shard_cols
        nb = max(1, -(-cols // max(1, nblocks)))
        blocks = [(c0, min(c0 + nb, cols)) for c0 in range(0, cols, nb)]
        x_ready = []
        with torch.cuda.stream(s_in):
            for c0, c1 in blocks:
                Xt[c0:c1].copy_(X_host[c0:c1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_in)
                x_ready.append(ev)
        with torch.cuda.stream(comm):
            for w in works:
                w.wait()
            a_ready = torch.cuda.Event()
            a_ready.record(comm)
        main.wait_event(a_ready)
        for (c0, c1), ev in zip(blocks, x_ready):
            main.wait_event(ev)
            self.local_gemm(D_shard[:, c0:c1], A, X_shard[:, c0:c1], False, self.kernel)
            done = torch.cuda.Event()
            done.record(main)
            s_out.wait_event(done)
            with torch.cuda.stream(s_out):
                D_host[c0:c1].copy_(Dt[c0:c1], non_blocking=True)
        main.wait_stream(s_out)
        return D_host
