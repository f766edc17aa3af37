#!/usr/bin/env python
"""bench.py -- headline benchmark of the jmul! path on B200 (contract: see the task brief, section 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|c4a|c4b|c5|fb|fb32] [--kernel auto|dmma|simt]

One "step" = one product D = A*X over one batch of synthetic N(0,1) input (mrandn, src/randmat.jl:5-14).
  N = 1  : BASELINE.json configs[1]  -- Float64 M=N=K=8192 (the config the metric is quoted on).
  N > 1  : the column-sharded mode: every rank owns an 8192-column block of X and D, A (8192x8192)
           lives on rank 0 and is broadcast in K panels overlapped with the local GEMM -> weak scaling,
           N = 1 being exactly configs[1].  `--workload c5` runs BASELINE configs[4] instead (32768^3 strong).
`--workload fb` (N = 1 only) is the SURVEY 8f-1 row: 10^6 independent 16x32x14 Float64 fastmul! products in one launch, the
one shape the reference publishes a time for (test/runtests.jl:110-122); HBM-bound, so its roofline object is in GB/s.
Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference's own loop nest
(oracle/jmul_baseline.c; Julia is not available) on rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED_A = 0x6A424C41
SEED_X = SEED_A + 1
METRIC = "GEMM TFLOP/s (FP64, FP32) and % of roofline at 1/2/4/8 B200 vs CPU jBLAS"
FP64_NOMINAL_TFLOPS = 37.2  # 148 SM x 64 FMA/clk x 2 x 1.965 GHz
FP32_NOMINAL_TFLOPS = 74.4
TF32X3_NOMINAL_TFLOPS = 1125.0 / 3  # dense TF32 tensor peak (half the 2.25 PFLOP/s bf16 figure) / 3 MMAs per product

WORKLOADS = {
    # name: (dtype, M, N, K, description)
    "c2": ("float64", 8192, 8192, 8192, "Float64 jmul! M=N=K=8192 (BASELINE configs[1])"),
    "c3": ("float32", 16384, 16384, 16384, "Float32 jmul! M=N=K=16384 exact SIMT (BASELINE configs[2])"),
    "c4a": ("float64", 1023, 777, 4097, "Float64 jmul! ragged M=1023 N=777 K=4097 (BASELINE configs[3])"),
    "c4b": ("float64", 65536, 64, 64, "Float64 jmul! tall-skinny M=65536 N=64 K=64 (BASELINE configs[3])"),
    "c5": ("float64", 32768, 32768, 32768, "Float64 jmul! M=N=K=32768 column-sharded (BASELINE configs[4])"),
}


# --------------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region
# --------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0: float, t1: float) -> dict:
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------
# CPU legs (the ONLY place bench.py touches oracle/: as the reported baseline, never as the product)
# --------------------------------------------------------------------------------------------------------
class CpuSlab:
    """The reference's jmul! loop nest (restatement, oracle/jmul_baseline.c) on a bounded slab of column tiles."""

    def __init__(self, dtype: str, M: int, N: int, K: int, target_s: float, nthreads: int = 1):
        import numpy as np

        import oracle

        self.oracle, self.np = oracle, np
        self.dtype, self.M, self.N, self.K, self.nthreads = dtype, M, N, K, nthreads
        dt = np.float64 if dtype == "float64" else np.float32
        self.vl, self.rows, self.cols = oracle.jmul_baseline_tile(np.dtype(dt).itemsize)
        rng = np.random.Generator(np.random.PCG64(SEED_A))
        self.max_tiles = max(1, N // self.cols)
        self.A = np.asfortranarray(rng.standard_normal((K, M)).T.astype(dt))
        probe_tiles = min(self.max_tiles, 4)
        Xp = np.asfortranarray(rng.standard_normal((probe_tiles * self.cols, K)).T.astype(dt))
        Dp = np.empty((M, probe_tiles * self.cols), dtype=dt, order="F")
        t = time.perf_counter()
        oracle.jmul_baseline(Dp, self.A, Xp, nthreads=nthreads)
        per_tile = (time.perf_counter() - t) / probe_tiles
        self.tiles = int(min(self.max_tiles, max(probe_tiles, target_s / max(per_tile, 1e-9))))
        self.X = np.asfortranarray(rng.standard_normal((self.tiles * self.cols, K)).T.astype(dt))
        self.D = np.empty((M, self.tiles * self.cols), dtype=dt, order="F")

    def run(self):
        t = time.perf_counter()
        r, c = self.oracle.jmul_baseline(self.D, self.A, self.X, nthreads=self.nthreads)
        sec = time.perf_counter() - t
        flops = 2.0 * r * c * self.K  # only what the reference's tile loops cover (it skips remainder rows/cols)
        return flops, sec, (r, c)

    def describe(self, r, c, sec):
        return (f"restatement of the jmul! loop nest (Julia unavailable): tile {self.rows}x{self.cols} from pick_kernel_size, {self.tiles} of "
                f"{self.max_tiles} column tiles of the {self.M}x{self.N}x{self.K} {self.dtype} workload = {r}x{c} outputs in {sec:.1f} s, "
                f"{self.nthreads} thread (the reference is single-threaded); host has {os.cpu_count()} cores")


def cpu_baseline_sample(dtype: str, M: int, N: int, K: int, target_s: float = 12.0, nthreads: int = 1):
    slab = CpuSlab(dtype, M, N, K, target_s, nthreads)
    flops, sec, (r, c) = slab.run()
    return {"value": flops / sec / 1e12, "unit": "TFLOP/s", "cores": nthreads, "kind": "port", "sample": slab.describe(r, c, sec)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    dtype, M, N, K, desc = WORKLOADS[args.workload or "c2"]
    slab = CpuSlab(dtype, M, N, K, target_s=4.0)
    for _ in range(args.warmup):
        slab.run()
    tot_flops = tot_s = 0.0
    r = c = 0
    for _ in range(args.steps):
        f, s_, (r, c) = slab.run()
        tot_flops += f
        tot_s += s_
    value = tot_flops / tot_s / 1e12
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_s / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if dtype == "float64" else "f32", "data": "synthetic N(0,1) (mrandn), host PCG64",
        "config": {"workload": desc, "note": "each step = a bounded slab of column tiles of this workload on 1 host core (the reference is single-threaded)"},
        "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": 1, "kind": "port", "sample": slab.describe(r, c, tot_s / max(args.steps, 1))},
        "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


FB_SHAPE = (16, 32, 14, 1_000_000)  # jBLAS names M, N, P (D MxP = A MxN * X NxP) and the batch
FB_PUBLISHED_NS = 127.921            # minimum time of ONE such product on the author's CPU, 1 thread


def run_batched(args):
    """--workload fb: batched fastmul! on one GPU.  Same JSON contract; the dominant kernel is HBM-bound."""
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch

    if args.gpus != 1 or int(os.environ.get("WORLD_SIZE", "1")) != 1:
        raise SystemExit("--workload fb is a single-GPU workload (independent products: N GPUs = N replicas)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    from jblas.jl_b200 import build

    build.build()
    import jblas.jl_b200 as jb

    jb.init(0)
    M, N, P, batch = FB_SHAPE
    f32 = args.workload == "fb32"  # same shape in Float32 (the reference is generic in T), twice the products
    dtn, es = ("float32", 4) if f32 else ("float64", 8)
    if f32:
        batch *= 2
    A = jb.mrandn_batch(batch, M, N, dtn, seed=SEED_A)
    X = jb.mrandn_batch(batch, N, P, dtn, seed=SEED_X)
    D = jb.empty_colmajor_batch(batch, M, P, dtn, fill=float("nan"))
    flops = 2.0 * M * N * P * batch
    algo_bytes = (M * N + N * P + M * P) * es * batch  # every matrix read or written exactly once
    for _ in range(max(args.warmup, 3)):
        jb.fastmul_batched_(D, A, X)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.25)
    launches0 = jb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        jb.fastmul_batched_(D, A, X)
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1) / args.steps
    launches = jb.launch_count() - launches0
    clocks = sampler.summary(t0, t1)
    sampler.stop()
    assert not torch.isnan(D).any().item()
    peaks, peak_src = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = algo_bytes / (ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(f"{'fastmul_batched_f32_warp' if f32 else 'fastmul_batched_dmma'}@{M}x{N}x{P}x{batch}")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                "kernel": "fastmul_batched_f32_warp_kernel<8,14>" if f32 else "fastmul_batched_dmma_kernel<2,2,8,1,true>", "kernel_ms": ms, "algorithmic_bytes_per_launch": algo_bytes,
                "flops_per_launch": flops, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src}; a copy bandwidth, 1 read : 1 write -- this "
                "kernel reads 2.4 bytes per byte written, so a fraction slightly above 1 is possible)"}
    # e2e: the C-ABI host-pointer entry on pinned host batches (chunked H2D / kernel / D2H pipeline inside the call)
    from jblas.jl_b200 import _lib

    L = _lib.lib()
    Ah = np.ascontiguousarray(A.transpose(1, 2).cpu().numpy()).transpose(0, 2, 1)  # (batch, M, N), column-major matrices
    Xh = np.ascontiguousarray(X.transpose(1, 2).cpu().numpy()).transpose(0, 2, 1)
    Dh = np.full((batch, P, M), np.nan, dtype=Ah.dtype).transpose(0, 2, 1)
    bases = [a.base if a.base is not None else a for a in (Ah, Xh, Dh)]
    for a in bases:
        _lib.check(L.jblas_b200_host_register(a.ctypes.data, a.nbytes))
    steps = max(1, min(args.steps, 5))
    try:
        jb.fastmul_batched_(Dh, Ah, Xh)
        t = time.perf_counter()
        for _ in range(steps):
            jb.fastmul_batched_(Dh, Ah, Xh)
        sec = (time.perf_counter() - t) / steps
    finally:
        for a in bases:
            L.jblas_b200_host_unregister(a.ctypes.data)
    assert not np.isnan(Dh).any()
    e2e = {"value": flops / sec / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": (M * N + N * P) * es * batch, "d2h_bytes_per_step": M * P * es * batch,
           "steps": steps, "ms_per_step": 1e3 * sec, "pcie_gbs": algo_bytes / sec / 1e9,
           "api": f"jblas_b200_fastmul_batched_{'f32' if f32 else 'f64'} (host pointers, pinned by jblas_b200_host_register; PCIe-bound: 1.5 flop per byte moved)"}
    cpu = None
    if not args.no_cpu_baseline and not f32:  # the fastmul! restatement is Float64 (the published number is)
        import oracle

        nb = 200_000  # bounded sample: 200k of the 10^6 products (1.9 GB of the workload), repeated until ~10 s have passed
        rng = np.random.Generator(np.random.PCG64(SEED_A))
        Ac = rng.standard_normal((nb, N, M)).transpose(0, 2, 1)
        Xc = rng.standard_normal((nb, P, N)).transpose(0, 2, 1)
        Dc = np.empty((nb, P, M)).transpose(0, 2, 1)
        oracle.fastmul_baseline_batched(Dc[:1000], Ac[:1000], Xc[:1000])
        reps, sec_cpu = 0, 0.0
        while sec_cpu < 10.0 and reps < 200:
            t = time.perf_counter()
            oracle.fastmul_baseline_batched(Dc, Ac, Xc)
            sec_cpu += time.perf_counter() - t
            reps += 1
        ns_each = sec_cpu / (reps * nb) * 1e9
        cpu = {"value": 2.0 * M * N * P / (ns_each * 1e-9) / 1e12, "unit": "TFLOP/s", "cores": 1, "kind": "port",
               "sample": f"restatement of fastmul! (oracle/jmul_baseline.c, register-resident 2x14 zmm accumulators) over {nb} of the {batch} "
                         f"products x {reps} passes, streaming from DRAM: {ns_each:.0f} ns per product on 1 thread (the reference publishes "
                         f"{FB_PUBLISHED_NS} ns for one cache-resident product on the author's CPU); host has {os.cpu_count()} cores"}
    value = flops / (ms * 1e-3) / 1e12
    published_tflops = 2.0 * M * N * P / (FB_PUBLISHED_NS * 1e-9) / 1e12
    line = {"metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None if f32 else value / published_tflops, "dtype": "f32" if f32 else "f64",
            "data": "synthetic N(0,1), device-generated Philox (mrandn analogue)",
            "config": {"workload": f"fastmul! batched: {batch} independent {'Float32' if f32 else 'Float64'} products D({M}x{P}) = A({M}x{N}) * X({N}x{P}) per launch (SURVEY 8f-1)",
                       "products_per_s": batch / (ms * 1e-3), "ns_per_product": ms * 1e6 / batch,
                       "vs_baseline_note": f"BASELINE.md publishes {FB_PUBLISHED_NS} ns per product (1 CPU thread, author's machine) = {published_tflops:.4f} TFLOP/s",
                       "l2": f"inputs larger than L2: {algo_bytes / 2**20:.0f} MiB per step vs 126 MB L2"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    return 0


def parity_sample(D, A, X, n_rows=48, n_cols=24, seed=5, extra_rel=0.0):
    """Checker, outside every timed region: a sampled sub-grid of this rank's D shard against the CPU oracle chain on the
    identical input bits (each element's chain is independent -- jmul!'s tile loops, src/gemm.jl:313 -- so sampling is
    legitimate).  The shard's first/last rows and columns (the seams between ranks) are always in the sample."""
    import numpy as np
    import torch

    import oracle

    M, ns = D.shape
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = np.unique(np.concatenate([[0, 1, M - 2, M - 1, 127, 128], rng.integers(0, M, n_rows)]).clip(0, M - 1))
    cols = np.unique(np.concatenate([[0, 1, ns - 2, ns - 1], rng.integers(0, ns, n_cols)]).clip(0, ns - 1))
    tr, tc = torch.from_numpy(rows).to(D.device), torch.from_numpy(cols).to(D.device)
    As = np.asfortranarray(A[tr, :].cpu().numpy())
    Xs = np.asfortranarray(X[:, tc].cpu().numpy())
    got = np.asfortranarray(D[tr][:, tc].cpu().numpy())
    want = oracle.oracle_gemm(As, Xs)
    ok, worst = oracle.error_bound_ok(got, want, As, Xs, extra_rel=extra_rel)
    return {"rows": int(rows.size), "cols": int(cols.size), "bit_identical": bool(got.tobytes() == want.tobytes()),
            "within_bound": bool(ok), "worst_err_over_bound": float(worst), "nan_free": bool(not torch.isnan(D).any().item())}


def merge_parity(dist, dev, world, mine):
    """All ranks' samples -> one object (every rank must pass)."""
    import torch

    flags = torch.tensor([int(mine["bit_identical"]), int(mine["within_bound"]), int(mine["nan_free"]), -mine["worst_err_over_bound"]],
                         device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    return {"rows": mine["rows"], "cols": mine["cols"], "ranks": world, "bit_identical": bool(flags[0].item() == 1),
            "within_bound": bool(flags[1].item() == 1), "nan_free": bool(flags[2].item() == 1), "worst_err_over_bound": float(-flags[3].item()),
            "checker": "oracle/oracle_gemm.c chain on the same input bits; per rank: rows x cols of its D shard incl. the shard seams"}


def time_other_config(jb, name, kernel=None, extra_rel=0.0):
    """One of the other BASELINE configs on one GPU: device-resident, CUDA events, buffers rotated so that consecutive
    launches never find their operands in L2 (>= 300 MB of distinct A/X/D), sampled parity against the oracle."""
    import torch

    from jblas.jl_b200 import api

    dtype, M, N, K, desc = WORKLOADS[name]
    es = 8 if dtype == "float64" else 4
    bytes_ = (M * K + K * N + M * N) * es
    sets = max(1, min(8, -(-300_000_000 // bytes_))) if bytes_ < 300_000_000 else 1
    bufs = [(jb.mrandn(M, K, dtype, seed=SEED_A + 2 * i), jb.mrandn(K, N, dtype, seed=SEED_X + 2 * i), jb.empty_colmajor(M, N, dtype, fill=float("nan")))
            for i in range(sets)]
    flops = 2.0 * M * N * K
    reps = int(max(2, min(200, 0.4 / max(flops / 30e12, 1e-5))))
    for A, X, D in bufs:
        api._gemm(D, A, X, False, kernel)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        A, X, D = bufs[i % sets]
        api._gemm(D, A, X, False, kernel)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    A, X, D = bufs[0]
    par = parity_sample(D, A, X, extra_rel=extra_rel)
    nominal = FP64_NOMINAL_TFLOPS if dtype == "float64" else FP32_NOMINAL_TFLOPS
    if kernel is not None and dtype == "float32" and kernel == jb.F32_3XTF32:
        nominal = TF32X3_NOMINAL_TFLOPS  # three TF32 MMAs per credited FMA
    out = {"config": name, "workload": desc, "kernel": jb.plan(M, K, N, dtype, kernel=kernel)["kernel"], "ms": ms, "tflops": flops / (ms * 1e-3) / 1e12,
           "frac_of_nominal": flops / (ms * 1e-3) / 1e12 / nominal, "nominal_peak": nominal, "algorithmic_gbs": bytes_ / (ms * 1e-3) / 1e9,
           "reps": reps, "l2": f"{sets} rotating operand sets ({sets * bytes_ / 1e6:.0f} MB)" if sets > 1 else f"operands {bytes_ / 1e6:.0f} MB > L2",
           "parity_check": par}
    del bufs, A, X, D
    torch.cuda.empty_cache()
    return out


def run_gpu(args):
    # stdout must carry exactly ONE JSON line.  Libraries (NCCL's version banner, torchrun notices) printf() to fd 1,
    # so fd 1 is pointed at stderr for the whole run and the JSON line is written to the saved real stdout at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_group = None
    if world != args.gpus:
        if world == 1 and args.gpus > 1:  # convenience: relaunch ourselves under torchrun (children get the real stdout)
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                   "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
            return subprocess.call(cmd, stdout=real_stdout)
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly ONE JSON line: NCCL printf()s a "NCCL version ..." banner to stdout at the VERSION and
        # WARN debug levels, so those two are dropped (INFO and above go through NCCL's logger and are left alone)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")  # host-side waits (an NCCL barrier would spin a kernel on every GPU)

    from jblas.jl_b200 import build

    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    import jblas.jl_b200 as jb
    from jblas.jl_b200 import _lib
    from jblas.jl_b200.multigpu import ShardedGemm

    jb.init(local_rank)
    wl = args.workload or "c2"
    dtype, M, N, K, desc = WORKLOADS[wl]
    if world > 1 and wl == "c2":
        n_total, scaling = N * world, "weak"     # every rank owns an 8192-column block
        desc = f"Float64 jmul! column-sharded: M=K=8192, N=8192 per GPU x {world} GPUs, A broadcast from rank 0 in K panels"
    else:
        n_total, scaling = N, ("strong" if world > 1 else "weak")
    if dtype == "float32":
        selector = {"auto": None, "simt": jb.F32_EXACT, "tf32x3": jb.F32_3XTF32}.get(args.kernel)
        if args.kernel not in ("auto", "simt", "tf32x3"):
            raise SystemExit("--kernel for a Float32 workload must be auto, simt or tf32x3")
    else:
        if args.kernel == "tf32x3":
            raise SystemExit("--kernel tf32x3 needs a Float32 workload (--workload c3)")
        selector = {"auto": None, "dmma": jb.F64_DMMA, "simt": jb.F64_SIMT}[args.kernel]
    sg = ShardedGemm(M, K, n_total, panel_k=args.panel_k, kernel=selector, first_panel_k=args.first_panel_k or None, bcast=args.bcast)
    A = jb.mrandn(M, K, dtype, seed=SEED_A) if rank == 0 else jb.empty_colmajor(M, K, dtype)
    X = jb.mrandn(K, sg.shard_cols, dtype, seed=SEED_X, first_col=sg.c0)
    D = jb.empty_colmajor(M, sg.shard_cols, dtype, fill=float("nan"))
    es = 8 if dtype == "float64" else 4
    flops_step = 2.0 * M * K * n_total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # untimed pre-warm: ~1 s of steps so clocks and power reach their steady state before the W warm-up steps (a cold box
    # measured 32.2 ms/step, the same command a minute later 31.0)
    # (a fixed COUNT derived from the workload, identical on every rank: each step contains collectives)
    n_pre = int(min(40, 1.0 / max(flops_step / world / 36e12, 1e-3)))
    for _ in range(n_pre):
        sg(D, A, X)
    torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        sg(D, A, X)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    launches0 = jb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        sg(D, A, X)
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = jb.launch_count() - launches0
    total_ms = float(ms.item())
    value = flops_step * args.steps / (total_ms * 1e-3) / 1e12
    clocks = sampler.summary(t0, t1) if rank == 0 else None
    if rank == 0:
        sampler.stop()
    assert not torch.isnan(D).any().item(), "NaN sentinel survived: some element of D was not written"
    # ---- parity of what was just timed (checker, outside the timed region): every rank samples its own D shard ----
    parity = merge_parity(dist, dev, world, parity_sample(D, A, X, extra_rel=(2.0 ** -18 if args.kernel == "tf32x3" else 0.0)))
    assert parity["within_bound"] and parity["nan_free"], f"parity check failed: {parity}"

    # ---- roofline of the dominant kernel (the GEMM kernel itself: at N=1 the step IS one launch) ----
    roofline = None
    if rank == 0:
        peaks, peak_src = load_peaks()
        if dtype == "float64":
            # register-resident probes: the DMMA warp-tile pattern (what the tensor pipe can sustain with free operands)
            # and DFMA in the 8x8 outer-product pattern (what a SIMT micro-kernel can sustain)
            dmma, _ = jb.probe_pipe("dmma_tile", 4000)
            dfma, _ = jb.probe_pipe("dfma_tile", 10000)
            peak, nominal = max(dfma, dmma), FP64_NOMINAL_TFLOPS
            probe = {"dmma_tile_tflops": dmma, "dfma_tile_tflops": dfma}
        else:
            ffma, _ = jb.probe_pipe("ffma", 40000)
            ffma2, _ = jb.probe_pipe("ffma2_tile", 10000)
            peak, nominal = max(ffma, ffma2), FP32_NOMINAL_TFLOPS
            probe = {"ffma_chain_tflops": ffma, "ffma2_tile_tflops": ffma2}
            if args.kernel == "tf32x3":
                # 3 TF32 MMAs per credited FMA: the bound is the dense TF32 tensor peak / 3; MEASURED_PEAKS.json has a
                # measured bf16 figure, TF32 runs at half the bf16 rate on this part (nominal 1.1 vs 2.25 PFLOP/s)
                # the timed region is a long run of back-to-back launches: the SUSTAINED bf16 figure applies (this pool's B200s
                # are power-capped under tensor load: the 3xTF32 kernel runs at 1.26-1.5 GHz with sw_power_cap set)
                bf16 = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
                peak, nominal = bf16 / 2.0 / 3.0, TF32X3_NOMINAL_TFLOPS
                probe = {"bf16_tflops_sustained_" + peak_src: bf16, "bf16_tflops_burst": peaks.get("bf16_tflops"), "tf32_over_3": peak}
        # kernel-only duration: time the dominant local kernel launch (the largest K panel) alone on this stream
        k0, k1 = max(sg.panels, key=lambda p: p[1] - p[0])
        per_launch_flops = 2.0 * M * (k1 - k0) * sg.shard_cols
        ke0, ke1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, min(args.steps, 10))
        from jblas.jl_b200 import api
        api._gemm(D, A[:, k0:k1], X[k0:k1, :], False, selector)
        torch.cuda.synchronize()
        ke0.record()
        for _ in range(reps):
            api._gemm(D, A[:, k0:k1], X[k0:k1, :], False, selector)
        ke1.record()
        torch.cuda.synchronize()
        kms = ke0.elapsed_time(ke1) / reps
        achieved = per_launch_flops / (kms * 1e-3) / 1e12
        algo_bytes = (M * (k1 - k0) + (k1 - k0) * sg.shard_cols + M * sg.shard_cols) * es
        pl = jb.plan(M, k1 - k0, sg.shard_cols, dtype, kernel=selector)
        traffic = None
        try:  # DRAM bytes per launch of this kernel on this shape, from the committed ncu capture (if one exists)
            tmap = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            traffic = tmap.get(f"{pl['kernel']}@{M}x{sg.shard_cols}x{k1 - k0}")
        except Exception:
            traffic = None
        roofline = {
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
            "kernel": pl["kernel"], "kernel_ms": kms, "flops_per_launch": per_launch_flops, "algorithmic_bytes_per_launch": algo_bytes,
            "peak_source": ("live register-only pipe probe in this run (MEASURED_PEAKS.json carries only HBM and bf16 figures): " + json.dumps(probe)),
            "frac_of_nominal": achieved / nominal, "nominal_peak": nominal,
            "hbm_gbs_if_bytes_bound": algo_bytes / (kms * 1e-3) / 1e9, "hbm_peak_gbs": peaks.get("hbm_gbs"), "hbm_peak_source": peak_src,
        }

    # ---- e2e: the reference-facing call with HOST buffers (H2D of A and X, D2H of D inside the timed region) ----
    e2e = None if args.no_e2e else run_e2e(args, jb, _lib, sg, A, X, D, dtype, M, K, world, rank, dev, flops_step, selector, cpu_group)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample(dtype, M, N, K, target_s=12.0)

    # ---- the other BASELINE configs, so that every config's speed AND parity is in the driver's own run ----
    other = None
    strong_c5 = None
    if wl == "c2" and not args.no_other_configs:
        del D, X, A
        torch.cuda.empty_cache()
        if world == 1:
            other = [time_other_config(jb, "c4a"), time_other_config(jb, "c4b"), time_other_config(jb, "c3"),
                     time_other_config(jb, "c3", kernel=jb.F32_3XTF32, extra_rel=2.0 ** -18)]
            other[-1]["config"] = "c3_3xtf32"
        strong_c5 = run_strong_c5(args, jb, dist, dev, world, rank, selector, cpu_group)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64" if dtype == "float64" else "f32", "data": "synthetic N(0,1), device-generated Philox (mrandn analogue), seeds jBLA/jBLA+1",
            "config": {"workload": desc, "M": M, "K": K, "N_total": n_total, "N_per_gpu": sg.shard_cols, "kernel_selector": args.kernel,
                       "parallelism": f"column-shard x{world}" + (f", A broadcast ({sg.bcast}) in {len(sg.panels)} K-panels (first {sg.panels[0][1] - sg.panels[0][0]}, then {args.panel_k})" if world > 1 else ""),
                       "l2": (f"inputs larger than L2: A+X+D = {(M * K + K * sg.shard_cols + M * sg.shard_cols) * es / 2**20:.0f} MiB per GPU vs 126 MB L2"
                              if (M * K + K * sg.shard_cols + M * sg.shard_cols) * es > 2 * 126e6 else
                              f"L2-WARM and host-call bound: A+X+D = {(M * K + K * sg.shard_cols + M * sg.shard_cols) * es / 2**20:.0f} MiB per GPU fit the 126 MB L2 and one step is one "
                              "Python call; the cold-operand figure of this config (rotating sets > L2, 200 back-to-back launches) is the other_configs entry of the default line")},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "parity_check": parity, "other_configs": other, "strong_c5": strong_c5,
            "build": dict(build.build_info(), mode=build.LAST_BUILD_MODE,
                          note="mode = what build.build() did in THIS process: 'reused' = the in-tree .so that travelled with the snapshot was newer than every source"),
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


C5_ONE_GPU_FILE = os.path.join(ROOT, "profiles", "c5_1gpu.json")


def run_strong_c5(args, jb, dist, dev, world, rank, selector, cpu_group=None):
    """BASELINE configs[4]: Float64 32768^3, column-sharded over `world` GPUs (strong scaling), A (8 GiB) resident on rank 0
    and delivered in K panels behind the multiplies.  1 warm-up + 3 timed steps, max over ranks, sampled parity per rank.
    `efficiency` = T1 / (world * T_world) with T1 the one-GPU time of the same code: measured in this run when world == 1
    (and written to gpurun_out/), otherwise the committed profiles/c5_1gpu.json."""
    import torch

    from jblas.jl_b200.multigpu import ShardedGemm

    n = 32768
    sg = ShardedGemm(n, n, n, panel_k=args.panel_k, kernel=selector, first_panel_k=args.first_panel_k or None, bcast=args.bcast)
    A = jb.mrandn(n, n, "float64", seed=SEED_A) if rank == 0 else jb.empty_colmajor(n, n, "float64")
    X = jb.mrandn(n, sg.shard_cols, "float64", seed=SEED_X, first_col=sg.c0)
    D = jb.empty_colmajor(n, sg.shard_cols, "float64", fill=float("nan"))
    sg(D, A, X)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    steps = 3 if world > 1 else 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sg(D, A, X)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    parity = merge_parity(dist, dev, world, parity_sample(D, A, X, n_rows=24, n_cols=12))
    bcast = sg.bcast
    sg.close()
    del A, X, D, sg
    torch.cuda.empty_cache()
    flops = 2.0 * n * n * n
    # the same config END TO END: host A, X, D (3 x 8 GiB, pinned) through the C-ABI host-pointer entry on `world` GPUs of one
    # process (rank 0 makes the call; at world == 1 it is the plain one-GPU entry).  2 * 8 GiB up and 8 GiB back per call.
    e2e = None
    if not args.no_e2e:
        import numpy as np

        if world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=cpu_group)
        if rank == 0:
            shard = n // world
            seams = [c for g in range(1, world) for c in (g * shard - 1, g * shard)]
            sec, h2d, d2h, par = mgpu_host_call(jb, np, "float64", n, n, n, world, selector, 1 if world == 1 else 2, SEED_A, SEED_X, seams)
            e2e = {"value": flops / sec / 1e12, "unit": "TFLOP/s", "ms_per_step": 1e3 * sec, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "parity_check": par, "api": "jblas_b200_mgpu_gemm_f64, one host process, pinned host matrices" if world > 1 else "jblas_b200_gemm_f64, pinned host matrices"}
        if world > 1:
            dist.barrier(group=cpu_group)
    out = {"workload": WORKLOADS["c5"][4], "n_gpus": world, "steps": steps, "ms_per_step": ms, "tflops": flops / (ms * 1e-3) / 1e12, "e2e": e2e,
           "frac_of_nominal_per_gpu": flops / (ms * 1e-3) / 1e12 / world / FP64_NOMINAL_TFLOPS, "scaling": "strong", "transport": bcast,
           "parity_check": parity}
    if world == 1:
        out["efficiency"] = 1.0
        out["t1_ms"], out["t1_source"] = ms, "this run"
    else:
        try:
            ref = json.load(open(C5_ONE_GPU_FILE))
            out["t1_ms"], out["t1_source"] = float(ref["ms_per_step"]), "profiles/c5_1gpu.json: " + ref.get("source", "")
            out["efficiency"] = out["t1_ms"] / (world * ms)
            if e2e is not None and ref.get("e2e_ms_per_step"):
                out["e2e"]["efficiency"] = float(ref["e2e_ms_per_step"]) / (world * e2e["ms_per_step"])
        except Exception:
            out["efficiency"] = None
    return out


def host_link_probe(n_gpus, mib=256, reps=3):
    """What THIS box's host side can move when `n_gpus` GPUs copy at once from/to pinned host memory (one process, one
    stream per GPU and direction): GiB/s summed over the GPUs, each direction alone and both together.  It bounds every
    end-to-end figure that moves A, X and D through host memory -- on the pool's 8-GPU VMs the host side, not the GPUs'
    own links, is the limit (profiles/r2_pcie_probe_8gpu.txt)."""
    import torch

    hin = [torch.empty(mib << 20, dtype=torch.uint8).pin_memory() for _ in range(n_gpus)]
    hout = [torch.empty(mib << 20, dtype=torch.uint8).pin_memory() for _ in range(n_gpus)]
    din = [torch.empty(mib << 20, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n_gpus)]
    dout = [torch.empty(mib << 20, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n_gpus)]
    sin = [torch.cuda.Stream(device=d) for d in range(n_gpus)]
    sout = [torch.cuda.Stream(device=d) for d in range(n_gpus)]

    def run(h2d, d2h):
        def once():
            for d in range(n_gpus):
                if h2d:
                    with torch.cuda.stream(sin[d]):
                        din[d].copy_(hin[d], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(sout[d]):
                        hout[d].copy_(dout[d], non_blocking=True)
            for d in range(n_gpus):
                torch.cuda.synchronize(d)

        once()
        t = time.perf_counter()
        for _ in range(reps):
            once()
        return mib / 1024 * n_gpus * reps / (time.perf_counter() - t)

    out = {"gpus": n_gpus, "mib_per_gpu_per_direction": mib, "h2d_only_gibs": run(True, False), "d2h_only_gibs": run(False, True),
           "each_direction_when_both_gibs": run(True, True)}
    del hin, hout, din, dout
    torch.cuda.empty_cache()
    return out


def link_floor_ms(probe, h2d_bytes, d2h_bytes):
    """Lower bound of one call's transfer time on this box: the better of 'uploads, then downloads' and 'both at once'."""
    gi = 2.0 ** 30
    seq = h2d_bytes / gi / probe["h2d_only_gibs"] + d2h_bytes / gi / probe["d2h_only_gibs"]
    both = max(h2d_bytes, d2h_bytes) / gi / probe["each_direction_when_both_gibs"]
    return 1e3 * min(seq, both)


def mgpu_host_call(jb, np, dtype, M, K, n_total, world, selector, steps, gen_seed_a, gen_seed_x, sample_seams, extra_rel=0.0):
    """Rank 0 only: the single-process multi-GPU entry of the C ABI on pinned host matrices (what jmul!(D, A, X; gpus = N) hits);
    returns (seconds per call, h2d bytes, d2h bytes, parity object)."""
    import torch

    import oracle

    ndt = np.float64 if dtype == "float64" else np.float32
    chunk = 4096
    Ah = np.empty((M, K), dtype=ndt, order="F")
    for c0 in range(0, K, chunk):
        Ah[:, c0:min(c0 + chunk, K)] = jb.mrandn(M, min(chunk, K - c0), dtype, seed=gen_seed_a, first_col=c0).cpu().numpy()
    Xh = np.empty((K, n_total), dtype=ndt, order="F")
    for c0 in range(0, n_total, chunk):  # the one global X (the ranks' shards are column blocks of it), generated on this GPU
        Xh[:, c0:min(c0 + chunk, n_total)] = jb.mrandn(K, min(chunk, n_total - c0), dtype, seed=gen_seed_x, first_col=c0).cpu().numpy()
    Dh = np.full((M, n_total), np.nan, dtype=ndt, order="F")
    torch.cuda.empty_cache()
    with jb.pinned(Ah, Xh, Dh):
        jb.jmul_(Dh, Ah, Xh, kernel=selector, gpus=world)  # warm-up: per-GPU contexts, workspaces, peer access
        t = time.perf_counter()
        for _ in range(steps):
            jb.jmul_(Dh, Ah, Xh, kernel=selector, gpus=world)
        sec = (time.perf_counter() - t) / steps
    assert not np.isnan(Dh).any(), "NaN sentinel survived in the multi-GPU host result"
    # parity of the end-to-end result (checker, after the timed region): sampled rows x columns of the host D, the columns on
    # both sides of every GPU's shard seam included, against the oracle chain on the same host bits
    rng = np.random.Generator(np.random.PCG64(7))
    rows = np.unique(np.concatenate([[0, 1, M - 2, M - 1], rng.integers(0, M, 20)]))
    cols = np.unique(np.concatenate([[0, n_total - 1], sample_seams, rng.integers(0, n_total, 10)]).astype(np.int64))
    As, Xs = np.asfortranarray(Ah[rows, :]), np.asfortranarray(Xh[:, cols])
    want = oracle.oracle_gemm(As, Xs)
    got = np.asfortranarray(Dh[np.ix_(rows, cols)])
    ok, worst = oracle.error_bound_ok(got, want, As, Xs, extra_rel=extra_rel)
    assert ok, f"multi-GPU end-to-end result outside the tolerance: {worst}"
    parity = {"rows": int(rows.size), "cols": int(cols.size), "shard_seams_sampled": len(sample_seams), "bit_identical": bool(got.tobytes() == want.tobytes()),
              "within_bound": bool(ok), "worst_err_over_bound": float(worst)}
    h2d, d2h = int(Ah.nbytes + Xh.nbytes), int(Dh.nbytes)
    del Ah, Xh, Dh
    return sec, h2d, d2h, parity


def run_e2e(args, jb, _lib, sg, A, X, D, dtype, M, K, world, rank, dev, flops_step, selector, cpu_group=None):
    """Same metric through the public host-facing API: every step copies that step's inputs host->device from
    pinned host memory and reads the result back."""
    import numpy as np
    import torch
    import torch.distributed as dist

    es = 8 if dtype == "float64" else 4
    steps = max(1, min(args.steps, 5))
    L = _lib.lib()
    h2d = d2h = 0
    if world == 1:
        # the C-ABI host-pointer entry (what a Julia ccall hits); caller buffers pinned once with jblas_b200_host_register
        Ah = np.asfortranarray(A.cpu().numpy())
        Xh = np.asfortranarray(X.cpu().numpy())
        Dh = np.full((M, sg.shard_cols), np.nan, dtype=Ah.dtype, order="F")
        for a in (Ah, Xh, Dh):
            _lib.check(L.jblas_b200_host_register(a.ctypes.data, a.nbytes))
        try:
            jb.jmul_(Dh, Ah, Xh, kernel=selector)  # warm-up (workspace allocation)
            t = time.perf_counter()
            for _ in range(steps):
                jb.jmul_(Dh, Ah, Xh, kernel=selector)
            sec = time.perf_counter() - t
        finally:
            for a in (Ah, Xh, Dh):
                L.jblas_b200_host_unregister(a.ctypes.data)
        assert not np.isnan(Dh).any()
        h2d, d2h = Ah.nbytes + Xh.nbytes, Dh.nbytes
        return {"value": flops_step * steps / sec / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": steps, "ms_per_step": 1e3 * sec / steps, "api": f"jblas_b200_gemm_{'f64' if dtype == 'float64' else 'f32'} (host pointers, pinned by jblas_b200_host_register)"}
    # N > 1: the SINGLE-PROCESS multi-GPU entry of the C ABI, jblas_b200_mgpu_gemm_* -- what a Julia `jmul!(D, A, X; gpus = N)`
    # hits: one host process, the whole A / X / D in pinned host memory, N GPUs driven from it.  Rank 0 makes the call on
    # all N GPUs; the other ranks of this torchrun job hold no GPU work meanwhile and wait on a HOST-side (gloo) barrier.
    n_total = sg.n_total
    result = [None]
    torch.cuda.synchronize()
    dist.barrier(group=cpu_group)
    if rank == 0:
        seams = [c for g in range(1, world) for c in (g * sg.shard_cols - 1, g * sg.shard_cols)]
        sec, h2d, d2h, parity = mgpu_host_call(jb, np, dtype, M, K, n_total, world, selector, steps, SEED_A, SEED_X, seams,
                                               extra_rel=(2.0 ** -18 if args.kernel == "tf32x3" else 0.0))
        probe = host_link_probe(world)
        floor = link_floor_ms(probe, h2d, d2h)
        result[0] = {"value": flops_step / sec / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": steps,
                     "ms_per_step": 1e3 * sec, "parity_check": parity,
                     "host_link_probe": probe, "host_link_floor_ms": floor, "frac_of_host_link_floor": floor / (1e3 * sec),
                     "note": "host_link_floor_ms = the time this box's host side needs just to move the step's H2D and D2H bytes with all "
                             f"{world} GPUs copying at once (measured in this run); when it exceeds the compute time the end-to-end figure is bound by the host, not by the GPUs",
                     "api": f"jblas_b200_mgpu_gemm_{'f64' if dtype == 'float64' else 'f32'} (ONE host process, host pointers pinned by jblas_b200_host_register, "
                            f"{world} GPUs: A slices over every GPU's own PCIe link + NVLink peer pulls, X / D column blocks per GPU; wall clock around the call)"}
    dist.barrier(group=cpu_group)
    return result[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS) + ["fb", "fb32"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "dmma", "simt", "tf32x3"])
    ap.add_argument("--panel-k", type=int, default=2048)
    ap.add_argument("--first-panel-k", type=int, default=256, help="shorter first K panel of the A broadcast (0 = same as the others)")
    ap.add_argument("--bcast", default="auto", choices=["auto", "nccl", "p2p"],
                    help="N > 1: how A reaches the other GPUs (multigpu.py); auto = copy-engine pulls over CUDA IPC if every rank can map A, else NCCL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (very large workloads)")
    ap.add_argument("--no-other-configs", action="store_true", help="default workload only: skip the other BASELINE configs (c3, c4a, c4b at N=1; 32768^3 strong at every N)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.workload in ("fb", "fb32"):
            raise SystemExit("--impl reference times the jmul! loop nest on the GEMM workloads; the fb leg reports its CPU figure in cpu_baseline")
        return run_reference(args)
    if args.workload in ("fb", "fb32"):
        return run_batched(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
