#!/usr/bin/env python
"""Host-link ceilings of this box: pinned H2D / D2H bandwidth with 1, 2, 4, 8 GPUs copying at the same time (one host process,
one stream per GPU and direction).  These numbers bound every end-to-end figure that moves A, X and D through host memory:
    python tools/pcie_probe.py > gpurun_out/pcie_probe.txt"""
import time

import torch

n_dev = torch.cuda.device_count()
MB = 256
host_in = [torch.empty(MB << 20, dtype=torch.uint8).pin_memory() for _ in range(n_dev)]
host_out = [torch.empty(MB << 20, dtype=torch.uint8).pin_memory() for _ in range(n_dev)]
dev_in = [torch.empty(MB << 20, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n_dev)]
dev_out = [torch.empty(MB << 20, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n_dev)]
s_in = [torch.cuda.Stream(device=d) for d in range(n_dev)]
s_out = [torch.cuda.Stream(device=d) for d in range(n_dev)]


def run(n, h2d, d2h, reps=4):
    def once():
        for d in range(n):
            if h2d:
                with torch.cuda.stream(s_in[d]):
                    dev_in[d].copy_(host_in[d], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s_out[d]):
                    host_out[d].copy_(dev_out[d], non_blocking=True)
        for d in range(n):
            torch.cuda.synchronize(d)

    once()
    t = time.perf_counter()
    for _ in range(reps):
        once()
    sec = (time.perf_counter() - t) / reps
    return MB / 1024 * n / sec  # GiB/s per direction, all GPUs together


print(f"{n_dev} GPUs visible; {MB} MiB per GPU per direction, pinned host memory (torch cudaHostAlloc), GiB/s summed over the GPUs")
print(f"{'GPUs':>4} {'H2D only':>10} {'D2H only':>10} {'H2D (both)':>11} {'D2H (both)':>11}")
for n in (1, 2, 4, 8):
    if n > n_dev:
        break
    a, b, c = run(n, True, False), run(n, False, True), run(n, True, True)
    print(f"{n:>4} {a:>10.1f} {b:>10.1f} {c:>11.1f} {c:>11.1f}")
