#!/usr/bin/env python
"""Static check of the FFMA2 kernels' SASS: how many FFMA2 read their 64-bit A pair and their 64-bit accumulator pair from the
same register bank pair (register index mod 4 equal).  ncu's per-instruction stall samples on B200 put the math-pipe stalls of
the exact FP32 kernel on exactly those instructions.    python tools/ffma2_bank_check.py [kernel-substring]"""
import re, subprocess, sys, collections
so = sys.argv[2] if len(sys.argv) > 2 else "jblas/jl_b200/libjblas_b200.so"
want = sys.argv[1] if len(sys.argv) > 1 else "SimtCfgIfLi2ELi4ELi32ELi3ELi2EEELb1ELb0"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, stats = None, collections.defaultdict(lambda: [0, 0, 0])
pat = re.compile(r"FFMA2\s+R(\d+),\s+R(\d+)\.F32x2\.HI_LO,\s+R(\d+)\.F32,\s+R(\d+)\.F32x2\.HI_LO")
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and "f32x2" in cur and want in cur:
        m = pat.search(line)
        if m:
            d, a, b, c = map(int, m.groups())
            st = stats[cur]
            st[0] += 1
            st[1] += (a % 4) == (c % 4)          # A pair and accumulator pair in the same bank pair
            st[2] += (b % 4) in (c % 4, c % 4 + 1) or (b % 4) in (a % 4, a % 4 + 1)  # scalar shares a bank with a pair
for k, (n, pc, sc) in stats.items():
    print(f"{k[:110]}\n   FFMA2 {n}  A/acc pair conflicts {pc} ({100*pc/max(n,1):.0f} %)  scalar-bank overlaps {sc} ({100*sc/max(n,1):.0f} %)")
