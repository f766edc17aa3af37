#!/usr/bin/env python
"""Ablation of the exact-FP32 inner loop: FFMA2 stream with / without its shared-memory operand loads (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jblas.jl_b200 import build as _build
_build.build(probes=True)  # the ablation probes are not in the shipped library
import jblas.jl_b200 as jb
jb.init(0)
print("ffma2_tile (registers only)", jb.probe_pipe("ffma2_tile", 10000)[0])
names = {0: "LDS.128 A + LDS X", 1: "+ syncthreads", 2: "A via LDS.64", 3: "A via LDS.64 + sync", 4: "no X loads", 8: "no A loads", 12: "no loads at all",
         5: "no X loads + sync", 9: "no A loads + sync", 13: "no loads + sync", 6: "A via LDS.64, no X loads"}
names.update({16: "MAP1: LDS.128 A + LDS X", 18: "MAP1: A via LDS.64", 20: "MAP1: no X loads", 24: "MAP1: no A loads", 22: "MAP1: A LDS.64, no X"})
names.update({32: "MAP2: LDS.128 A + LDS X", 33: "MAP2: + syncthreads", 36: "MAP2: no X loads", 40: "MAP2: no A loads"})
only = [int(a) for a in sys.argv[1:]]
for m in (only or (12, 13, 8, 4, 0, 1, 2, 3, 6, 16, 18, 20, 24, 22)):
    print(f"mode {m:2d} {names[m]:28s}", round(jb.probe_pipe(f"ffma2_lds{m}", 2000)[0], 2), flush=True)
for m, nm in () if only else ((2, "8x16 tile, no loads"), (0, "8x16 tile, LDS"), (1, "8x16 tile, LDS + sync")):
    print(f"wide {m} {nm:28s}", round(jb.probe_pipe(f"ffma2_wide{m}", 4000)[0], 2), flush=True)
