#!/usr/bin/env python
"""Print the stage timeline of one host-pointer jmul! call (JBLAS_B200_TRACE=1): python tools/trace_host.py [n]"""
import os, sys, time
os.environ["JBLAS_B200_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import jblas.jl_b200 as jb
from jblas.jl_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
jb.init(0)
A = np.asfortranarray(jb.mrandn(n, n, seed=1).cpu().numpy()); X = np.asfortranarray(jb.mrandn(n, n, seed=2).cpu().numpy())
D = np.empty((n, n), order="F")
L = _lib.lib()
for a in (A, X, D):
    _lib.check(L.jblas_b200_host_register(a.ctypes.data, a.nbytes))
os.environ.pop("JBLAS_B200_TRACE")
jb.jmul_(D, A, X)
os.environ["JBLAS_B200_TRACE"] = "1"
t = time.perf_counter(); jb.jmul_(D, A, X); print("wall ms", 1e3 * (time.perf_counter() - t), file=sys.stderr)
