"""Build recipe for libjblas_b200.so (hand-written CUDA for sm_100a + the C ABI of include/jblas_b200.h).

    python -m jblas.jl_b200.build [--force] [--verbose] [--probes] [--fast]

The library is built IN-TREE (jblas/jl_b200/libjblas_b200.so) with an explicit nvcc command so the
prebuilt .so travels with the repository snapshot to the GPU box.  `-gencode arch=compute_100a,code=sm_100a`
(not plain -arch=sm_100a, which also emits a compute_100 PTX pass that rejects tcgen05) and `-lineinfo`
so ncu's source page maps SASS back to these files.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libjblas_b200.so")

SOURCES = ["capi.cu"]
LAST_BUILD_MODE = "not built in this process"  # "rebuilt" | "reused" (the in-tree .so was newer than every source)
NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libjblas_b200.so cannot be built (there is no CPU fallback)")


def build_info() -> dict:
    """How the in-tree library came to be (bench.py records it): written next to the .so by the build that made it."""
    try:
        import json

        with open(SO + ".buildinfo.json") as f:
            info = json.load(f)
    except Exception:
        info = {"nvcc": None, "flags": None, "host": None, "built_at": None}
    info["so_mtime"] = os.path.getmtime(SO) if os.path.exists(SO) else None
    info["so_bytes"] = os.path.getsize(SO) if os.path.exists(SO) else None
    return info


def _stale() -> bool:
    if not os.path.exists(SO):
        return True
    flags = build_info().get("flags") or ""
    if "TUNING_PROBES" in flags or "split-compile" in flags:  # a development build must not be taken for the shipped one
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "jblas_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, probes: bool = False, fast: bool = False) -> str:
    """probes: also compile the FFMA2 shared-memory ablation probes (tools/probe_ffma2.py); fast: nvcc --split-compile=0
    (development iterations only: code generation differs slightly, never used for a measured build)."""
    global LAST_BUILD_MODE
    if not (force or probes or _stale()):
        LAST_BUILD_MODE = "reused"
        return SO
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", SO, *[os.path.join(CSRC, s) for s in SOURCES], "-lcuda"]
    if probes:
        cmd.insert(1, "-DJBLAS_B200_TUNING_PROBES")
    if fast or os.environ.get("JBLAS_B200_FAST_BUILD"):
        cmd.insert(1, "--split-compile=0")
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), flush=True)
    # the image exports CC/CXX pointing at a wrapper gcc; nvcc wants the distro host compiler
    env = dict(os.environ)
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building libjblas_b200.so")
    LAST_BUILD_MODE = "rebuilt"
    try:
        import json
        import socket
        import time

        ver = subprocess.run([_nvcc(), "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
        with open(SO + ".buildinfo.json", "w") as f:
            json.dump({"nvcc": ver, "flags": " ".join(c for c in cmd[1:] if not c.startswith("/")), "host": socket.gethostname(),
                       "built_at": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}, f)
    except Exception:
        pass
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, probes="--probes" in sys.argv, fast="--fast" in sys.argv))
