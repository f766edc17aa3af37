"""CPU tests (-m "not gpu"): the C-ABI library loads, exports every symbol include/jblas_b200.h declares,
its host-side planner behaves, and every compute entry fails LOUDLY without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "jblas_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jblas_b200_[a-z0-9_]+)\s*\(", src)))


def _has_cuda():
    import torch

    return torch.cuda.is_available()


def test_header_declares_the_hot_path():
    syms = _declared_symbols()
    for must in ["jblas_b200_gemm_f64", "jblas_b200_gemm_f32", "jblas_b200_jmul_f64", "jblas_b200_kernel_f64",
                 "jblas_b200_initkernel_f64", "jblas_b200_fastmul_f64", "jblas_b200_gemm_f64_dev", "jblas_b200_randn_fill"]:
        assert must in syms


def test_library_exports_every_declared_symbol(jb):
    from jblas.jl_b200 import _lib

    L = ctypes.CDLL(_lib.SO_PATH)
    missing = [s for s in _declared_symbols() if not hasattr(L, s)]
    assert not missing, missing
    # and the ctypes prototype table covers exactly the header
    assert sorted(_lib.PROTOTYPES) == _declared_symbols()
    assert _lib.lib().jblas_b200_version() == 100


def test_header_cites_reference_lines():
    src = open(HEADER).read()
    for cite in ["src/gemm.jl:244", "src/kernels.jl:239", "src/kernels.jl:273", "src/kernels.jl:202", "src/randmat.jl:5-14",
                 "test/runtests.jl:97-101"]:
        assert cite in src


def test_planner_is_pure_host_logic(jb):
    p = jb.plan(8192, 8192, 8192)
    # AUTO on an aligned big square: the persistent TMA warp-specialised DMMA kernel, one CTA per SM
    assert p["kernel"].startswith("dmma_tma_f64") and p["tile_m"] == 128 and p["tile_n"] == 128 and p["grid"] == 148
    s = jb.plan(8192, 8192, 8192, kernel=jb.F64_SIMT)
    assert s["kernel"].startswith("simt_f64") and s["staging"] == "cp.async 16B"
    assert s["grid"] == (8192 // s["tile_m"]) * (8192 // s["tile_n"])  # one CTA per tile (non-persistent)
    assert jb.plan(8192, 8192, 8192, kernel=jb.F64_DMMA)["kernel"].startswith("dmma")
    # ragged leading dimensions (M = 1023 doubles per column) cannot use TMA boxes: mid-size products run the persistent
    # kernels with their element-wise (ragged) producers, very large ones are re-aligned into scratch first and then take the
    # 128 x 128 TMA kernel, small ones use the plain element-wise cp.async kernels
    r = jb.plan(1023, 4097, 777)
    assert "ragged producer" in r["staging"] and r["kernel"].startswith("dmma_tma_f64") and 56 <= r["grid"] <= 296
    huge = jb.plan(8191, 8191, 8191)
    assert "re-aligning" in huge["staging"] and huge["kernel"] == "dmma_tma_f64_128x128x32_s3"
    assert "re-aligning" in jb.plan(1021, 1027, 515, "float32", lda=1023, ldx=1033)["staging"]
    small = jb.plan(129, 17, 127)
    assert small["staging"].startswith("cp.async element-wise") and small["kernel"].startswith("dmma")
    assert jb.plan(16384, 16384, 16384, "float32")["kernel"].startswith("simt_f32")
    names = jb.kernel_names()
    for i, n in enumerate(names):
        dt = "float64" if "f64" in n else "float32"
        assert jb.plan(512, 512, 512, dt, kernel=jb.EXPLICIT_BASE + i)["kernel"] == n
    with pytest.raises(jb.JblasB200Error):
        jb.plan(0, 1, 1)
    with pytest.raises(jb.JblasB200Error):
        jb.plan(512, 512, 512, "float32", kernel=jb.EXPLICIT_BASE + 0)  # a Float64 kernel for Float32 data


def test_argument_checks_mirror_reference_dispatch_errors(jb):
    D = np.zeros((4, 5), order="F")
    A = np.zeros((4, 3), order="F")
    X = np.zeros((3, 5), order="F")
    with pytest.raises(ValueError):
        jb.jmul_(D, A, np.zeros((4, 5), order="F"))  # inner dimension mismatch (MethodError in Julia)
    with pytest.raises(TypeError):
        jb.jmul_(D, A.astype(np.float32), X)  # mixed element types
    with pytest.raises(ValueError):
        jb.jmul_(D, np.zeros((4, 3), order="C"), X)  # row-major storage
    with pytest.raises(TypeError):
        jb.jmul_(D.astype(np.int64), A.astype(np.int64), X.astype(np.int64))
    with pytest.raises(ValueError):
        jb.kernel_(np.zeros(10), np.zeros(100), np.zeros(100), jb.Kernel(8, 4, 8, 5, 5))  # pD too small


def test_no_cpu_fallback_without_gpu(jb):
    if _has_cuda():
        pytest.skip("a GPU is present")
    D = np.full((4, 5), np.nan, order="F")
    A = np.ones((4, 3), order="F")
    X = np.ones((3, 5), order="F")
    with pytest.raises(jb.JblasB200Error) as e:
        jb.jmul_(D, A, X)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    assert np.isnan(D).all()  # nothing was computed behind our back
    from jblas.jl_b200 import _lib

    L = _lib.lib()
    assert L.jblas_b200_gemm_f64_dev(1, 1, 1, 4, 3, 5, 4, 4, 3, 0, 0, None) == -3  # ENOTINIT
    assert b"no CPU fallback" in L.jblas_b200_last_error()
    assert L.jblas_b200_jmul_f64(D.ctypes.data, A.ctypes.data, X.ctypes.data, 4, 3, 5) == -3
    # the newer entries fail the same way: fused forms, batched fastmul! (host and device), peer memory
    C = np.ones((4, 5), order="F")
    for fn in (jb.gemm_plus_c_, jb.gemm_x_plus_c_):
        with pytest.raises(jb.JblasB200Error):
            fn(D, A, X, C if fn is jb.gemm_plus_c_ else np.ones((3, 5), order="F"))
    bA, bX = np.ones((2, 3, 4)).transpose(0, 2, 1), np.ones((2, 5, 3)).transpose(0, 2, 1)
    bD = np.full((2, 5, 4), np.nan).transpose(0, 2, 1)
    with pytest.raises(jb.JblasB200Error):
        jb.fastmul_batched_(bD, bA, bX)
    assert np.isnan(D).all() and np.isnan(bD).all()
    assert L.jblas_b200_fastmul_batched_f64_dev(1, 1, 1, 4, 3, 5, 2, 20, 12, 15, None) == -3
    assert L.jblas_b200_copy_async(1, 1, 8, None) == -3
    import ctypes

    off = ctypes.c_int64()
    assert L.jblas_b200_ipc_export(1, (ctypes.c_ubyte * 64)(), ctypes.byref(off)) == -3


def test_multigpu_transport_argument_is_validated():
    from jblas.jl_b200.multigpu import ShardedGemm

    with pytest.raises(ValueError):
        ShardedGemm(8, 8, 8, bcast="rdma")
    assert ShardedGemm(8, 8, 8, bcast="auto").bcast == "auto"  # decided collectively at the first GPU call


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "jblas")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


@pytest.mark.parametrize("fname,dtype,limit", [("r1_size_sweep_per_kernel.json", "float64", 0.05), ("r2_size_sweep_f32_per_kernel.json", "float32", 0.12)])
def test_planner_pick_is_close_to_the_measured_best(jb, fname, dtype, limit):
    """Regression guard for the planner's cost model (capi.cu: make_plan): on every shape of the committed per-kernel B200
    sweeps (tools/size_sweep.py --all) the kernel it picks must be within `limit` of the fastest registered kernel.
    (Without a GPU the occupancy of the one-CTA-per-tile kernels falls back to their launch bounds; the picks may then
    differ from the on-device ones, the bound still has to hold.)"""
    import json

    path = os.path.join(ROOT, "profiles", fname)
    sweep = json.load(open(path))
    checked = 0
    for key, row in sweep.items():
        per = row.get("per_kernel_ms")
        if not per or not key.startswith(dtype):
            continue
        M, N, K = (int(v) for v in key.split("_")[1].split("x"))
        pick = jb.plan(M, K, N, dtype)["kernel"]
        assert pick in per, (key, pick)
        best = min(per.values())
        assert per[pick] <= (1.0 + limit) * best, (key, pick, per[pick], best)
        checked += 1
    assert checked >= 10


def test_planner_tall_skinny_rules(jb):
    """Pure host logic: which tall-skinny Float64 shapes leave the tile kernels (DESIGN.md s4 "Tall-skinny", profiles/r2_skinny_compare.txt)."""
    team, xres, xreg = "dmma_skinny_f64_16x16_xreg_team_w16", "dmma_skinny_f64_16x64_xres_w12", "dmma_skinny_f64_16x16_xreg_w8"
    assert jb.plan(65536, 64, 64)["kernel"] == xreg and jb.plan(40000, 64, 48)["kernel"] == xreg  # A fits in L2: private boxes
    assert jb.plan(131072, 32, 48)["kernel"] == team
    assert jb.plan(131072, 64, 64)["kernel"] == xreg and jb.plan(300000, 64, 64)["kernel"] == team and jb.plan(1 << 20, 64, 64)["kernel"] == team
    assert jb.plan(65536, 64, 32)["kernel"] == xres and jb.plan(262144, 32, 9)["kernel"] == team and jb.plan(65536, 72, 48)["kernel"] == xres
    for shape in [(4096, 64, 64), (65536, 256, 64), (65536, 64, 128), (65535, 64, 64), (65536, 36, 64)]:
        assert "skinny" not in jb.plan(*shape)["kernel"], shape
    assert "skinny" not in jb.plan(65536, 64, 64, kernel=jb.F64_SIMT)["kernel"]


def test_documented_kernel_names_exist(jb):
    """Every full kernel name quoted in DESIGN.md / README.md / INTEGRATION.md is (a prefix of) a registered kernel: the documents
    must not drift from the registry the planner and the explicit selectors use."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set(jb.kernel_names())
    missing = []
    for fn in ("DESIGN.md", "README.md", "INTEGRATION.md"):
        with open(os.path.join(root, fn)) as f:
            for n in re.findall(r"`((?:dmma|simt|tf32x3)_[a-z0-9_]+)`", f.read()):
                if n not in names and not any(k.startswith(n) for k in names):
                    missing.append((fn, n))
    assert not missing, sorted(set(missing))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` needs no GPU: ONE JSON line on stdout with the contract keys of the reference arm -- the CPU
    restatement of the jmul! loop nest timed on a bounded slab of the same workload (the only other place bench.py may execute oracle/)."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=240, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "TFLOP/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"].startswith("GEMM TFLOP/s") and "8192" in d["config"]["workload"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
