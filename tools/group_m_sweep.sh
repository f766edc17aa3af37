#!/bin/bash
# Raster group height (JBLAS_B200_GROUP_M) vs DRAM traffic and time of the headline kernel (FP64 8192^3 AUTO).
# usage (on a GPU box): bash tools/group_m_sweep.sh > gpurun_out/group_m_sweep.txt
set -u
mkdir -p gpurun_out
echo "# FP64 8192^3 AUTO (dmma_tma_f64_128x128x32_s3): raster group height vs DRAM bytes per launch (ncu) and CUDA-event time (no profiler)"
echo "# group_m  dram_read_GB  dram_write_GB  traffic/algorithmic(1.61 GB)  ms_per_launch(events, 6 launches)"
for h in 0 2; do
export JBLAS_B200_L2HINT=$h
echo "# JBLAS_B200_L2HINT=$h (0: no eviction hints; 2: A evict_last, X normal)"
for g in 4 8 12 16 24 64; do
  JBLAS_B200_GROUP_M=$g timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:gemm_dmma_tma" -s 1 -c 1 --csv \
     --log-file gpurun_out/gm_$g.csv python tools/ncu_target.py float64 8192 8192 8192 auto 2 > /dev/null 2>&1
  ms=$(JBLAS_B200_GROUP_M=$g python - <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
import jblas.jl_b200 as jb
from jblas.jl_b200 import api
jb.init(0)
A = jb.mrandn(8192, 8192, "float64", seed=1); X = jb.mrandn(8192, 8192, "float64", seed=2); D = jb.empty_colmajor(8192, 8192, "float64")
for _ in range(3): api._gemm(D, A, X, False, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(6): api._gemm(D, A, X, False, None)
e1.record(); torch.cuda.synchronize()
print("%.3f" % (e0.elapsed_time(e1) / 6))
PY
)
  python - "$g" "$ms" <<'PY'
import csv, sys
g, ms = sys.argv[1], sys.argv[2]
rd = wr = 0.0
for row in csv.reader(open(f"gpurun_out/gm_{g}.csv")):
    if len(row) > 3 and "dram__bytes_read.sum" in row: 
        v = float(row[-1].replace(",", "")); u = row[-2]
        rd = v * {"Gbyte": 1, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}.get(u, 1)
    if len(row) > 3 and "dram__bytes_write.sum" in row:
        v = float(row[-1].replace(",", "")); u = row[-2]
        wr = v * {"Gbyte": 1, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}.get(u, 1)
print(f"{g:>8s}  {rd:12.3f}  {wr:13.3f}  {(rd + wr) / 1.6106:28.2f}  {ms}")
PY
done
done
