#!/bin/bash
# ncu --set full captures of the batched fastmul kernels and the mid-size GEMM; summaries are made ON the box (the
# .ncu-rep files together exceed the 64 MiB that travels back), one report is kept for source-level reading.
set -u
mkdir -p gpurun_out
run() { name=$1; regex=$2; note=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:${regex}" -s 1 -c 1 -o "gpurun_out/prof_${name}" -f python tools/ncu_target.py "$@" > "gpurun_out/ncu_${name}.log" 2>&1; echo "ncu ${name} rc=$?"
  python tools/ncu_summary.py "gpurun_out/prof_${name}.ncu-rep" "gpurun_out/sum_${name}.txt" "${note}" > /dev/null 2>&1; echo "summary ${name} rc=$?"
  [ "${name}" = "fb_dmma" ] || rm -f "gpurun_out/prof_${name}.ncu-rep"
}
run fb_dmma fastmul_batched_dmma "batched fastmul! 16x32x14 f64 x 1e6: one warp per product, DMMA fragments straight from HBM" batched 16 32 14 1000000 2
run fb_tma64 gemm_dmma_tma "batched 64x64x64 f64 x 60000: TMA/DMMA GEMM kernel over 3-D tensor maps (64x64x32_s3_x2, BATCHED)" batched 64 64 64 60000 2
run fb_tiny fastmul_batched_dmma "batched 8x8x8 f64 x 4e6: U = 4 products per warp iteration" batched 8 8 8 4000000 2
run dmma_2048 gemm_dmma_tma "FP64 2048^3 AUTO (dynamic tile scheduler)" float64 2048 2048 2048 auto 2
ls -la gpurun_out | tail
