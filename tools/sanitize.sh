#!/bin/bash
# compute-sanitizer memcheck over small parity tests of the newer kernels (dynamic tile scheduler, batched DMMA / 3-D TMA, fused
# forms, re-align pass).  Slow under the sanitizer: small shapes only.
set -u
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest -x -q \
  "tests/test_fastmul_batched_gpu.py::test_fastmul_batched_strided_batch_and_special_values" \
  "tests/test_fastmul_batched_gpu.py::test_fastmul_batched_bit_identical" \
  "tests/test_fused_gpu.py" \
  "tests/test_gemm_gpu.py::test_dmma_tma_kernels_even_strides_accumulate_edges" \
  "tests/test_gemm_gpu.py::test_exact_kernels_strided_leading_dimensions" \
  -k "not 64x64x64 and not 97x11x33 and not 300x1000x260" > gpurun_out/sanitize.log 2>&1
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|passed|failed" gpurun_out/sanitize.log | head -20
tail -3 gpurun_out/sanitize.log
