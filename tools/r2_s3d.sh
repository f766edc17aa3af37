#!/bin/bash
set -u
mkdir -p gpurun_out
bash tools/ncu_skinny.sh 3 300000
timeout 900 python -m pytest tests/test_tf32x3_gpu.py tests/test_gemm_gpu.py -m gpu -x -q -k "tf32" > gpurun_out/pytest_tf32.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_tf32.log
timeout 300 python bench.py --workload c3 --kernel tf32x3 --no-cpu-baseline > gpurun_out/bench_c3_tf32x3.json 2> gpurun_out/bench_c3_tf32x3.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_c3_tf32x3.json
