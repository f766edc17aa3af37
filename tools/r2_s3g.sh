#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k "skinny" > gpurun_out/pytest_skinny.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_skinny.log
timeout 600 python tools/skinny_compare.py > gpurun_out/skinny_compare.txt 2>&1; cat gpurun_out/skinny_compare.txt | tail -14
