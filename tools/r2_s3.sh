#!/bin/bash
# round-2 (session 3) 1-GPU call: smoke, GPU tests, bench (both arms), tall-skinny probe on cold operands
set -u
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
./tools/skinny_probe > gpurun_out/skinny_probe.txt 2>&1; echo "skinny rc=$?"; cat gpurun_out/skinny_probe.txt
SKINNY_TRACE=1 ./tools/skinny_probe 65536 64 64 0 > gpurun_out/skinny_trace.txt 2>&1; cat gpurun_out/skinny_trace.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; cat gpurun_out/bench_1gpu.json; tail -5 gpurun_out/bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
