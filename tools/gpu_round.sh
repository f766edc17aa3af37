#!/bin/bash
# One gpurun call: tests, sweep, bench, ncu launch list + full capture of the top kernels.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tests|notests]'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
if [ "${1:-tests}" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
  tail -15 gpurun_out/pytest_gpu.log
fi
timeout 600 python tools/sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/sweep.log | cut -c1-1500
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; echo "ref rc=$?"
cat gpurun_out/bench_ref.json
# ncu: launch list of the bench command, then full captures of the two FP64 candidates
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma_tma -s 1 -c 1 -o gpurun_out/prof_dmma_tma -f \
   python tools/ncu_target.py float64 8192 8192 8192 auto 2 > gpurun_out/ncu_dmma_tma.log 2>&1; echo "ncu dmma_tma rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_simt -s 1 -c 1 -o gpurun_out/prof_simt_f32 -f \
   python tools/ncu_target.py float32 8192 8192 8192 auto 2 > gpurun_out/ncu_simt_f32.log 2>&1; echo "ncu simt_f32 rc=$?"
ls -la gpurun_out
