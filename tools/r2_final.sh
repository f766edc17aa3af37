#!/bin/bash
# final 1-GPU round of round 2: smoke, all GPU tests, bench (both arms), launch list of the bench command, sanitizer on the new kernels
set -u
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_1gpu.json; tail -3 gpurun_out/bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --workload c4b --no-cpu-baseline > gpurun_out/bench_c4b.json 2> gpurun_out/bench_c4b.err; echo "c4b rc=$?"; cut -c1-900 gpurun_out/bench_c4b.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python tools/launch_summary.py gpurun_out/launches.csv gpurun_out/launches_summary.txt > /dev/null 2>&1; head -12 gpurun_out/launches_summary.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "skinny and not 65536 and not 40000" > gpurun_out/sanitizer_skinny.log 2>&1; echo "memcheck skinny rc=$?"; tail -4 gpurun_out/sanitizer_skinny.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_tf32x3_gpu.py -m gpu -q -x -k "pair and not 4096" > gpurun_out/sanitizer_pair.log 2>&1; echo "memcheck pair rc=$?"; tail -4 gpurun_out/sanitizer_pair.log
