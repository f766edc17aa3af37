#!/bin/bash
# full 1-GPU round: smoke, all GPU tests, bench (both arms), launch list, ncu captures of the new kernels
set -u
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_1gpu.json; tail -3 gpurun_out/bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python tools/launch_summary.py gpurun_out/launches.csv gpurun_out/launches_summary.txt > /dev/null 2>&1; head -30 gpurun_out/launches_summary.txt
# full captures
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:gemm_tf32x3_pair" -s 1 -c 1 -o gpurun_out/prof_tf32x3_pair -f python tools/ncu_target.py float32 8192 8192 8192 tf32x3 3 > gpurun_out/ncu_pair.log 2>&1; echo "ncu pair rc=$?"
python tools/ncu_summary.py gpurun_out/prof_tf32x3_pair.ncu-rep gpurun_out/sum_tf32x3_pair.txt "FP32 8192^3 3xTF32 AUTO = tf32x3_tcgen05_2cta_f32_256x256x32_s3 (cta_group::2), kernel only" > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:gemm_skinny" -s 7 -c 1 -o gpurun_out/prof_skinny_lib -f python tools/ncu_target.py float64 65536 64 64 auto 12 6 > gpurun_out/ncu_skinny_lib.log 2>&1; echo "ncu skinny rc=$?"
python tools/ncu_summary.py gpurun_out/prof_skinny_lib.ncu-rep gpurun_out/sum_skinny_lib.txt "FP64 65536x64x64 AUTO = dmma_skinny_f64_16x64_xres_w12, COLD operands (6 rotating sets = 403 MB, 8th launch captured)" > /dev/null 2>&1
grep -E "kernel:|time_duration|tensor_cycles|dram__bytes|cycles_elapsed" gpurun_out/sum_tf32x3_pair.txt gpurun_out/sum_skinny_lib.txt
rm -f gpurun_out/prof_tf32x3_pair.ncu-rep
