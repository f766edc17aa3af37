#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:fastmul_batched_f32_warp" -s 1 -c 1 -o gpurun_out/prof_fb32 -f python tools/ncu_target.py batched 16 32 14 2000000 2 float32 > gpurun_out/ncu_fb32.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_fb32.ncu-rep gpurun_out/sum_fb32.txt "batched fastmul! 16x32x14 f32 x 2e6: warp-private TMA-bulk staging + FFMA2" > /dev/null 2>&1; echo "summary rc=$?"
rm -f gpurun_out/prof_fb32.ncu-rep
grep -E "time_duration|dram__bytes|registers|warps_active|pipe_fma_cycles|wavefronts_mem_shared|dram_throughput" gpurun_out/sum_fb32.txt
timeout 300 python bench.py --workload fb32 --steps 100 > gpurun_out/bench_fb32.json 2> gpurun_out/bench_fb32.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_fb32.err; cut -c1-1500 gpurun_out/bench_fb32.json
