#!/bin/bash
# ncu --set full of the headline kernel (FP64 8192^3 AUTO), summarised on the box
set -u
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:gemm_dmma_tma" -s 1 -c 1 -o gpurun_out/prof_headline -f python tools/ncu_target.py float64 8192 8192 8192 auto 2 > gpurun_out/ncu_headline.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_headline.ncu-rep gpurun_out/sum_headline.txt "FP64 8192^3 AUTO = dmma_tma_f64_128x128x32_s3 with the dynamic tile scheduler (final round-1 build)" > /dev/null 2>&1; echo "summary rc=$?"
rm -f gpurun_out/prof_headline.ncu-rep
grep -E "kernel:|time_duration|tensor_cycles|dram__bytes|issue_active|barrier|long_scoreboard|sector_hit" gpurun_out/sum_headline.txt
