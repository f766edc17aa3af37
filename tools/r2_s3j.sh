#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k "skinny" > gpurun_out/pytest_skinny.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_skinny.log
timeout 600 python tools/skinny_compare.py > gpurun_out/skinny_compare.txt 2>&1; cat gpurun_out/skinny_compare.txt | tail -14 | cut -c1-250
JBLAS_B200_NO_PDL=1 timeout 600 python tools/skinny_compare.py 2>&1 | head -2 | cut -c1-250
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:gemm_skinny" -s 7 -c 1 -o gpurun_out/prof_skinny_lib -f python tools/ncu_target.py float64 65536 64 64 auto 12 6 > gpurun_out/ncu_skinny_lib.log 2>&1; echo "ncu skinny rc=$?"
python tools/ncu_summary.py gpurun_out/prof_skinny_lib.ncu-rep gpurun_out/sum_skinny_lib.txt "FP64 65536x64x64 AUTO = dmma_skinny_f64_16x16_xreg_team_w16, COLD operands (6 rotating sets = 403 MB, 8th launch captured)" > /dev/null 2>&1
grep -E "kernel:|time_duration|tensor_cycles|dram__bytes|cycles_elapsed" gpurun_out/sum_skinny_lib.txt
timeout 300 ncu --set full --clock-control none -k "regex:gemm_skinny" -s 3 -c 1 -o gpurun_out/prof_skinny_1m -f python tools/ncu_target.py float64 1000000 64 64 auto 6 2 > gpurun_out/ncu_skinny_1m.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_skinny_1m.ncu-rep gpurun_out/sum_skinny_1m.txt "FP64 1000000x64x64 AUTO = dmma_skinny_f64_16x16_xreg_team_w16 (2 operand sets of 1 GB)" > /dev/null 2>&1
grep -E "time_duration|tensor_cycles|dram__bytes" gpurun_out/sum_skinny_1m.txt; rm -f gpurun_out/prof_skinny_1m.ncu-rep
