#!/bin/bash
# ncu --set full of a tall-skinny probe variant (index $1, rows $2), cold operands; summary on the box, report kept for the source page
set -u
mkdir -p gpurun_out
V=${1:-1}
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:gemm_skinny" -s 8 -c 1 -o gpurun_out/prof_skinny_$V -f ./tools/skinny_probe ${2:-65536} 64 64 $V > gpurun_out/ncu_skinny_$V.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_skinny_$V.ncu-rep gpurun_out/sum_skinny_$V.txt "tall-skinny probe variant $V, M = ${2:-65536}" > /dev/null 2>&1
grep -E "kernel:|time_duration|tensor_cycles|dram__bytes|issue_active|stalled|cycles_elapsed" gpurun_out/sum_skinny_$V.txt
