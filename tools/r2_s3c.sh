#!/bin/bash
set -u
mkdir -p gpurun_out
./tools/skinny_probe 65536 64 64 -2 > gpurun_out/dmma_pattern.txt 2>&1; cat gpurun_out/dmma_pattern.txt
timeout 600 python tools/tf32x3_pair_check.py > gpurun_out/tf32x3_pair_check.txt 2>&1; echo "pair rc=$?"; tail -30 gpurun_out/tf32x3_pair_check.txt
