#!/bin/bash
set -u
mkdir -p gpurun_out
./tools/skinny_probe > gpurun_out/skinny_probe2.txt 2>&1; echo "skinny rc=$?"; cat gpurun_out/skinny_probe2.txt
SKINNY_TRACE=1 ./tools/skinny_probe 65536 64 64 3 > gpurun_out/skinny_trace2.txt 2>&1; cat gpurun_out/skinny_trace2.txt
./tools/skinny_probe 65521 64 64 3; ./tools/skinny_probe 100000 48 72 3; ./tools/skinny_probe 4096 64 64 3; ./tools/skinny_probe 300000 64 64 3
bash tools/group_m_sweep.sh > gpurun_out/group_m_sweep.txt 2>&1; cat gpurun_out/group_m_sweep.txt
