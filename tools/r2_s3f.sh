#!/bin/bash
set -u
mkdir -p gpurun_out
./tools/skinny_probe > gpurun_out/skinny_probe4.txt 2>&1; cat gpurun_out/skinny_probe4.txt
for v in 0 1 2; do ./tools/skinny_probe 300000 64 64 $v | tail -1; done
./tools/skinny_probe 100000 48 72 1 | tail -1; ./tools/skinny_probe 65536 32 64 6 | tail -1; ./tools/skinny_probe 1000 64 64 1 | tail -1; ./tools/skinny_probe 65536 64 128 1 | tail -1;  ./tools/skinny_probe 20 64 64 1 | tail -1
SKINNY_TRACE=1 ./tools/skinny_probe 65536 64 64 1
