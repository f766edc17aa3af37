#!/bin/bash
set -u
mkdir -p gpurun_out
./tools/skinny_probe > gpurun_out/skinny_probe3.txt 2>&1; cat gpurun_out/skinny_probe3.txt
./tools/skinny_probe 300000 64 64 0; ./tools/skinny_probe 300000 64 64 2; ./tools/skinny_probe 100000 48 72 2; ./tools/skinny_probe 65536 32 64 2
SKINNY_TRACE=1 ./tools/skinny_probe 65536 64 64 2
