#!/bin/bash
# round-2 first GPU call: box topology, smoke, GPU tests, 1-GPU bench, 2-GPU bench
set -u
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.limit,pci.bus_id --format=csv; nvidia-smi topo -m; lscpu | head -30; ls /sys/devices/system/node/ | head; free -g; } > gpurun_out/box.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; cat gpurun_out/bench_1gpu.json; tail -5 gpurun_out/bench_1gpu.err
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  JBLAS_B200_TRACE=1 timeout 600 python bench.py --gpus 2 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; cat gpurun_out/bench_2gpu.json; grep -v "trace" gpurun_out/bench_2gpu.err | tail -5
fi
