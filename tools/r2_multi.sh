#!/bin/bash
# round-2 multi-GPU call (8-GPU box): topology, the 4- and 8-GPU cases of the mgpu / multigpu tests, bench at 8, 4 (and 2) GPUs
set -u
mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; free -g | head -2; } > gpurun_out/box8.txt 2>&1
timeout 900 python -m pytest tests/test_mgpu_gpu.py tests/test_multigpu_gpu.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_multi.log
for n in 8 4 2; do
  JBLAS_B200_TRACE=1 timeout 900 python bench.py --gpus $n > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err; echo "bench$n rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${n}gpu.json"))
    print($n, "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", d["e2e"] and round(d["e2e"]["value"],1), d["e2e"] and round(d["e2e"]["ms_per_step"],2), "parity", d["parity_check"]["bit_identical"], "c5", d["strong_c5"] and (round(d["strong_c5"]["ms_per_step"],1), d["strong_c5"]["efficiency"], d["strong_c5"]["parity_check"]["bit_identical"]))
except Exception as e:
    print("bench $n parse failed", e)
PY
  grep -v "trace\]" gpurun_out/bench_${n}gpu.err | tail -3
done
timeout 600 python bench.py --gpus 8 --bcast nccl --no-e2e > gpurun_out/bench_8gpu_nccl.json 2> gpurun_out/bench_8gpu_nccl.err; echo "bench8 nccl rc=$?"
