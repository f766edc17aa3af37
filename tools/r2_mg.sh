#!/bin/bash
# multi-GPU call on an N-GPU box: bash tools/r2_mg.sh N   (tests that fit N GPUs, then bench --gpus N through torchrun)
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
timeout 900 python -m pytest tests/test_mgpu_gpu.py tests/test_multigpu_gpu.py -m gpu -q > gpurun_out/pytest_multi_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_multi_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench$N rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${N}gpu.json"))
    print($N, "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", d["e2e"] and round(d["e2e"]["value"],1), d["e2e"] and round(d["e2e"]["ms_per_step"],2), "parity", d["parity_check"]["bit_identical"])
    c=d["strong_c5"]; print("c5", round(c["ms_per_step"],1), round(c["tflops"],1), "eff", c["efficiency"], "parity", c["parity_check"]["bit_identical"], "e2e", c.get("e2e") and (round(c["e2e"]["value"],1), c["e2e"]["parity_check"]["bit_identical"]))
except Exception as e:
    print("bench $N parse failed", e)
PY
tail -3 gpurun_out/bench_${N}gpu.err
