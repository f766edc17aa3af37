#!/bin/bash
set -u
bash tools/r2_mg.sh 8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --bcast nccl --no-e2e --no-other-configs > gpurun_out/bench_8gpu_nccl.json 2> gpurun_out/bench_8gpu_nccl.err; echo "bench8 nccl rc=$?"; cut -c1-400 gpurun_out/bench_8gpu_nccl.json
