#!/bin/bash
# N-GPU A/B of the two A transports (multigpu.py): NCCL broadcast vs CUDA-IPC copy-engine pulls
N=${1:-2}
run() { tag=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --warmup 3 --no-e2e "$@" > gpurun_out/mg_$tag.json 2> gpurun_out/mg_$tag.err; python -c "import json;d=json.load(open('gpurun_out/mg_$tag.json'));print('$tag', round(d['ms_per_step'],3), round(d['value'],2), d['config']['parallelism'])" || tail -5 gpurun_out/mg_$tag.err; }
run nccl --steps 10
run p2p --steps 10 --bcast p2p
run nccl2 --steps 10
run p2p2 --steps 10 --bcast p2p
if [ "${2:-}" = "c5" ]; then
run c5_nccl --workload c5 --steps 3
run c5_p2p --workload c5 --steps 3 --bcast p2p
fi
