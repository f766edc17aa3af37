#!/bin/bash
N=${1:-8}
run() { tag=$1; shift; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --warmup 3 "$@" > gpurun_out/mg${N}_$tag.json 2> gpurun_out/mg${N}_$tag.err; python -c "import json;d=json.load(open('gpurun_out/mg${N}_$tag.json'));print('$tag', round(d['ms_per_step'],3), round(d['value'],2), d['config']['parallelism'])" || tail -5 gpurun_out/mg${N}_$tag.err; }
run p2p --steps 10 --bcast p2p --no-e2e
run nccl --steps 10 --no-e2e
run c5_p2p --workload c5 --steps 3 --bcast p2p --no-e2e
