#!/bin/bash
# 8-GPU round: weak-scaling headline (with e2e), BASELINE configs[4] strong-scaled with two panel widths
set -u
N=${1:-8}
trun() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
trun 29521 --steps 10 --warmup 3 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err; echo "c2 rc=$?"; cut -c1-400 gpurun_out/bench_g$N.json
if [ "$N" = "8" ]; then
trun 29522 --workload c5 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_c5_g8.json 2> gpurun_out/bench_c5_g8.err; echo "c5 rc=$?"; cut -c1-400 gpurun_out/bench_c5_g8.json
trun 29523 --workload c5 --steps 3 --warmup 3 --no-e2e --panel-k 4096 > gpurun_out/bench_c5_g8_p4096.json 2> gpurun_out/bench_c5_g8_p4096.err; echo "c5 p4096 rc=$?"; cut -c1-400 gpurun_out/bench_c5_g8_p4096.json
fi
