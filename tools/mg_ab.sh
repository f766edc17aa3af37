run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e $EXTRA > gpurun_out/mg_$tag.json 2> gpurun_out/mg_$tag.err; python -c "import json;d=json.load(open('gpurun_out/mg_$tag.json'));print('$tag', d['ms_per_step'], d['value'], d['config']['parallelism'])"; }
EXTRA="" run dyn A=1
EXTRA="" run static JBLAS_B200_STATIC_TILES=1
EXTRA="--first-panel-k 0" run dyn_nofirst A=1
EXTRA="--first-panel-k 0" run static_nofirst JBLAS_B200_STATIC_TILES=1
EXTRA="" run dyn2 A=1
