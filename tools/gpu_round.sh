#!/bin/bash
# One gpurun call: tests, sweep, bench, ncu launch list + full capture of the top kernels.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tests|notests] [bench|nobench] [ncu targets...]'
#   ncu target = "<name>:<kernel regex>:<dtype> <M> <N> <K> <selector>"
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
if [ "${1:-tests}" = "tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
  tail -25 gpurun_out/pytest_gpu.log
fi
timeout 600 python tools/sweep.py ${SWEEP_ARGS:-} > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/sweep.log | cut -c1-2500
if [ "${2:-bench}" = "bench" ]; then
  timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; echo "ref rc=$?"
  cat gpurun_out/bench_ref.json
  # ncu: launch list of the bench command
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
fi
shift 2 2>/dev/null
for tgt in "$@"; do
  name="${tgt%%:*}"; rest="${tgt#*:}"; regex="${rest%%:*}"; args="${rest#*:}"
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:${regex}" -s 1 -c 1 -o "gpurun_out/prof_${name}" -f \
     python tools/ncu_target.py ${args} 2 > "gpurun_out/ncu_${name}.log" 2>&1; echo "ncu ${name} rc=$?"
done
ls -la gpurun_out | tail -30
