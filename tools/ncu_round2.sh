set -u
mkdir -p gpurun_out
run() { name=$1; regex=$2; shift 2; timeout 300 ncu --set full --clock-control none --import-source on -k "regex:${regex}" -s 1 -c 1 -o "gpurun_out/prof_${name}" -f python tools/ncu_target.py "$@" > "gpurun_out/ncu_${name}.log" 2>&1; echo "ncu ${name} rc=$?"; }
run fb_dmma fastmul_batched_dmma batched 16 32 14 1000000 2
run fb_tma64 gemm_dmma_tma batched 64 64 64 60000 2
run fb_tiny fastmul_batched_dmma batched 8 8 8 4000000 2
run dmma_2048 gemm_dmma_tma float64 2048 2048 2048 auto 2
ls -la gpurun_out/*.ncu-rep | tail
