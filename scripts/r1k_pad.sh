# Operand pitch padding (power-of-two pitches -> +128 B) on/off, single CTAs and pairs
set -u
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q --timeout 120 -k "tc or bip or c4 or c3" ) > gpurun_out/r1k_test.log 2>&1
echo "rc=$?" >> gpurun_out/r1k_test.log
run() { # tag pad cg workload prec
  ISB_TC_PAD=$2 ISB_TC_CG=$3 timeout 200 python bench.py --workload $4 --prec $5 --no-cpu-baseline > gpurun_out/r1k_bench_$4_$5_$1.json 2> gpurun_out/r1k_bench_$4_$5_$1.err
}
run pad_cg1 1 1 c3 bf16x1
run pad_cg2 1 2 c3 bf16x1
run nopad_cg2 0 2 c3 bf16x1
run pad_cg2 1 2 c3 bf16x3
run pad_cg1 1 1 c4 bf16x1
run pad_cg2 1 2 c4 bf16x1
run pad_cg2 1 2 c4 bf16x3
