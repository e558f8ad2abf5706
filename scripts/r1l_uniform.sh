# Warp-uniform producer / MMA issuer (elect.sync): tests, then bench lines for single CTAs and pairs
set -u
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_rowshard.py -m gpu -x -q --timeout 120 -k "tc or bip or c4 or c3 or shard" ) > gpurun_out/r1l_test.log 2>&1
echo "rc=$?" >> gpurun_out/r1l_test.log
run() { # tag cg workload prec
  ISB_TC_CG=$2 timeout 200 python bench.py --workload $3 --prec $4 --no-cpu-baseline > gpurun_out/r1l_bench_$3_$4_$1.json 2> gpurun_out/r1l_bench_$3_$4_$1.err
}
run cg1 1 c3 bf16x1
run cg2 2 c3 bf16x1
run cg1 1 c3 bf16x3
run cg2 2 c3 bf16x3
run cg1 1 c4 bf16x1
run cg2 2 c4 bf16x1
run cg1 1 c4 bf16x3
run cg2 2 c4 bf16x3
