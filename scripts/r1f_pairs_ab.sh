# CTA pairs, second pass: polling waits + 6-slot ring (default build) vs 4 slots vs suspended waits.
set -u
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "cta_pairs or chain_resident" ) > gpurun_out/r1f_pairs_test.log 2>&1
echo "pairs rc=$?" >> gpurun_out/r1f_pairs_test.log
run() { # tag lib cg workload prec
  ISING_B200_LIB=$2 ISB_TC_CG=$3 timeout 200 python bench.py --workload $4 --prec $5 --no-cpu-baseline > gpurun_out/r1f_bench_$4_$5_$1.json 2> gpurun_out/r1f_bench_$4_$5_$1.err
}
D=$PWD/isingmodel.jl_b200/libising_b200.so
run cg2 $D 2 c3 bf16x1
run cg2 $D 2 c3 bf16x3
run cg2 $D 2 c4 bf16x1
run cg2 $D 2 c4 bf16x3
run cg2st4 $PWD/scratch_ab/lib_st4.so 2 c3 bf16x1
run cg2st4 $PWD/scratch_ab/lib_st4.so 2 c4 bf16x1
run cg2sleep $PWD/scratch_ab/lib_sleep.so 2 c3 bf16x1
run cg1 $D 1 c3 bf16x1
