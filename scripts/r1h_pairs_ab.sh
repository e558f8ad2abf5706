# CTA pairs, third pass: plain (non-release.cluster) remote arrives; probe without the peer's arrive.
set -u
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "cta_pairs" ) > gpurun_out/r1h_pairs_test.log 2>&1
echo "pairs rc=$?" >> gpurun_out/r1h_pairs_test.log
run() { # tag lib cg workload prec
  ISING_B200_LIB=$2 ISB_TC_CG=$3 timeout 200 python bench.py --workload $4 --prec $5 --no-cpu-baseline > gpurun_out/r1h_bench_$4_$5_$1.json 2> gpurun_out/r1h_bench_$4_$5_$1.err
}
D=$PWD/isingmodel.jl_b200/libising_b200.so
run cg2 $D 2 c3 bf16x1
run cg2 $D 2 c3 bf16x3
run cg2 $D 2 c4 bf16x1
run cg2 $D 2 c4 bf16x3
run cg2noarr $PWD/scratch_ab/lib_noarr.so 2 c3 bf16x1
run cg2noarr $PWD/scratch_ab/lib_noarr.so 2 c4 bf16x1
