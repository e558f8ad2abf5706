# Temperature load hoisted above the accumulator wait; producer / MMA issuer poll (default) vs all waits suspended
set -u
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q --timeout 120 -k "tc or bip or c4 or c3" ) > gpurun_out/r1j_test.log 2>&1
echo "rc=$?" >> gpurun_out/r1j_test.log
run() { # tag lib cg workload prec
  ISING_B200_LIB=$2 ISB_TC_CG=$3 timeout 200 python bench.py --workload $4 --prec $5 --no-cpu-baseline > gpurun_out/r1j_bench_$4_$5_$1.json 2> gpurun_out/r1j_bench_$4_$5_$1.err
}
D=$PWD/isingmodel.jl_b200/libising_b200.so
A=$PWD/scratch_ab/lib_allsleep.so
run cg1 $D 1 c4 bf16x1
run cg1 $D 1 c4 bf16x3
run cg2 $D 2 c4 bf16x1
run cg2 $D 2 c4 bf16x3
run cg1allsleep $A 1 c4 bf16x1
run cg1allsleep $A 1 c4 bf16x3
run cg1 $D 1 c3 bf16x1
run cg2 $D 2 c3 bf16x1
