# What bounds the C3 contraction?  ring depth (pairs: 6 vs 4 slots), the sampling epilogue (probe build without it)
set -u
mkdir -p gpurun_out
run() { # tag lib cg workload prec
  ISING_B200_LIB=$2 ISB_TC_CG=$3 timeout 200 python bench.py --workload $4 --prec $5 --no-cpu-baseline > gpurun_out/r1i_bench_$4_$5_$1.json 2> gpurun_out/r1i_bench_$4_$5_$1.err
}
D=$PWD/isingmodel.jl_b200/libising_b200.so
run cg2 $D 2 c3 bf16x1
run cg2st4 $PWD/scratch_ab/lib_st4.so 2 c3 bf16x1
run cg1noepi $PWD/scratch_ab/lib_noepi.so 1 c3 bf16x1
run cg2noepi $PWD/scratch_ab/lib_noepi.so 2 c3 bf16x1
run cg1noepi $PWD/scratch_ab/lib_noepi.so 1 c4 bf16x1
run cg2noepi $PWD/scratch_ab/lib_noepi.so 2 c4 bf16x1
ISB_TC_CG=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:bip_tc -s 8 -c 1 -f -o gpurun_out/r1i_c3_cg2 \
  python bench.py --workload c3 --prec bf16x1 --steps 1 --warmup 1 --sca-steps 4 --no-cpu-baseline > gpurun_out/r1i_ncu_cg2.log 2>&1
