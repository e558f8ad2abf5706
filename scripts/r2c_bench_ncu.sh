# the default bench line (every config), the reference arm, the C caller, then ncu of the int8 tensor kernel (C4, C3) and
# of the sparse sweep kernel (C1)
set -u
mkdir -p gpurun_out
( timeout 1500 python bench.py > gpurun_out/r2c_bench_all.json 2> gpurun_out/r2c_bench_all.err ); echo "bench all rc=$? $(wc -c < gpurun_out/r2c_bench_all.json) bytes"
( timeout 600 python bench.py --impl reference > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err ); echo "bench ref rc=$?"
( timeout 300 python -m pytest tests/test_c_abi.py -m gpu -x -q ) > gpurun_out/r2c_c_abi.log 2>&1; echo "c abi rc=$?"; tail -3 gpurun_out/r2c_c_abi.log
C4="python bench.py --workload c4 --prec i8x3 --steps 1 --warmup 1 --sca-steps 100 --no-cpu-baseline"
C3="python bench.py --workload c3 --prec i8x3 --steps 1 --warmup 1 --sca-steps 3 --no-cpu-baseline"
C1="python bench.py --workload c1 --steps 1 --warmup 1 --c1-sweeps 300 --no-cpu-baseline"
$C4 > gpurun_out/r2c_plain_c4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:bip_tc_kernel -s 1 -c 1 -f -o gpurun_out/r2c_c4_i8x3 $C4 > gpurun_out/r2c_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
$C3 > gpurun_out/r2c_plain_c3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:bip_tc_kernel -s 8 -c 1 -f -o gpurun_out/r2c_c3_i8x3 $C3 > gpurun_out/r2c_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
$C1 > gpurun_out/r2c_plain_c1.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:ssf_sparse_kernel -s 1 -c 1 -f -o gpurun_out/r2c_c1_sparse $C1 > gpurun_out/r2c_ncu_c1.log 2>&1; echo "ncu c1 rc=$?"
ls -la gpurun_out/*.ncu-rep
