# ncu of the final kernels (one capture each) and the launch list of the default bench command
set -u
mkdir -p gpurun_out
C4="python bench.py --workload c4 --prec i8x3 --steps 1 --warmup 1 --sca-steps 100 --no-cpu-baseline"
C3="python bench.py --workload c3 --prec i8x3 --steps 1 --warmup 1 --sca-steps 3 --no-cpu-baseline"
C1="python bench.py --workload c1 --steps 1 --warmup 1 --c1-sweeps 300 --no-cpu-baseline"
$C4 > gpurun_out/r2o_plain_c4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:bip_tc_kernel -s 1 -c 1 -f -o gpurun_out/r2o_c4_i8x3 $C4 > gpurun_out/r2o_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
$C3 > gpurun_out/r2o_plain_c3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:bip_tc_kernel -s 8 -c 1 -f -o gpurun_out/r2o_c3_i8x3 $C3 > gpurun_out/r2o_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
$C1 > gpurun_out/r2o_plain_c1.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:ssf_lattice_kernel -s 1 -c 1 -f -o gpurun_out/r2o_c1_lattice $C1 > gpurun_out/r2o_ncu_c1.log 2>&1; echo "ncu c1 rc=$?"
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sub-warmup 1"
$B > gpurun_out/r2o_plain_bench.log 2>&1 && timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2o_bench_launches.csv $B > gpurun_out/r2o_ncu_bench.log 2>&1; echo "launch list rc=$?"
