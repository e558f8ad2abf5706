# TMEM loads waited for after the Philox rounds: parity, cycle accounting, C4 / C3 throughput
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_i8.py tests/test_gpu_rowshard.py tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/r2x_test.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2x_test.log
for p in i8x3 bf16x1; do ISING_B200_LIB=$PWD/scratch_ab/lib_timing.so timeout 120 python scripts/tc_timing.py $p; done 2>&1 | tee gpurun_out/r2x_tc_timing.txt
run() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --no-cpu-baseline --steps 10 $ARGS > gpurun_out/r2x_bench_${tag}.json 2> gpurun_out/r2x_bench_${tag}.err
  echo "$tag rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2x_bench_${tag}.json')); r=d['roofline']
print('  value %.4g frac %.3f half-step %.4f ms clocks %s' % (d['value'], r['frac'], r['kernel_ms_per_half_step'], d['clocks']['sm_mhz']))"
}
ARGS="--workload c4 --prec i8x3"
run c4 A=1
ARGS="--workload c4 --prec i8x2"
run c4_i8x2 A=1
ARGS="--workload c4 --prec bf16x1"
run c4_bf16x1 A=1
ARGS="--workload c4 --prec fp16x2"
run c4_fp16x2 A=1
ARGS="--workload c3 --prec i8x3"
run c3 A=1
ARGS="--workload c3 --prec bf16x1"
run c3_bf16x1 A=1
