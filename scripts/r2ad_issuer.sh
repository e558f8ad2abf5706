# MMA issue loop without the division (16-bit formats), Philox keys from the parameter bank, if-chain instead of a jump table
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_i8.py tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/r2ad_test.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2ad_test.log
ISING_B200_LIB=$PWD/scratch_ab/lib_timing.so timeout 120 python scripts/tc_timing.py i8x3 2>&1 | grep -v Warning | tee gpurun_out/r2ad_tc_timing.txt
run() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --no-cpu-baseline --steps 10 $ARGS > gpurun_out/r2ad_bench_${tag}.json 2> gpurun_out/r2ad_bench_${tag}.err
  echo "$tag rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2ad_bench_${tag}.json')); r=d['roofline']
print('  value %.4g frac %.3f half-step %.4f ms clocks %s' % (d['value'], r['frac'], r['kernel_ms_per_half_step'], d['clocks']['sm_mhz']))"
}
ARGS="--workload c3 --prec bf16x1"
run c3_bf16x1 A=1
ARGS="--workload c3 --prec fp16x2"
run c3_fp16x2 A=1
ARGS="--workload c3 --prec i8x3"
run c3_i8x3 A=1
ARGS="--workload c4 --prec i8x3"
run c4_i8x3 A=1
ARGS="--workload c4 --prec bf16x1"
run c4_bf16x1 A=1
