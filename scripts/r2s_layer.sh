# layer-templated tile code in the sampling warps: parity, C4 / C3 throughput, ncu
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_i8.py tests/test_gpu_rowshard.py tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/r2s_test.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2s_test.log
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 10 $ARGS > gpurun_out/r2s_bench_${tag}.json 2> gpurun_out/r2s_bench_${tag}.err
  echo "$tag rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2s_bench_${tag}.json')); r=d['roofline']
print('  value %.4g frac %.3f half-step %.4f ms clocks %s' % (d['value'], r['frac'], r['kernel_ms_per_half_step'], d['clocks']['sm_mhz']))"
}
ARGS="--workload c4 --prec i8x3"
run c4_new A=1
ARGS="--workload c4 --prec i8x2"
run c4_i8x2 A=1
ARGS="--workload c4 --prec bf16x1"
run c4_bf16x1 A=1
ARGS="--workload c3 --prec i8x3"
run c3_new A=1
C4="python bench.py --workload c4 --prec i8x3 --steps 1 --warmup 1 --sca-steps 100 --no-cpu-baseline"
$C4 > gpurun_out/r2s_plain_c4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:bip_tc_kernel -s 1 -c 1 -f -o gpurun_out/r2s_c4_i8x3 $C4 > gpurun_out/r2s_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
