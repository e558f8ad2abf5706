# leaner sampling epilogue of the int8 tensor kernel (per-half-step temperature constants, int32 plane recombination,
# K-tail MMAs skipped): parity, then C4 / C3 throughput with A/B switches
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_i8.py tests/test_gpu_rowshard.py tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/r2q_test.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2q_test.log
run() {  # tag, env..., -- args
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 10 $ARGS > gpurun_out/r2q_bench_${tag}.json 2> gpurun_out/r2q_bench_${tag}.err
  echo "$tag rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2q_bench_${tag}.json')); r=d['roofline']
print('  value %.4g frac %.3f half-step %.4f ms clocks %s' % (d['value'], r['frac'], r['kernel_ms_per_half_step'], d['clocks']['sm_mhz']))"
}
ARGS="--workload c4 --prec i8x3"
run c4_new A=1
run c4_comb1 ISB_I8_COMB=1
run c4_comb0 ISB_I8_COMB=0
run c4_epi16 ISING_B200_LIB=$PWD/scratch_ab/lib_epi16.so
ARGS="--workload c4 --prec i8x2"
run c4_i8x2 A=1
ARGS="--workload c3 --prec i8x3"
run c3_new A=1
