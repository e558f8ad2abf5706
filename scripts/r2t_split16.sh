# 16 sampling warps with a setmaxnreg register split (service warps 32, sampling warps 112 registers)
set -u
mkdir -p gpurun_out
export SPLIT=$PWD/scratch_ab/lib_split16.so
( ISING_B200_LIB=$SPLIT timeout 600 python -m pytest tests/test_gpu_i8.py tests/test_gpu_parity.py -m gpu -x -q -k "tc or i8 or bip" ) > gpurun_out/r2t_test.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2t_test.log
run() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --no-cpu-baseline --steps 10 $ARGS > gpurun_out/r2t_bench_${tag}.json 2> gpurun_out/r2t_bench_${tag}.err
  echo "$tag rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2t_bench_${tag}.json')); r=d['roofline']
print('  value %.4g frac %.3f half-step %.4f ms clocks %s' % (d['value'], r['frac'], r['kernel_ms_per_half_step'], d['clocks']['sm_mhz']))"
}
ARGS="--workload c4 --prec i8x3"
run c4_base A=1
run c4_split16 ISING_B200_LIB=$SPLIT
ARGS="--workload c4 --prec bf16x1"
run c4_bf16x1_split16 ISING_B200_LIB=$SPLIT
ARGS="--workload c3 --prec i8x3"
run c3_split16 ISING_B200_LIB=$SPLIT
