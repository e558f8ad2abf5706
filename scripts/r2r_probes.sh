# where does a C4 half-step go?  contraction only / epilogue only / no Philox probes on the leaner epilogue, and an ncu capture
set -u
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 10 $ARGS > gpurun_out/r2r_bench_${tag}.json 2> gpurun_out/r2r_bench_${tag}.err
  echo "$tag rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2r_bench_${tag}.json')); r=d['roofline']
print('  value %.4g frac %.3f half-step %.4f ms clocks %s' % (d['value'], r['frac'], r['kernel_ms_per_half_step'], d['clocks']['sm_mhz']))"
}
ARGS="--workload c4 --prec i8x3"
run c4_new A=1
run c4_k1 ISING_B200_LIB=$PWD/scratch_ab/lib_k1.so
run c4_noepi ISING_B200_LIB=$PWD/scratch_ab/lib_noepi.so
run c4_nophilox ISING_B200_LIB=$PWD/scratch_ab/lib_nophilox.so
run c4_fine2 ISB_TC_FINE=2
run c4_fine0 ISB_TC_FINE=0
run c4_single ISB_TC_CG=1
C4="python bench.py --workload c4 --prec i8x3 --steps 1 --warmup 1 --sca-steps 100 --no-cpu-baseline"
$C4 > gpurun_out/r2r_plain_c4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:bip_tc_kernel -s 1 -c 1 -f -o gpurun_out/r2r_c4_i8x3 $C4 > gpurun_out/r2r_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
