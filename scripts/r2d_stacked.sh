# stacked int8 planes (one MMA of N = P x bn per K block): parity, then throughput with the tile width swept
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_i8.py tests/test_gpu_rowshard.py -m gpu -x -q ) > gpurun_out/r2d_i8_test.log 2>&1
echo "i8 tests rc=$?"; tail -5 gpurun_out/r2d_i8_test.log
( timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s -k "c3_depth" ) > gpurun_out/r2d_c3depth.log 2>&1
echo "c3 depth rc=$?"; grep "decisions differ" gpurun_out/r2d_c3depth.log
for w in c3 c4; do for bn in 0 48 64 80; do
  ISB_I8_BN=$bn timeout 300 python bench.py --workload $w --prec i8x3 --no-cpu-baseline --steps 10 > gpurun_out/r2d_bench_${w}_i8x3_bn${bn}.json 2> gpurun_out/r2d_bench_${w}_bn${bn}.err
  echo "$w bn=$bn rc=$?"; python -c "
import json,sys
d=json.load(open('gpurun_out/r2d_bench_${w}_i8x3_bn${bn}.json')); r=d['roofline']
print('  value %.4g frac %.3f half-step %.4f ms clocks %s' % (d['value'], r['frac'], r['kernel_ms_per_half_step'], d['clocks']['sm_mhz']))"
done; done
for w in c3 c4; do
  timeout 300 python bench.py --workload $w --prec i8x2,i8x4 --no-cpu-baseline --steps 10 > gpurun_out/r2d_bench_${w}_i8x2x4.json 2> gpurun_out/r2d_bench_${w}_i8x2x4.err; echo "$w i8x2,i8x4 rc=$?"
done
