set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "commensurate" ) > gpurun_out/r2k_tie_test.log 2>&1; echo "tie test rc=$?"; grep -E "commensurate|passed|failed|Error" gpurun_out/r2k_tie_test.log | tail -8
( timeout 900 python -m pytest tests/test_gpu_i8.py tests/test_gpu_parity.py -m gpu -x -q -k "not commensurate" ) > gpurun_out/r2k_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r2k_parity.log
for w in c4 c3; do
  timeout 300 python bench.py --workload $w --prec i8x3,bf16x1 --no-cpu-baseline --steps 10 > gpurun_out/r2k_bench_${w}.json 2> gpurun_out/r2k_bench_${w}.err; echo "$w rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/r2k_bench_${w}.json')); r=d['roofline']
print('  i8x3 value %.4g frac %.3f half-step %.4f ms' % (d['value'], r['frac'], r['kernel_ms_per_half_step']))
for k,v in d['precisions'].items(): print('  ',k,'value %.4g frac %.3f' % (v['value'], v['roofline']['frac']))"
done
timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r2k_bench_c2.json 2> gpurun_out/r2k_bench_c2.err; echo "c2 rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2k_bench_c2.json')); print('c2 value %.4g ms %.1f launches %d' % (d['value'], d['ms_per_step'], d['gpu_launches']))"
