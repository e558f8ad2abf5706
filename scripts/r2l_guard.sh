set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sparse.py tests/test_gpu_reference_tests.py -m gpu -x -q ) > gpurun_out/r2l_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r2l_parity.log
for g in default 0; do
  if [ $g = 0 ]; then export ISB_SSF_GUARD=0; fi
  timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r2l_bench_c2_guard_$g.json 2> gpurun_out/r2l_bench_c2_$g.err; echo "c2 guard=$g rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2l_bench_c2_guard_$g.json')); print('  c2 value %.4g ms %.1f launches %d' % (d['value'], d['ms_per_step'], d['gpu_launches']))"
done
