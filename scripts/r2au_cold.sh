# third kernel of the segmented sweep run (fields in shared memory, 28 chains per SM): parity, then C2 with thresholds
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_reference_tests.py -m gpu -x -q ) > gpurun_out/r2au_test.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2au_test.log
for tc in 0.04 0.0 0.02 0.08 0.15; do
  ISB_SSF_COLD_THR=$tc timeout 300 python bench.py --workload c2 --no-cpu-baseline --steps 5 > gpurun_out/r2au_c2_cold$tc.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/r2au_c2_cold$tc.json')); print('c2 cold_thr=$tc value %.4g ms/step %.2f e2e %.4g launches %s E %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['e2e']['mean_final_energy']))"
done
