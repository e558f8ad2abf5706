# sequential sweeps run segment by segment, streaming kernel while hot, plain kernel once cold: parity, then C2
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_reference_tests.py tests/test_gpu_sparse.py -m gpu -x -q ) > gpurun_out/r2ar_test.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2ar_test.log
for v in 1 0; do
  ISB_SSF_SEGMENT=$v timeout 300 python bench.py --workload c2 --no-cpu-baseline --steps 5 > gpurun_out/r2ar_c2_seg$v.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/r2ar_c2_seg$v.json')); print('c2 segment=$v value %.4g ms/step %.2f e2e %.4g launches %s E %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['e2e']['mean_final_energy']))"
done
