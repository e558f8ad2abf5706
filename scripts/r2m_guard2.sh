set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "commensurate" ) > gpurun_out/r2m_tie_test.log 2>&1; echo "tie test rc=$?"; grep -E "commensurate|passed|failed|Error|assert" gpurun_out/r2m_tie_test.log | tail -12
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sparse.py tests/test_gpu_reference_tests.py tests/test_gpu_threads.py tests/test_gpu_lattice.py tests/test_c_abi.py -m gpu -x -q ) > gpurun_out/r2m_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r2m_parity.log
timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r2m_bench_c2.json 2> gpurun_out/r2m_bench_c2.err; echo "c2 rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2m_bench_c2.json')); print('  c2 value %.4g ms %.1f launches %d' % (d['value'], d['ms_per_step'], d['gpu_launches']))"
