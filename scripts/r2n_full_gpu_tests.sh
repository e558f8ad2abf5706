set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2n_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -15 gpurun_out/r2n_pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2n_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2n_smoke.log
