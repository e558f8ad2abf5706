# two-colour sweep order of the lattice kernel: parity against the oracle's site-list run, exact mean energy, throughput
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_sparse.py -m gpu -x -q ) > gpurun_out/r2am_test.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2am_test.log
timeout 600 python bench.py --workload c1 --no-cpu-baseline --steps 3 > gpurun_out/r2am_bench_c1.json 2> gpurun_out/r2am_bench_c1.err; echo "c1 rc=$?"
python - <<'P'
import json
d=json.load(open("gpurun_out/r2am_bench_c1.json"))
print("sequential: value %.4g single chain %.2f us/sweep E %.1f" % (d["value"], d["single_chain_latency"]["us_per_sweep"], d["e2e"]["mean_final_energy"]))
c=d["checkerboard_order"]; print("checkerboard: value %.4g single chain %.2f us/sweep E %.1f kernel %.1f ms" % (c["value"], c["single_chain_us_per_sweep"], c["mean_final_energy"], c["kernel_ms"]))
P
