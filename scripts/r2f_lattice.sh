# lattice specialisation of the sparse path: parity, then C1 throughput and single-chain latency; C2 per-temperature profile
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_sparse.py tests/test_gpu_reference_tests.py -m gpu -x -q ) > gpurun_out/r2f_lattice_test.log 2>&1
echo "lattice tests rc=$?"; tail -15 gpurun_out/r2f_lattice_test.log
timeout 600 python bench.py --workload c1 --no-cpu-baseline --steps 3 > gpurun_out/r2f_bench_c1_lattice.json 2> gpurun_out/r2f_bench_c1_lattice.err; echo "c1 rc=$?"
ISB_LATTICE=0 timeout 600 python bench.py --workload c1 --no-cpu-baseline --steps 3 > gpurun_out/r2f_bench_c1_generic.json 2> gpurun_out/r2f_bench_c1_generic.err; echo "c1 generic rc=$?"
python - <<'P'
import json
for f in ("lattice","generic"):
    d=json.load(open(f"gpurun_out/r2f_bench_c1_{f}.json"))
    print(f, "value %.4g ms/step %.1f e2e %.4g Emean %.2f latency %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["mean_final_energy"], d["single_chain_latency"]))
P
timeout 600 python scripts/ssf_profile_schedule.py > gpurun_out/r2f_c2_schedule_profile.txt 2>&1; tail -25 gpurun_out/r2f_c2_schedule_profile.txt
