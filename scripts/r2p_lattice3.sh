set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_sparse.py -m gpu -x -q ) > gpurun_out/r2p_lattice_test.log 2>&1
echo "lattice tests rc=$?"; tail -5 gpurun_out/r2p_lattice_test.log
timeout 600 python bench.py --workload c1 --no-cpu-baseline --steps 3 > gpurun_out/r2p_bench_c1_lattice.json 2> gpurun_out/r2p_bench_c1_lattice.err; echo "c1 rc=$?"
python - <<'P'
import json
d=json.load(open("gpurun_out/r2p_bench_c1_lattice.json"))
print("value %.4g ms/step %.1f e2e %.4g Emean %.2f latency %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["mean_final_energy"], d["single_chain_latency"]))
P
C1="python bench.py --workload c1 --steps 1 --warmup 1 --c1-sweeps 300 --no-cpu-baseline"
$C1 > gpurun_out/r2p_plain_c1.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:ssf_lattice_kernel -s 1 -c 1 -f -o gpurun_out/r2p_c1_lattice $C1 > gpurun_out/r2p_ncu_c1.log 2>&1; echo "ncu c1 rc=$?"
