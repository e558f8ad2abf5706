# closing run of the round: the whole GPU suite, smoke, the driver's two bench commands, launch list, ncu of the final C4 kernel
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2ah_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 gpurun_out/r2ah_pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2ah_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2ah_smoke.log
( timeout 900 python bench.py > gpurun_out/r2ah_bench_default.json 2> gpurun_out/r2ah_bench_default.err ); echo "bench rc=$?"
( timeout 600 python bench.py --impl reference > gpurun_out/r2ah_bench_reference.json 2> gpurun_out/r2ah_bench_reference.err ); echo "reference arm rc=$?"
python - <<'P'
import json
d=json.load(open("gpurun_out/r2ah_bench_default.json"))
def show(k,w):
    r=w.get("roofline") or {}
    print(k, "value %.4g"%w["value"], "ms/step %.1f"%w["ms_per_step"], r.get("bound"), "frac %.3f"%r.get("frac",0), "e2e %.4g"%((w.get("e2e") or {}).get("value") or 0), "clk", (w.get("clocks") or {}).get("samples"), (w.get("clocks") or {}).get("sm_mhz"))
    for pn,pr in (w.get("precisions") or {}).items(): print("    ", pn, "value %.4g frac %.3f"%(pr["value"], pr["roofline"]["frac"]))
show("c2", d)
for k,w in d["workloads"].items(): show(k,w)
r=json.load(open("gpurun_out/r2ah_bench_reference.json")); print("reference arm: %.4g"%r["value"], r["cpu_baseline"]["cores"], "same config:", r["config"]==d["config"])
P
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sub-warmup 3"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2ah_bench_launches.csv $B > gpurun_out/r2ah_ncu_bench.log 2>&1; echo "launch list rc=$?"
C4="python bench.py --workload c4 --prec i8x3 --steps 1 --warmup 1 --sca-steps 100 --no-cpu-baseline"
$C4 > gpurun_out/r2ah_plain_c4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:bip_tc_kernel -s 1 -c 1 -f -o gpurun_out/r2ah_c4_i8x3 $C4 > gpurun_out/r2ah_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
