# closing run (after the two-colour sweep order and the segmented delivery and the cold kernel of the sweeps): the whole GPU suite, smoke, the driver's two bench commands
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2bf_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 gpurun_out/r2bf_pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2bf_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2bf_smoke.log
( timeout 900 python bench.py > gpurun_out/r2bf_bench_default.json 2> gpurun_out/r2bf_bench_default.err ); echo "bench rc=$?"
( timeout 600 python bench.py --impl reference > gpurun_out/r2bf_bench_reference.json 2> gpurun_out/r2bf_bench_reference.err ); echo "reference arm rc=$?"
python - <<'P'
import json
d=json.load(open("gpurun_out/r2bf_bench_default.json"))
def show(k,w):
    r=w.get("roofline") or {}
    print(k, "value %.4g"%w["value"], "ms/step %.1f"%w["ms_per_step"], r.get("bound"), "frac %.3f"%r.get("frac",0), "burst %.3f"%r.get("frac_of_burst",0), "traffic", r.get("traffic"), "e2e %.4g"%((w.get("e2e") or {}).get("value") or 0), "clk", (w.get("clocks") or {}).get("samples"), (w.get("clocks") or {}).get("sm_mhz"))
    for pn,pr in (w.get("precisions") or {}).items(): print("    ", pn, "value %.4g frac %.3f"%(pr["value"], pr["roofline"]["frac"]))
show("c2", d)
for k,w in d["workloads"].items(): show(k,w)
print("c1 checkerboard", d["workloads"]["c1"]["checkerboard_order"]["value"])
r=json.load(open("gpurun_out/r2bf_bench_reference.json")); print("reference arm: %.4g"%r["value"], r["cpu_baseline"]["cores"], "same config:", r["config"]==d["config"], "e2e ratio %.0f" % (d["e2e"]["value"]/r["e2e"]["value"]))
P
