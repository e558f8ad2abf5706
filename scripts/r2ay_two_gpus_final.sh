# 2 GPUs, closing check of the multi-GPU bench path after the bench.py changes of the round (sustained / burst denominators,
# traffic fields): the driver's torchrun command, and a C2-only launch list for the share of the sweep kernel in a step
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
( timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2ay_bench_all_2gpu.json 2> gpurun_out/r2ay_bench_all_2gpu.err ); echo "bench all 2gpu rc=$?"
( timeout 300 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2ay_bench_ref_2gpu.json 2> gpurun_out/r2ay_bench_ref_2gpu.err ); echo "reference arm under torchrun rc=$?"; wc -l gpurun_out/r2ay_bench_ref_2gpu.json
python - <<'P'
import json
d=json.load(open("gpurun_out/r2ay_bench_all_2gpu.json"))
print("C2 2 GPUs: value %.4g ms/step %.1f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
w=d["workloads"]["c5"]; r=w["roofline"]
print("C5: value %.4g half-step %.4f ms frac %.3f (burst %.3f) frac_of_fused_target %.3f exchange %s e2e %.4g region %.1f s" % (w["value"], r["kernel_ms_per_half_step"], r["frac"], r["frac_of_burst"], r["frac_of_fused_target"], w["config"].get("exchange"), (w.get("e2e") or {}).get("value", 0), r["timed_region_s"]))
P
B="python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2ay_c2_launches.csv $B > gpurun_out/r2ay_ncu_c2.log 2>&1; echo "c2 launch list rc=$?"
