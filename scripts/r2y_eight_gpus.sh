# 8 GPUs: the NCCL / ABI row-shard parity tests at the widest rank count the test supports, then the default bench line
# under torchrun exactly as the driver launches it (C2 headline + C5 sub-result, N = 65536), then C5 alone per exchange mode
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -8
( timeout 600 python -m pytest tests/test_gpu_rowshard.py -m gpu -q -x ) > gpurun_out/r2y_rowshard.log 2>&1; echo "rowshard rc=$?"; tail -3 gpurun_out/r2y_rowshard.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
( timeout 900 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2y_bench_all_8gpu.json 2> gpurun_out/r2y_bench_all_8gpu.err ); echo "bench all 8gpu rc=$?"
python - <<'P'
import json
d=json.load(open("gpurun_out/r2y_bench_all_8gpu.json"))
print("C2 8 GPUs: value %.4g ms/step %.1f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
w=d["workloads"]["c5"]; r=w["roofline"]
print("C5: value %.4g half-step %.4f ms frac %.3f frac_of_fused_target %.3f exchange %s e2e %.4g" % (w["value"], r["kernel_ms_per_half_step"], r["frac"], r["frac_of_fused_target"], w["config"].get("exchange"), (w.get("e2e") or {}).get("value", 0)))
for k,v in (w.get("precisions") or {}).items(): print("   ", k, "value %.4g frac %.3f fused %.3f" % (v["value"], v["roofline"]["frac"], v["roofline"]["frac_of_fused_target"]))
P
for ex in abi-nccl copy nccl; do
  ISB_C5_EXCHANGE=$ex timeout 300 $TR bench.py --gpus 8 --workload c5 --prec i8x3 --steps 5 --warmup 3 > gpurun_out/r2y_c5_i8x3_${ex}_8gpu.json 2> gpurun_out/r2y_c5_i8x3_${ex}_8gpu.err
  echo "c5 $ex rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2y_c5_i8x3_${ex}_8gpu.json')); r=d['roofline']
print('  value %.4g half-step %.4f ms frac_of_fused_target %.3f frac %.3f exchange %s' % (d['value'], r['kernel_ms_per_half_step'], r['frac_of_fused_target'], r['frac'], d['config']['exchange']))"
done
