# 2 GPUs: the step loop inside the library (isb_shard_run_*) against the emulation, NCCL and copy-engine exchange
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_rowshard.py -m gpu -q -x ) > gpurun_out/r2i_rowshard.log 2>&1; echo "rowshard rc=$?"; tail -15 gpurun_out/r2i_rowshard.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for ex in abi-copy abi-nccl copy; do
  ISB_C5_EXCHANGE=$ex timeout 300 $TR bench.py --gpus 2 --workload c5 --prec i8x3 --steps 5 --warmup 3 > gpurun_out/r2i_c5_i8x3_${ex}_2gpu.json 2> gpurun_out/r2i_c5_i8x3_${ex}_2gpu.err
  echo "c5 $ex rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2i_c5_i8x3_${ex}_2gpu.json')); r=d['roofline']
print('  value %.4g half-step %.4f ms frac_of_fused_target %.3f frac %.3f exchange %s e2e %.4g' % (d['value'], r['kernel_ms_per_half_step'], r['frac_of_fused_target'], r['frac'], d['config']['exchange'], d['e2e']['value']))"
done
