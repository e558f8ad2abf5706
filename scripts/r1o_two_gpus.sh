# 2 GPUs: the NCCL row-shard test and a short C5 bench (N = 16384 over 2 GPUs) on the pair / uniform-issuer kernel
set -u
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests/test_gpu_rowshard.py -m gpu -q --timeout 150 ) > gpurun_out/r1o_rowshard.log 2>&1
echo "rc=$?" >> gpurun_out/r1o_rowshard.log
for prec in bf16x1 bf16x3; do
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c5 --prec $prec --steps 3 --warmup 3 > gpurun_out/r1o_bench_c5_${prec}_2gpu.json 2> gpurun_out/r1o_bench_c5_${prec}_2gpu.err
done
