# First gpurun call of the next round: everything prepared after round 1's GPU budget was spent.
#   gpurun --timeout 900 -- 'bash scripts/r2_first_call.sh'     (build the epi12 variant first: scripts/build_variant.sh epi12 -DISB_TC_EPI_WARPS=12)
set -u
mkdir -p gpurun_out
# 1. sparse sweep kernel with register-prefetched CSR rows: parity, then C1 both ways
( ISB_TEST_UNVERIFIED=1 timeout 300 python -m pytest tests/test_gpu_sparse.py -m gpu -q -k prefetch ) > gpurun_out/r2a_prefetch_test.log 2>&1
for pf in 0 1; do
  ISB_SPARSE_PREFETCH=$pf timeout 300 python bench.py --workload c1 --no-cpu-baseline > gpurun_out/r2a_bench_c1_prefetch$pf.json 2>/dev/null
done
# 2. fp16 terms: throughput (expected = the bf16x2 figures)
for w in c3 c4; do for prec in fp16x2 bf16x2; do
  timeout 200 python bench.py --workload $w --prec $prec --no-cpu-baseline > gpurun_out/r2a_bench_${w}_${prec}.json 2>/dev/null
done; done
# 3. TMEM accumulator layouts (decides the next step on C4)
( cd scripts/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_layout_probe tmem_layout_probe.cu \
  && timeout 60 ./tmem_layout_probe ) > gpurun_out/r2a_tmem_layout.txt 2>&1
# 4. cold sweeps: fields in registers vs in shared memory
timeout 300 python scripts/ssf_cold_sparse_probe.py > gpurun_out/r2a_cold_probe.txt 2>&1
# 5. 12 epilogue warps (128 registers, no spills) vs 16
if [ -f scratch_ab/lib_epi12.so ]; then for w in c3 c4; do
  ISING_B200_LIB=$PWD/scratch_ab/lib_epi12.so timeout 200 python bench.py --workload $w --prec bf16x1 --no-cpu-baseline > gpurun_out/r2a_bench_${w}_bf16x1_epi12.json 2>/dev/null
  timeout 200 python bench.py --workload $w --prec bf16x1 --no-cpu-baseline > gpurun_out/r2a_bench_${w}_bf16x1_epi16.json 2>/dev/null
done; fi
ls -la gpurun_out | grep r2a
