# CTA-pair (cta_group::2) variant of the tcgen05 kernel: parity first, then the whole GPU suite, then A/B bench lines.
set -u
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "cta_pairs" ) > gpurun_out/r1e_pairs_test.log 2>&1
echo "pairs rc=$?" >> gpurun_out/r1e_pairs_test.log
( time timeout 900 python -m pytest tests -m gpu -q --timeout 300 ) > gpurun_out/r1e_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r1e_pytest_gpu.log
for cg in 1 2; do for w in c3 c4; do for prec in bf16x1 bf16x3; do
  ISB_TC_CG=$cg timeout 200 python bench.py --workload $w --prec $prec --no-cpu-baseline > gpurun_out/r1e_bench_${w}_${prec}_cg$cg.json 2> gpurun_out/r1e_bench_${w}_${prec}_cg$cg.err
done; done; done
