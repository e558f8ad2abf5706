# Round-1 closing measurement of HEAD: GPU tests, smoke, bench lines, launch list + one full ncu capture (C4 kernel).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r1d_gpu.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 ) > gpurun_out/r1d_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r1d_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r1d_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r1d_smoke.log
timeout 300 python bench.py > gpurun_out/r1d_bench_c2.json 2> gpurun_out/r1d_bench_c2.err
for w in c4 c3; do for prec in bf16x1 bf16x3; do
  timeout 200 python bench.py --workload $w --prec $prec --no-cpu-baseline > gpurun_out/r1d_bench_${w}_${prec}.json 2> gpurun_out/r1d_bench_${w}_${prec}.err
done; done
for f in 0 2 99; do
  ISB_TC_FINE=$f timeout 200 python bench.py --workload c4 --prec bf16x1 --no-cpu-baseline > gpurun_out/r1d_bench_c4_bf16x1_fine$f.json 2>/dev/null
done
timeout 200 python bench.py --workload c1 --no-cpu-baseline > gpurun_out/r1d_bench_c1.json 2> gpurun_out/r1d_bench_c1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r1d_c4_launches.csv \
  python bench.py --workload c4 --prec bf16x1 --steps 1 --warmup 1 --sca-steps 20 --no-cpu-baseline > gpurun_out/r1d_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bip_tc -s 1 -c 1 -f -o gpurun_out/r1d_c4_tc \
  python bench.py --workload c4 --prec bf16x1 --steps 1 --warmup 1 --sca-steps 20 --no-cpu-baseline > gpurun_out/r1d_ncu_full.log 2>&1
ls -la gpurun_out
