# Closing run of the round on the committed kernels (CTA pairs default): whole GPU suite, smoke, the default bench line,
# its launch list, and ncu --set full of the tcgen05 kernel on C3 and C4.
set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q --timeout 300 ) > gpurun_out/r1m_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r1m_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r1m_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r1m_smoke.log
timeout 300 python bench.py > gpurun_out/r1m_bench_c2.json 2> gpurun_out/r1m_bench_c2.err
timeout 200 python bench.py --workload c3 --prec bf16x2 --no-cpu-baseline > gpurun_out/r1m_bench_c3_bf16x2.json 2>/dev/null
timeout 200 python bench.py --workload c4 --prec bf16x2 --no-cpu-baseline > gpurun_out/r1m_bench_c4_bf16x2.json 2>/dev/null
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1m_c2_default_launches.csv \
  python bench.py --no-cpu-baseline > gpurun_out/r1m_ncu_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:bip_tc -s 8 -c 1 -f -o gpurun_out/r1m_c3_tc \
  python bench.py --workload c3 --prec bf16x1 --steps 1 --warmup 1 --sca-steps 4 --no-cpu-baseline > gpurun_out/r1m_ncu_c3.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:bip_tc -s 1 -c 1 -f -o gpurun_out/r1m_c4_tc \
  python bench.py --workload c4 --prec bf16x1 --steps 1 --warmup 1 --sca-steps 20 --no-cpu-baseline > gpurun_out/r1m_ncu_c4.log 2>&1
