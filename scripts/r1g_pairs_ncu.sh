# ncu --set full of the tcgen05 half-step on C3 (one bf16 pass): single CTAs vs CTA pairs
set -u
mkdir -p gpurun_out
for cg in 2 1; do
ISB_TC_CG=$cg timeout 300 ncu --set full --clock-control none --import-source on -k regex:bip_tc -s 8 -c 1 -f -o gpurun_out/r1g_c3_cg$cg \
  python bench.py --workload c3 --prec bf16x1 --steps 1 --warmup 1 --sca-steps 4 --no-cpu-baseline > gpurun_out/r1g_ncu_cg$cg.log 2>&1
done
ls -la gpurun_out | grep r1g
