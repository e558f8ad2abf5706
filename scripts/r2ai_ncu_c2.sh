# fresh ncu capture of the sweep kernel (C2) of the closing library
set -u
mkdir -p gpurun_out
C2="python bench.py --workload c2 --steps 1 --warmup 1 --sweeps 200 --no-cpu-baseline"
$C2 > gpurun_out/r2ai_plain_c2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:ssf_kernel -s 1 -c 1 -f -o gpurun_out/r2ai_c2_ssf $C2 > gpurun_out/r2ai_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
