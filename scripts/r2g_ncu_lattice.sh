set -u
mkdir -p gpurun_out
C1="python bench.py --workload c1 --steps 1 --warmup 1 --c1-sweeps 300 --no-cpu-baseline"
$C1 > gpurun_out/r2g_plain_c1.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:ssf_lattice_kernel -s 1 -c 1 -f -o gpurun_out/r2g_c1_lattice $C1 > gpurun_out/r2g_ncu_c1.log 2>&1; echo "ncu c1 rc=$?"
