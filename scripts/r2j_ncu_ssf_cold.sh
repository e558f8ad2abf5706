set -u
mkdir -p gpurun_out
python scripts/ssf_cold_ncu.py > gpurun_out/r2j_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:ssf_kernel -s 1 -c 1 -f -o gpurun_out/r2j_ssf_cold python scripts/ssf_cold_ncu.py > gpurun_out/r2j_ncu.log 2>&1; echo "ncu rc=$?"; cat gpurun_out/r2j_plain.log
