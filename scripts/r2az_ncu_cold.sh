# ncu of the cold-regime sweep kernel (last 200 sweeps of the C2 schedule, forced mode 0)
set -u
mkdir -p gpurun_out
cat > /tmp/cold_run.py <<'P'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import isingmodel_jl_b200 as pkg
from isingmodel_jl_b200 import _lib, synth
N, R = 1024, 4096
ctx = _lib.context(0)
J = synth.sk_J(N, 2)
T = synth.geometric_schedule(2.0, 0.05, 1000)
ens = _lib.Ensemble(_lib.Model.dense(ctx, J, np.zeros(N), _lib.PREC_F64), R)
ens.set_spins(synth.spins(3, R, N))
os.environ.pop("ISB_SSF_MODE", None)
ens.ssf_run(_lib.RULE_GLAUBER, 800 * N, seed=1, T=T[:800], steps_per_T=N)          # anneal down to sweep 800 (product path)
os.environ["ISB_SSF_MODE"] = "0"
for rep in range(2):
    ens.ssf_run(_lib.RULE_GLAUBER, 200 * N, seed=1, step_offset=800 * N, T=T[800:], steps_per_T=N)
    print("cold kernel, sweeps 800-1000: %.2f ms, acceptance %.5f" % (ens.last_stats()["kernel_ms"], ens.last_stats()["flips"] / (200 * N * R)))
P
python /tmp/cold_run.py > gpurun_out/r2az_plain_cold.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:ssf_cold_kernel -s 1 -c 1 -f -o gpurun_out/r2az_c2_cold python /tmp/cold_run.py > gpurun_out/r2az_ncu_cold.log 2>&1; echo "ncu cold rc=$?"; tail -2 gpurun_out/r2az_plain_cold.log
