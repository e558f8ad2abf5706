"""Cold-regime probe of the sweep kernel (T = 0.08: ~0.3 % acceptance, on-demand row delivery) for ncu."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isingmodel_jl_b200 as pkg
from isingmodel_jl_b200 import _lib, synth
N, R = 1024, 4096
ctx = _lib.context(0)
J = synth.sk_J(N, 2)
ens = _lib.Ensemble(_lib.Model.dense(ctx, J, np.zeros(N), _lib.PREC_F64), R)
ens.set_spins(synth.spins(3, R, N))
T = synth.geometric_schedule(2.0, 0.08, 60)
ens.ssf_run(_lib.RULE_GLAUBER, 60 * N, seed=1, T=T, steps_per_T=N)          # anneal down first (untimed)
for rep in range(2):
    ens.ssf_run(_lib.RULE_GLAUBER, 100 * N, seed=1, step_offset=(60 + 100 * rep) * N, T=np.array([0.08]), steps_per_T=100 * N)
    st = ens.last_stats()
    print("cold: accept", st["flips"] / (100 * N * R), st["kernel_ms"] / 100 * 1e3, "us/sweep")
