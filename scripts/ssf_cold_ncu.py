"""Two launches of the C2 sweep kernel: 800 annealing sweeps (hot -> cold), then 50 cold sweeps (T 0.104 -> 0.087, acceptance
0.35 %) — the second launch is the one to profile (ncu -k regex:ssf_kernel -s 1 -c 1)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isingmodel_jl_b200 as pkg  # noqa: F401
from isingmodel_jl_b200 import _lib, synth

N, R = 1024, 4096
ctx = _lib.context(0)
J = synth.sk_J(N, 2)
T = synth.geometric_schedule(2.0, 0.05, 1000)
ens = _lib.Ensemble(_lib.Model.dense(ctx, J, np.zeros(N), _lib.PREC_F64), R)
ens.set_spins(synth.spins(3, R, N))
ens.ssf_run(_lib.RULE_GLAUBER, 800 * N, seed=1, T=T[:800], steps_per_T=N)
ens.ssf_run(_lib.RULE_GLAUBER, 50 * N, seed=1, step_offset=800 * N, T=T[800:850], steps_per_T=N)
st = ens.last_stats()
print(f"cold launch: acceptance {st['flips'] / (50 * N * R):.4f}, {st['kernel_ms'] / 50 * 1e3:.1f} us per sweep")
