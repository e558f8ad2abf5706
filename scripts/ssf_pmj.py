"""+-J spin glass (couplings +-1/32: exactly representable in float, so ISB_PREC_AUTO stores J as float while the
fields and decisions stay Float64): the C2 schedule on the lossless-float storage path."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isingmodel_jl_b200 as pkg
from isingmodel_jl_b200 import _lib, synth
N, R, SWEEPS = 1024, 4096, 1000
ctx = _lib.context(0)
J = np.sign(synth.sk_J(N, 2)) / 32.0
T = synth.geometric_schedule(2.0, 0.05, SWEEPS)
for prec, name in ((_lib.PREC_AUTO, "auto (float J storage, Float64 fields)"), (_lib.PREC_F64, "f64 J storage")):
    ens = _lib.Ensemble(_lib.Model.dense(ctx, J, np.zeros(N), prec), R)
    for rep in range(2):
        ens.set_spins(synth.spins(3, R, N))
        ens.ssf_run(_lib.RULE_GLAUBER, SWEEPS * N, seed=1, T=T, steps_per_T=N)
        st = ens.last_stats()
    print(name, "ms", st["kernel_ms"], "Gupd/s", SWEEPS * N * R / st["kernel_ms"] / 1e6, "accept", st["flips"] / (SWEEPS * N * R))
