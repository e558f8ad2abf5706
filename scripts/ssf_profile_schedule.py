"""Per-temperature profile of the C2 workload: runs the annealing schedule in chunks and prints, per chunk,
the temperature range, acceptance rate and device time (used to tune the adaptive row delivery)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isingmodel_jl_b200 as pkg
from isingmodel_jl_b200 import _lib, synth

N, R, SWEEPS, CHUNK = 1024, 4096, 1000, 50
ctx = _lib.context(0)
J = synth.sk_J(N, 2)
T = synth.geometric_schedule(2.0, 0.05, SWEEPS)
ens = _lib.Ensemble(_lib.Model.dense(ctx, J, np.zeros(N), _lib.PREC_F64), R)
ens.set_spins(synth.spins(3, R, N))
tot = 0.0
for c in range(0, SWEEPS, CHUNK):
    ens.ssf_run(_lib.RULE_GLAUBER, CHUNK * N, seed=1, step_offset=c * N, T=T[c:c + CHUNK], steps_per_T=N)
    st = ens.last_stats()
    tot += st["kernel_ms"]
    print(f"sweeps {c:4d}-{c+CHUNK:4d}  T {T[c]:.3f}->{T[c+CHUNK-1]:.3f}  accept {st['flips']/(CHUNK*N*R):.4f}  "
          f"{st['kernel_ms']/CHUNK*1e3:8.1f} us/sweep  {CHUNK*N*R/st['kernel_ms']/1e6:8.2f} Gupd/s")
print("total ms", tot)
