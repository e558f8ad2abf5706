"""Probe for DESIGN §7 item 3: would the cold sweeps of the C2 anneal run faster with the chains' fields in shared
memory (one wave of 24 chains per SM, the sparse kernel's configuration) than in registers (two waves of 14)?  The same
dense SK couplings are loaded once as a dense model and once as a neighbour-list model (every site a neighbour of every
other); both ensembles are annealed down, then the same cold sweeps are timed on each (bit-identical trajectories)."""
import os, sys
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isingmodel_jl_b200 as pkg  # noqa: F401
from isingmodel_jl_b200 import _lib, synth

N, R = 1024, 4096
ctx = _lib.context(0)
J = synth.sk_J(N, 2)
S0 = synth.spins(3, R, N)
T = synth.geometric_schedule(2.0, 0.05, 1000)
ens = {"dense (fields in registers)": _lib.Ensemble(_lib.Model.dense(ctx, J, np.zeros(N), _lib.PREC_F64), R),
       "neighbour lists (fields in shared memory)": _lib.Ensemble(_lib.Model.sparse(ctx, sp.csc_matrix(J), np.zeros(N)), R)}
first = None
for name, e in ens.items():
    e.set_spins(S0)
    e.ssf_run(_lib.RULE_GLAUBER, 600 * N, seed=1, T=T[:600], steps_per_T=N)   # anneal down (untimed; the hot part)
    for lo in (600, 800):
        e.ssf_run(_lib.RULE_GLAUBER, 200 * N, seed=1, step_offset=lo * N, T=T[lo:lo + 200], steps_per_T=N)
        st = e.last_stats()
        print(f"{name}: sweeps {lo}..{lo + 200}: acceptance {st['flips'] / (200 * N * R):.4f}, {st['kernel_ms'] / 200 * 1e3:.1f} us per sweep")
    S = e.get_spins()
    first = S if first is None else first
    print("  same final spins as the first model:", bool(np.array_equal(S, first)))
