"""Generates tests/golden/*.npz — golden trajectories for the spin-update hot path.

Provenance: Julia is not installed in this image (nor on the GPU boxes), so the reference package cannot be
executed; the vectors are produced by the CPU oracle (oracle/ising_oracle.c), which restates the reference
line by line, and are cross-checked here against the independent pure-Python twin (oracle/oracle_np.py)
before being written.  The seven fixtures the reference's own test-suite holds (test/runtests.jl:20-31) are
asserted in tests/test_oracle.py.  Run:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import oracle_np  # noqa: E402
import isingmodel_jl_b200  # noqa: E402,F401
from isingmodel_jl_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def fluct_for(rule, seed, n):
    if rule == oracle.GLAUBER:
        return synth.logistic(seed, n)
    if rule == oracle.METROPOLIS:
        return synth.exponential(seed, n)
    return np.zeros(n)


def ssf_case(name, J, h, s0, rule, nsteps, nodes, T, steps_per_T, seed, trace_every, twin=True, store_J=True):
    fl = fluct_for(rule, seed, nsteps)
    s, flips, E, M = oracle.ssf_run(rule, J, h, s0, nsteps, nodes=nodes, fluct=fl, T=T, steps_per_T=steps_per_T,
                                    trace_every=trace_every)
    if twin:
        s2, f2 = oracle_np.ssf_run(rule, J, h, s0, nsteps, nodes=nodes, fluct=fl, T=T, steps_per_T=steps_per_T)
        assert np.array_equal(s, s2) and flips == f2, name
    d = dict(rule=rule, h=h, s0=s0, nsteps=nsteps, nodes=np.array([]) if nodes is None else nodes, fluct=fl, T=T,
             steps_per_T=steps_per_T, trace_every=trace_every, s_final=s, flips=flips, E=E, M=M, J_sha=sha(J))
    if store_J:
        d["J"] = J
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "flips", flips, "E_final", E[-1] if len(E) else None)


def bip_case(name, W, h, b, s0, t0, rule, nsteps, T, seed, store_W=True, twin=True):
    nv, nh = W.shape
    gen = synth.logistic if rule == oracle.SCA else synth.exponential
    Fv, Fh = gen(seed, (nsteps, nv), 1), gen(seed, (nsteps, nh), 2)
    s, t, E = oracle.bip_run(rule, W, h, b, s0, t0, nsteps, Fv, Fh, T, want_E=True)
    if twin:
        s2, t2 = oracle_np.bip_run(rule, W, h, b, s0, t0, nsteps, Fv, Fh, T)
        assert np.array_equal(s, s2) and np.array_equal(t, t2), name
    d = dict(rule=rule, h=h, b=b, s0=s0, t0=t0, nsteps=nsteps, Fv=Fv, Fh=Fh, T=T, s_final=s, t_final=t, E=E,
             W_sha=sha(W))
    if store_W:
        d["W"] = W
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "E_final", E[-1])


def main():
    # --- 2-spin ferromagnet of test/runtests.jl:20
    J2 = np.array([[0.0, 1.0], [1.0, 0.0]])
    for rule, nm in ((0, "hopfield"), (1, "glauber"), (2, "metropolis")):
        ssf_case(f"ssf_2spin_{nm}", J2, np.zeros(2), np.array([-1, 1], dtype=np.int8), rule, 12,
                 synth.nodes(5, 2, 12), 10.0 ** (-np.arange(1, 13.0)), 1, 11 + rule, 1)
    # --- 3x3 periodic antiferromagnet of demo.jl:60-62, annealed
    J9 = synth.lattice_J(3, -1.0)
    for rule, nm in ((1, "glauber"), (2, "metropolis")):
        ssf_case(f"ssf_3x3_{nm}", J9, np.zeros(9), synth.spins(3, 1, 9)[0], rule, 400, synth.nodes(7, 9, 400),
                 synth.geometric_schedule(4.0, 0.05, 400), 1, 20 + rule, 10)
    # --- C1: 32x32 ferromagnet, Metropolis at T = 2.269, sequential sweeps (16 sweeps)
    J1 = synth.lattice_J(32)
    ssf_case("ssf_c1_32x32_metropolis", J1, np.zeros(1024), synth.spins(1, 1, 1024)[0], 2, 16 * 1024, None,
             np.array([2.269]), 16 * 1024, 1, 1024, twin=False, store_J=False)
    # --- SK N = 64 with a field, random sites, Glauber annealing + Hopfield quench
    J64 = synth.sk_J(64, 2)
    h64 = synth.gaussian(9, 64) * 0.1
    ssf_case("ssf_sk64_glauber", J64, h64, synth.spins(2, 1, 64)[0], 1, 3000, synth.nodes(8, 64, 3000),
             synth.geometric_schedule(2.0, 0.05, 30), 100, 31, 100)
    ssf_case("ssf_sk64_hopfield", J64, h64, synth.spins(2, 1, 64)[0], 0, 640, None, np.array([0.0]), 640, 32, 64)
    # --- C2 shape: SK N = 1024, Glauber, 2 annealing sweeps (J regenerated from the seed; sha stored)
    J1k = synth.sk_J(1024, 2)
    ssf_case("ssf_c2_sk1024_glauber", J1k, np.zeros(1024), synth.spins(4, 1, 1024)[0], 1, 2 * 1024, None,
             np.array([2.0, 1.0]), 1024, 33, 1024, twin=False, store_J=False)
    # --- bipartite: the 2x3 system of test/runtests.jl:28 and a C4-shaped one
    W23 = np.ones((2, 3))
    for rule, nm in ((0, "sca"), (1, "ma")):
        bip_case(f"bip_2x3_{nm}", W23, np.zeros(2), np.zeros(3), np.array([-1, 1], dtype=np.int8),
                 np.array([-1, 1, -1], dtype=np.int8), rule, 8, 10.0 ** (-np.arange(1, 9.0)), 40 + rule)
    W, h, b = synth.bipartite_W(784, 512, 4)
    for rule, nm in ((0, "sca"), (1, "ma")):
        bip_case(f"bip_c4_784x512_{nm}", W, h, b, synth.spins(5, 1, 784)[0], synth.spins(6, 1, 512)[0], rule, 4,
                 np.array([1.0, 1.0, 0.5, 0.25]), 50 + rule, store_W=False, twin=False)
    Ws, hs, bs = synth.bipartite_W(24, 17, 6, 0.5)
    for rule, nm in ((0, "sca"), (1, "ma")):
        bip_case(f"bip_24x17_{nm}", Ws, hs, bs, synth.spins(7, 1, 24)[0], synth.spins(8, 1, 17)[0], rule, 20,
                 synth.geometric_schedule(2.0, 0.1, 20), 60 + rule)


if __name__ == "__main__":
    main()
