"""Threading rule of the handle model (include/ising_b200.h, "Threads"): calls on one context are serialised by the
library, so two host threads may drive two ensembles of the SAME context; and isb_ens_clone gives the independent
copy that `deepcopy(ss)` of the reference's host object needs (test/runtests.jl:22-24,30-31)."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lib():
    from isingmodel_jl_b200 import _lib
    return _lib


def test_two_threads_two_ensembles_one_context(ctx, orc, synth):
    """Two threads, each looping over short runs of its own ensemble (dense sweeps / block Gibbs) on one shared context
    — shared stream, events and scratch buffers.  ctypes releases the GIL during the calls, so they do overlap on the
    host; every run must still reproduce the oracle trajectory."""
    L = _lib()
    N, R, chunks, per = 96, 24, 12, 96
    J, h = synth.sk_J(N, 31), synth.gaussian(32, N) * 0.1
    S0 = synth.spins(33, R, N)
    fl = synth.logistic(34, (R, chunks * per))
    T = synth.geometric_schedule(2.0, 0.3, chunks)
    ens_a = L.Ensemble(L.Model.dense(ctx, J, h, L.PREC_F64), R)
    ens_a.set_spins(S0)
    nv, nh, Rb, nst = 80, 48, 16, 24
    W, hv, bh = synth.bipartite_W(nv, nh, 35, 0.3)
    Sv0, Sh0 = synth.spins(36, Rb, nv), synth.spins(37, Rb, nh)
    Fv, Fh = synth.logistic(38, (nst, nv), 1), synth.logistic(38, (nst, nh), 2)
    Tb = synth.geometric_schedule(1.5, 0.2, nst)
    ens_b = L.Ensemble(L.Model.bipartite(ctx, W, hv, bh, L.PREC_F64), Rb)
    ens_b.set_spins(Sv0)
    ens_b.set_hidden(Sh0)
    errors = []

    def run_a():
        try:
            for c in range(chunks):
                ens_a.ssf_run(L.RULE_GLAUBER, per, start=(c * per) % N, fluct=np.ascontiguousarray(fl[:, c * per:(c + 1) * per]),
                              fluct_per_replica=True, T=T[c:c + 1], steps_per_T=per)
                ens_a.energy()
        except Exception as ex:  # noqa: BLE001
            errors.append(ex)

    def run_b():
        try:
            for k in range(nst):
                ens_b.bip_run(L.BIP_SCA, 1, Fv=Fv[k:k + 1], Fh=Fh[k:k + 1], T=Tb[k:k + 1])
                ens_b.local_field()
        except Exception as ex:  # noqa: BLE001
            errors.append(ex)

    ta, tb = threading.Thread(target=run_a), threading.Thread(target=run_b)
    ta.start()
    tb.start()
    ta.join()
    tb.join()
    assert not errors, errors
    Sa = ens_a.get_spins()
    for r in range(R):
        s, *_ = orc.ssf_run(orc.GLAUBER, J, h, S0[r], chunks * per, fluct=fl[r], T=T, steps_per_T=per)
        assert np.array_equal(s, Sa[r]), f"replica {r}"
    Sb, Hb = ens_b.get_spins(), ens_b.get_hidden()
    for r in range(Rb):
        s, t, _ = orc.bip_run(orc.SCA, W, hv, bh, Sv0[r], Sh0[r], nst, Fv, Fh, Tb)
        assert np.array_equal(s, Sb[r]) and np.array_equal(t, Hb[r]), f"chain {r}"


def test_error_text_is_per_thread(ctx):
    L = _lib()
    got = {}

    def bad(name, n):
        try:
            L.Model.dense(ctx, np.zeros((2, 2)), np.zeros(2), 77 + n)
        except L.IsbError as ex:
            got[name] = str(ex)

    t = threading.Thread(target=bad, args=("thread", 1))
    t.start()
    t.join()
    bad("main", 0)
    assert "prec" in got["thread"] and "prec" in got["main"]


@pytest.mark.parametrize("kind", ["dense", "sparse", "bipartite", "bipartite_i8"])
def test_clone_is_independent(ctx, orc, synth, kind):
    """deepcopy semantics: the clone starts from the same state, then the two evolve independently."""
    import scipy.sparse as sp
    L = _lib()
    R = 6
    if kind in ("dense", "sparse"):
        N = 64
        J = synth.lattice_J(8) if kind == "sparse" else synth.sk_J(N, 41)
        h = np.zeros(N)
        m = L.Model.sparse(ctx, sp.csc_matrix(J), h) if kind == "sparse" else L.Model.dense(ctx, J, h, L.PREC_F64)
        S0 = synth.spins(42, R, N)
        e = L.Ensemble(m, R)
        e.set_spins(S0)
        fl = synth.logistic(43, (R, 2 * N))
        e.ssf_run(L.RULE_GLAUBER, N, fluct=np.ascontiguousarray(fl[:, :N]), fluct_per_replica=True, T=np.array([1.0]), steps_per_T=N)
        c = e.clone()
        S1 = e.get_spins()
        assert np.array_equal(c.get_spins(), S1)
        assert np.array_equal(c.energy(), e.energy())
        c.ssf_run(L.RULE_GLAUBER, N, fluct=np.ascontiguousarray(fl[:, N:]), fluct_per_replica=True, T=np.array([1.0]), steps_per_T=N)
        assert np.array_equal(e.get_spins(), S1)            # the original did not move
        Sc = c.get_spins()
        for r in range(R):
            s, *_ = orc.ssf_run(orc.GLAUBER, J, h, S0[r], 2 * N, fluct=fl[r], T=np.array([1.0]), steps_per_T=2 * N)
            assert np.array_equal(s, Sc[r])
        e.close()                                           # the clone outlives the original
        assert np.array_equal(c.get_spins(), Sc)
    else:
        nv, nh = 48, 40
        W, hv, bh = synth.bipartite_W(nv, nh, 44, 0.3)
        prec = L.PREC_F64 if kind == "bipartite" else L.PREC_I8X3
        m = L.Model.bipartite(ctx, W, hv, bh, prec)
        Weff = m.effective_couplings()
        Sv0, Sh0 = synth.spins(45, R, nv), synth.spins(46, R, nh)
        Fv, Fh = synth.logistic(47, (4, nv), 1), synth.logistic(47, (4, nh), 2)
        Tb = np.ones(4)
        e = L.Ensemble(m, R)
        e.set_spins(Sv0)
        e.set_hidden(Sh0)
        e.bip_run(L.BIP_SCA, 2, Fv=Fv[:2], Fh=Fh[:2], T=Tb[:2])
        c = e.clone()
        S1, H1 = e.get_spins(), e.get_hidden()
        assert np.array_equal(c.get_spins(), S1) and np.array_equal(c.get_hidden(), H1)
        c.bip_run(L.BIP_SCA, 2, Fv=Fv[2:], Fh=Fh[2:], T=Tb[2:])
        assert np.array_equal(e.get_spins(), S1) and np.array_equal(e.get_hidden(), H1)
        Sc, Hc = c.get_spins(), c.get_hidden()
        for r in range(R):
            s, t, _ = orc.bip_run(orc.SCA, Weff, hv, bh, Sv0[r], Sh0[r], 4, Fv, Fh, Tb)
            assert np.array_equal(s, Sc[r]) and np.array_equal(t, Hc[r])
