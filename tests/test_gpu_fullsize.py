"""GPU tests at BASELINE.json's full sizes.  The oracle cannot finish these in seconds, so parity is carried by
size-independent properties: (1) a replica's trajectory does not depend on how many other replicas run beside it
(chains are independent: the full-size run must reproduce a small run that the oracle-checked tests cover),
(2) the energies the sweep kernel maintains incrementally equal an independent recomputation from the final
spins, (3) a sample of replicas is checked against the oracle directly, (4) same seed -> same result."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lib():
    from isingmodel_jl_b200 import _lib
    return _lib


def _close(a, b, rtol=1e-9):
    return np.all(np.abs(a - b) <= rtol * np.maximum(1.0, np.abs(b)))


def test_c2_full_size_sk1024_4096_replicas(ctx, orc, synth):
    L = _lib()
    N, R, sweeps = 1024, 4096, 24
    J, h = synth.sk_J(N, 2), np.zeros(N)
    S0 = synth.spins(3, R, N)
    T = synth.geometric_schedule(2.0, 0.05, sweeps)   # spans streamed (hot) and on-demand (cold) epochs
    m = L.Model.dense(ctx, J, h, L.PREC_F64)
    e = L.Ensemble(m, R)
    e.set_spins(S0)
    out = e.ssf_run(L.RULE_GLAUBER, sweeps * N, seed=5, T=T, steps_per_T=N, trace_every=8 * N)
    S = e.get_spins()
    # (2) incremental energies == recomputation by an independent kernel
    assert _close(out["E"][-1], e.energy())
    assert np.array_equal(out["M"][-1], S.sum(1).astype(np.float64))
    assert 0 < out["flips"].sum() < sweeps * N * R
    # (1) replica independence: the first 40 replicas alone give the same trajectories
    e2 = L.Ensemble(m, 40)
    e2.set_spins(S0[:40])
    out2 = e2.ssf_run(L.RULE_GLAUBER, sweeps * N, seed=5, T=T, steps_per_T=N, trace_every=8 * N)
    assert np.array_equal(e2.get_spins(), S[:40]) and np.array_equal(out2["flips"], out["flips"][:40])
    assert np.array_equal(out2["E"], out["E"][:, :40])
    # (3) a sample of replicas against the oracle, fed the fluctuations the library dumps
    for r in (0, 1234, 4095):
        fl = ctx.philox_fluct(L.RULE_GLAUBER, 5, 0, r, 1, sweeps * N)[0]
        s, flips, Eo, _ = orc.ssf_run(1, J, h, S0[r], sweeps * N, fluct=fl, T=T, steps_per_T=N, trace_every=8 * N)
        assert np.array_equal(s, S[r]) and flips == out["flips"][r] and _close(out["E"][:, r], Eo)
    # (4) determinism
    e.set_spins(S0)
    e.ssf_run(L.RULE_GLAUBER, sweeps * N, seed=5, T=T, steps_per_T=N)
    assert np.array_equal(e.get_spins(), S)


def test_c4_full_size_784x512_16384_chains(ctx, orc, synth):
    L = _lib()
    nv, nh, R, nsteps = 784, 512, 16384, 6
    W, h, b = synth.bipartite_W(nv, nh, 4, 0.1)
    S0, T0 = synth.spins(5, R, nv), synth.spins(6, R, nh)
    T = np.ones(nsteps)
    m = L.Model.bipartite(ctx, W, h, b, L.PREC_BF16X3)
    e = L.Ensemble(m, R)                       # chain-resident persistent kernel (R >= 96 x SMs)
    e.set_spins(S0)
    e.set_hidden(T0)
    E = e.bip_run(L.BIP_SCA, nsteps, seed=9, T=T, trace_every=nsteps)
    S, H = e.get_spins(), e.get_hidden()
    assert e.last_stats()["launches"] <= 8     # one GEMM launch for all steps (+ layout conversions, energy)
    # (1) replica independence across kernel modes: 300 chains run alone (one launch per half-step)
    e2 = L.Ensemble(m, 300)
    e2.set_spins(S0[:300])
    e2.set_hidden(T0[:300])
    E2 = e2.bip_run(L.BIP_SCA, nsteps, seed=9, T=T, trace_every=nsteps)
    assert np.array_equal(e2.get_spins(), S[:300]) and np.array_equal(e2.get_hidden(), H[:300])
    assert np.array_equal(E2[0], E[0][:300])
    # (2) the traced energy is the Float64 energy of the final state
    Eo = np.array([orc.bip_energy(W, h, b, S[r], H[r]) for r in (0, 7777, 16383)])
    assert _close(E[0][[0, 7777, 16383]], Eo)
    # (3) one step from a common state vs the oracle: decisions may differ only on fp32 near ties
    e.set_spins(S0)
    e.set_hidden(T0)
    e.bip_run(L.BIP_SCA, 1, seed=9, T=T[:1])
    H1, S1 = e.get_hidden(), e.get_spins()
    Fh = ctx.philox_bip_fluct(L.BIP_SCA, 9, 0, 1, nh, 0, 4, 1)
    Fv = ctx.philox_bip_fluct(L.BIP_SCA, 9, 0, 0, nv, 0, 4, 1)
    for r in range(4):
        xh = 2.0 * orc.bip_aux_bias(W, b, S0[r]) - Fh[r, 0] * 1.0
        bad = H1[r] != np.where(xh < 0, -1, 1)
        assert np.all(np.abs(xh[bad]) < 5e-4)
        xv = 2.0 * orc.bip_local_field(W, h, H1[r]) - Fv[r, 0] * 1.0
        bad = S1[r] != np.where(xv < 0, -1, 1)
        assert np.all(np.abs(xv[bad]) < 5e-4)
    # mixing sanity: block Gibbs at T = 1 decorrelates from the initial state
    assert 0.3 < np.mean(S != S0) < 0.7


def test_c3_full_size_dense4096_8192_replicas(ctx, synth):
    L = _lib()
    N, R = 4096, 8192
    J = synth.sk_J(N, 3)
    q = 1.0                                   # eigmax(J)/2 -> 1 for J ~ N(0, 1/N); exact value irrelevant for the properties
    W = 0.5 * (J + q * np.eye(N))
    S0 = synth.spins(7, R, N)
    z = np.zeros(N)
    m = L.Model.bipartite(ctx, W, z, z, L.PREC_BF16X2)
    e = L.Ensemble(m, R)
    e.set_spins(S0)
    e.set_hidden(S0)
    E = e.bip_run(L.BIP_SCA, 2, seed=1, T=np.array([0.8, 0.4]), trace_every=2)
    S, H = e.get_spins(), e.get_hidden()
    e2 = L.Ensemble(m, 130)
    e2.set_spins(S0[:130])
    e2.set_hidden(S0[:130])
    E2 = e2.bip_run(L.BIP_SCA, 2, seed=1, T=np.array([0.8, 0.4]), trace_every=2)
    assert np.array_equal(e2.get_spins(), S[:130]) and np.array_equal(e2.get_hidden(), H[:130])
    assert np.array_equal(E2[0], E[0][:130])
    # embedding identity (demo.jl:90): H(s) = H_bip(s, s) + q N / 2 on the replicas whose two layers agree
    same = np.where((S == H).all(1))[0][:3]
    for r in same:
        Hs = -0.5 * S[r].astype(float) @ J @ S[r].astype(float)
        assert abs(Hs - (E[0][r] + 0.5 * q * N)) < 1e-6 * N


@pytest.mark.parametrize("prec", ["i8x3", "bf16x3", "fp16x2"])
def test_c3_depth_half_steps_against_the_oracle(ctx, orc, synth, prec):
    """One full SCA step (both half-steps, contraction depth K = 4096: the C3 shape) for 12 replicas against the
    Float64 oracle (src/OnBipartiteGraph.jl:35-42) with caller-supplied fluctuations.

    int8 digit planes: the contraction is exact, so the result must equal the oracle run on the grid couplings BIT FOR
    BIT; against the ORIGINAL couplings a decision may differ only where |2(W's + b) - F T| is below the deterministic
    quantisation bound K x quantum (every coupling is off by at most quantum / 2, each term enters with weight 2).
    bf16x3 / fp16x2: exact products, fp32 accumulation over K = 4096 terms: a decision may differ only in the band
    |x| <= 2 K 2^-24 max|partial sum| (each of the K additions rounds to fp32) plus the split error; the band is
    asserted and the number of differing decisions printed."""
    L = _lib()
    N, R = 4096, 12
    J = synth.sk_J(N, 3)
    q = 0.5 * 2.0                                  # ~ eigmax(J) / 2 (demo.jl:82) for J ~ N(0, 1/N)
    W = 0.5 * (J + q * np.eye(N))
    hb = synth.gaussian(31, N) * 0.05
    P = {"i8x3": L.PREC_I8X3, "bf16x3": L.PREC_BF16X3, "fp16x2": L.PREC_FP16X2}[prec]
    m = L.Model.bipartite(ctx, W, hb, hb, P)
    Weff = m.effective_couplings()
    S0 = synth.spins(32, R, N)
    Fv, Fh = synth.logistic(33, (R, 1, N), 1), synth.logistic(33, (R, 1, N), 2)
    T = 0.6
    e = L.Ensemble(m, R)
    e.set_spins(S0)
    e.set_hidden(S0)
    e.bip_run(L.BIP_SCA, 1, Fv=Fv, Fh=Fh, fluct_per_replica=True, T=np.array([T]))
    S, H = e.get_spins(), e.get_hidden()
    absW = np.abs(W).sum(1).max()
    if prec == "i8x3":
        quantum = 2.0 * np.abs(W - Weff)[~np.eye(N, dtype=bool)].max()
        assert quantum <= 2.0 ** -23 * np.abs(J).max()
        band = N * quantum                          # 2 x sum_k |dW_k| <= 2 K quantum / 2
    else:
        band = 2.0 * N * 2.0 ** -24 * absW + 2.0 * N * np.abs(W - Weff).max()
    ndiff = 0
    for r in range(R):
        if prec == "i8x3":   # exact on the grid couplings
            xh = 2.0 * orc.bip_aux_bias(Weff, hb, S0[r]) - Fh[r, 0] * T
            assert np.array_equal(H[r], np.where(xh < 0, -1, 1)), f"replica {r}: hidden layer"
            xv = 2.0 * orc.bip_local_field(Weff, hb, H[r]) - Fv[r, 0] * T
            assert np.array_equal(S[r], np.where(xv < 0, -1, 1)), f"replica {r}: visible layer"
        xh = 2.0 * orc.bip_aux_bias(W, hb, S0[r]) - Fh[r, 0] * T
        dh = H[r] != np.where(xh < 0, -1, 1)
        assert np.all(np.abs(xh[dh]) <= band), (r, np.abs(xh[dh]).max(), band)
        xv = 2.0 * orc.bip_local_field(W, hb, H[r]) - Fv[r, 0] * T
        dv = S[r] != np.where(xv < 0, -1, 1)
        assert np.all(np.abs(xv[dv]) <= band), (r, np.abs(xv[dv]).max(), band)
        ndiff += int(dh.sum() + dv.sum())
    print(f"\n[{prec}] K = {N}: {ndiff} of {2 * R * N} decisions differ from the Float64 oracle on the original "
          f"couplings (band {band:.3g})")
    assert ndiff <= 4
