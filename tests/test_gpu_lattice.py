"""GPU parity tests of the lattice specialisation of the sparse-coupling path (csrc/lattice.cu: periodic L x L lattices,
L a multiple of 32, recognised by isb_model_sparse): the window-parallel resolution of the sequential sweep must give the
oracle's trajectory bit for bit — any bond values (the field is recomputed in the reference's summation order), every
rule, partial windows, schedules that change inside a window, traces, both noise sources."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _close(a, b):
    return np.all(np.abs(a - b) <= 1e-9 * np.maximum(1.0, np.abs(b)))


def _lib():
    from isingmodel_jl_b200 import _lib
    return _lib


def _lattice(synth, Lside, bonds, seed):
    """Periodic square lattice, sites i = x + L y; bonds: 'ferro' (+1), 'pmj' (+-1), 'gauss' (non-dyadic Gaussian)."""
    import scipy.sparse as sp
    n = Lside * Lside
    idx = np.arange(n)
    x, y = idx % Lside, idx // Lside
    right = (x + 1) % Lside + y * Lside
    down = x + ((y + 1) % Lside) * Lside
    g = synth.gaussian(seed, 2 * n)
    if bonds == "ferro":
        v = np.ones(2 * n)
    elif bonds == "pmj":
        v = np.where(g > 0, 1.0, -1.0)
    else:
        v = g * 0.7 + 0.1
    A = sp.coo_matrix((np.concatenate([v[:n], v[n:]]), (np.concatenate([idx, idx]), np.concatenate([right, down]))), shape=(n, n))
    A = (A + A.T).tocsc()
    return A


CASES = [  # L, bonds, rule, R, nsteps, start, per-replica noise
    (32, "ferro", 2, 40, 32 * 32 * 3, 0, True), (32, "gauss", 1, 33, 32 * 32 * 2 + 77, 5, True),
    (32, "pmj", 0, 9, 32 * 32 * 2, 1000, False), (64, "gauss", 2, 20, 64 * 64 * 2 + 13, 64 * 7 + 31, True),
    (64, "pmj", 1, 6, 64 * 64 + 5, 0, False), (96, "gauss", 1, 5, 96 * 96 + 100, 96 * 95, True),
    (32, "gauss", 2, 300, 2048, 17, True),
]


@pytest.mark.parametrize("Lside,bonds,rule,R,nsteps,start,per_rep", CASES)
def test_lattice_bit_exact(ctx, orc, synth, monkeypatch, Lside, bonds, rule, R, nsteps, start, per_rep):
    L = _lib()
    n = Lside * Lside
    A = _lattice(synth, Lside, bonds, 50 + Lside)
    J = A.toarray()
    h = synth.gaussian(51, n) * (0.0 if bonds == "ferro" else 0.3)
    S0 = synth.spins(52, R, n)
    gen = synth.logistic if rule == 1 else synth.exponential
    fl = None if rule == 0 else gen(53, (R, nsteps) if per_rep else nsteps)
    T = synth.geometric_schedule(3.0, 0.3, 7)
    spT = (nsteps + 6) // 7                                   # the schedule changes in the middle of windows
    tr = max(1, nsteps // 3)
    m = L.Model.sparse(ctx, A, h)
    e = L.Ensemble(m, R)
    e.set_spins(S0)
    e.set_tie_eps(1e-12)
    out = e.ssf_run(rule, nsteps, start=start, fluct=fl, fluct_per_replica=per_rep, T=T, steps_per_T=spT, trace_every=tr,
                    want_S=True)
    S = e.get_spins()
    assert e.last_stats()["launches"] == 1                    # the lattice kernel alone (no field refresh launch)
    for r in list(range(min(R, 10))) + [R - 1]:
        s, flips, E, M = orc.ssf_run(rule, J, h, S0[r], nsteps, start=start, fluct=None if fl is None else (fl[r] if per_rep else fl),
                                     T=T, steps_per_T=spT, trace_every=tr)
        assert np.array_equal(s, S[r]), f"replica {r}"
        assert flips == out["flips"][r]
        assert np.array_equal(M, out["M"][:, r]) and _close(out["E"][:, r], E)
        assert np.array_equal(out["S"][-1][r], S[r]) if nsteps % tr == 0 else True
    assert _close(e.energy()[:3], np.array([orc.energy(J, h, S[r]) for r in range(3)]))
    # the generic neighbour-list kernel on the same model gives the same bits
    monkeypatch.setenv("ISB_LATTICE", "0")
    e2 = L.Ensemble(L.Model.sparse(ctx, A, h), R)
    e2.set_spins(S0)
    out2 = e2.ssf_run(rule, nsteps, start=start, fluct=fl, fluct_per_replica=per_rep, T=T, steps_per_T=spT, trace_every=tr)
    if bonds != "gauss":      # (non-dyadic bonds: the cached incremental field of the generic kernel may differ by ulps)
        assert np.array_equal(e2.get_spins(), S) and np.array_equal(out2["flips"], out["flips"])


def test_lattice_philox_equals_generic_kernel_and_dumped_noise(ctx, orc, synth, monkeypatch):
    """In-kernel noise: the lattice kernel hands out the Philox words of 128 steps with shuffles; the trajectory equals
    the neighbour-list kernel's (same words, same transforms) and the oracle's on the dumped fluctuations — also when the
    run is cut into pieces with odd step offsets."""
    import scipy.sparse as sp
    L = _lib()
    n, R = 1024, 70
    A = sp.csc_matrix(synth.lattice_J(32))
    S0 = synth.spins(61, R, n)
    T = np.array([2.269])
    e = L.Ensemble(L.Model.sparse(ctx, A, np.zeros(n)), R)
    e.set_spins(S0)
    pieces = [(0, 1000), (1000, 3333), (4333, 5 * n - 4333)]
    for off, cnt in pieces:
        e.ssf_run(L.RULE_METROPOLIS, cnt, start=off % n, seed=77, step_offset=off, T=T, steps_per_T=cnt)
    S = e.get_spins()
    monkeypatch.setenv("ISB_LATTICE", "0")
    e2 = L.Ensemble(L.Model.sparse(ctx, A, np.zeros(n)), R)
    e2.set_spins(S0)
    e2.ssf_run(L.RULE_METROPOLIS, 5 * n, seed=77, T=T, steps_per_T=5 * n)
    assert np.array_equal(e2.get_spins(), S)
    fl = ctx.philox_fluct(L.RULE_METROPOLIS, 77, 0, 0, 3, 5 * n)
    J = synth.lattice_J(32)
    for r in range(3):
        s, *_ = orc.ssf_run(L.RULE_METROPOLIS, J, np.zeros(n), S0[r], 5 * n, fluct=fl[r], T=T, steps_per_T=5 * n)
        assert np.array_equal(s, S[r])


def test_non_lattice_graphs_are_not_misdetected(ctx, orc, synth):
    """1024 sites with 4 neighbours each that are NOT the square lattice's (one bond moved): the generic kernel runs."""
    import scipy.sparse as sp
    L = _lib()
    J = synth.lattice_J(32)
    J[0, 1] = J[1, 0] = 0.0
    J[0, 2] = J[2, 0] = 1.0
    n, R, nsteps = 1024, 5, 2048
    S0 = synth.spins(71, R, n)
    fl = synth.exponential(72, (R, nsteps))
    e = L.Ensemble(L.Model.sparse(ctx, sp.csc_matrix(J), np.zeros(n)), R)
    e.set_spins(S0)
    e.ssf_run(2, nsteps, fluct=fl, fluct_per_replica=True, T=np.array([2.0]), steps_per_T=nsteps)
    assert e.last_stats()["launches"] == 2                    # field refresh + the neighbour-list kernel
    S = e.get_spins()
    for r in range(R):
        s, *_ = orc.ssf_run(2, J, np.zeros(n), S0[r], nsteps, fluct=fl[r], T=np.array([2.0]), steps_per_T=nsteps)
        assert np.array_equal(s, S[r])


CB_CASES = [  # L, bonds, rule, R, nsteps, start (position in the two-colour sweep order), per-replica noise
    (32, "gauss", 2, 12, 32 * 32 * 2 + 77, 0, True), (32, "pmj", 1, 7, 32 * 32 * 3, 523, False),
    (32, "ferro", 2, 33, 1500, 500, True), (64, "gauss", 1, 6, 64 * 64 * 2 + 13, 64 * 32 + 17, True),
    (64, "gauss", 0, 4, 64 * 64 + 5, 3, False), (96, "pmj", 2, 3, 96 * 96 + 100, 96 * 48 - 5, True),
]


@pytest.mark.parametrize("Lside,bonds,rule,R,nsteps,start,per_rep", CB_CASES)
def test_checkerboard_order_equals_its_site_list(ctx, orc, synth, Lside, bonds, rule, R, nsteps, start, per_rep):
    """ISB_ORDER_CHECKERBOARD is ONE site list of the reference's 3-argument update! (all sites of one colour in
    ascending index, then the other colour); sites of a colour do not interact, so the kernel decides 32 of them at
    once — and must give, bit for bit, what the oracle gives when it walks that list one site at a time
    (src/SingleSpinFlip.jl:31-36,46-55,65-74), and what the library's own list order gives on the same model."""
    L = _lib()
    n = Lside * Lside
    A = _lattice(synth, Lside, bonds, 60 + Lside)
    J = A.toarray()
    h = synth.gaussian(61, n) * (0.0 if bonds == "ferro" else 0.3)
    S0 = synth.spins(62, R, n)
    gen = synth.logistic if rule == 1 else synth.exponential
    fl = None if rule == 0 else gen(63, (R, nsteps) if per_rep else nsteps)
    T = synth.geometric_schedule(3.0, 0.3, 5)
    spT = (nsteps + 4) // 5                                   # the schedule changes in the middle of windows
    tr = max(1, nsteps // 3)
    sweep = synth.checkerboard_nodes(Lside)
    nodes = np.resize(np.roll(sweep, -start), nsteps)         # the list from position `start` on, sweep after sweep
    e = L.Ensemble(L.Model.sparse(ctx, A, h), R)
    e.set_spins(S0)
    out = e.ssf_run(rule, nsteps, order=L.ORDER_CHECKERBOARD, start=start, fluct=fl, fluct_per_replica=per_rep, T=T,
                    steps_per_T=spT, trace_every=tr, want_S=True)
    S = e.get_spins()
    assert e.last_stats()["launches"] == 1
    for r in list(range(min(R, 6))) + [R - 1]:
        s, flips, E, M = orc.ssf_run(rule, J, h, S0[r], nsteps, nodes=nodes, fluct=None if fl is None else (fl[r] if per_rep else fl),
                                     T=T, steps_per_T=spT, trace_every=tr)
        assert np.array_equal(s, S[r]), f"replica {r}"
        assert flips == out["flips"][r]
        assert np.array_equal(M, out["M"][:, r]) and _close(out["E"][:, r], E)
    e2 = L.Ensemble(L.Model.sparse(ctx, A, h), R)              # the list order of the library (neighbour-list kernel)
    e2.set_spins(S0)
    out2 = e2.ssf_run(rule, nsteps, nodes=nodes, fluct=fl, fluct_per_replica=per_rep, T=T, steps_per_T=spT, trace_every=tr)
    assert np.array_equal(e2.get_spins(), S) and np.array_equal(out2["flips"], out["flips"])


def test_checkerboard_order_needs_a_lattice(ctx, synth):
    L = _lib()
    n = 64
    J = synth.sk_J(n, 3)
    e = L.Ensemble(L.Model.dense(ctx, J, np.zeros(n)), 2)
    e.set_spins(synth.spins(1, 2, n))
    with pytest.raises(L.IsbError) as ei:
        e.ssf_run(L.RULE_GLAUBER, 10, order=L.ORDER_CHECKERBOARD, T=np.ones(1), steps_per_T=10)
    assert ei.value.code == L.ERR_UNSUPPORTED


def test_checkerboard_order_samples_the_exact_mean_energy(ctx, synth):
    """Two-colour sweeps of the 32 x 32 ferromagnet at T = 2.269 (Metropolis, in-kernel noise) give Kaufman's exact
    finite-lattice mean energy: the order changes the trajectory, not the stationary distribution (tolerance: 5 standard
    errors of the replica scatter + 0.2 % for the residual equilibration bias at the critical point)."""
    from exact_ising import mean_energy
    L = _lib()
    Ls, T, R = 32, 2.269, 2048
    n = Ls * Ls
    A = _lattice(synth, Ls, "ferro", 1)
    e = L.Ensemble(L.Model.sparse(ctx, A, np.zeros(n)), R)
    e.set_spins(synth.spins(5, R, n))
    e.ssf_run(L.RULE_METROPOLIS, 2000 * n, order=L.ORDER_CHECKERBOARD, seed=11, T=np.array([T]), steps_per_T=2000 * n)
    out = e.ssf_run(L.RULE_METROPOLIS, 500 * n, order=L.ORDER_CHECKERBOARD, seed=11, step_offset=2000 * n, T=np.array([T]),
                    steps_per_T=500 * n, trace_every=10 * n)
    Em = out["E"].mean(0)
    exact = mean_energy(Ls, T)
    err = Em.std() / np.sqrt(R)
    assert abs(Em.mean() - exact) < 5 * err + 2e-3 * abs(exact), (Em.mean(), exact, err)
