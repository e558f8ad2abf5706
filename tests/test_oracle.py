"""CPU-only: pins the oracle (oracle/) against every fixture the reference's own tests hold for this path
(test/runtests.jl:19-32), against its independent pure-Python twin, and against the committed golden files."""
import numpy as np
import pytest

from cases import BIP_GOLDEN, SSF_GOLDEN, golden_J, golden_W, nodes_of
from conftest import load_golden


# ---------------------------------------------------------------- the reference's seven fixtures
def test_ref_kat_energy_two_spin(orc):
    # test/runtests.jl:20-21:  calcEnergy(SpinSystem([-1,+1], [0 1;1 0], [0,0])) == 1.0
    assert orc.energy(np.array([[0.0, 1.0], [1.0, 0.0]]), np.zeros(2), np.array([-1, 1])) == 1.0


def test_ref_kat_energy_bipartite(orc):
    # test/runtests.jl:28-29:  calcEnergy(SpinSystemOnBipartiteGraph([-1,+1], [-1,+1,-1], ones(2,3), 0, 0)) == 0.0
    assert orc.bip_energy(np.ones((2, 3)), np.zeros(2), np.zeros(3), np.array([-1, 1]), np.array([-1, 1, -1])) == 0.0


def _run_annealer_ssf(orc, rule, rng):
    # test/runtests.jl:6-17: makeSampler!(ua, 5, annealingSchedule = n -> 10.0^(-n)) then one more update!
    J, h, s = np.array([[0.0, 1.0], [1.0, 0.0]]), np.zeros(2), np.array([-1, 1], dtype=np.int8)
    draw = {0: lambda n: rng.uniform(size=n), 1: lambda n: rng.logistic(size=n), 2: lambda n: rng.exponential(size=n)}[rule]
    nodes = rng.integers(0, 2, 5).astype(np.int32)
    fl = draw(5)
    T = 10.0 ** (-np.arange(1, 6.0))
    s, _, _, _ = orc.ssf_run(rule, J, h, s, 5, nodes=nodes, fluct=fl, T=T)
    s, _, _, _ = orc.ssf_run(rule, J, h, s, 1, nodes=rng.integers(0, 2, 1).astype(np.int32), fluct=draw(1), T=T[-1:])
    return s


@pytest.mark.parametrize("rule", [0, 1, 2])
def test_ref_annealer_single_spin(orc, rule):
    # test/runtests.jl:22-24: final state in {[1,1], [-1,-1]} (seed 128 there; any stream here, 64 of them)
    for seed in range(64):
        s = _run_annealer_ssf(orc, rule, np.random.default_rng(128 + seed))
        assert s.tolist() in ([1, 1], [-1, -1])


@pytest.mark.parametrize("rule", [0, 1])
def test_ref_annealer_bipartite(orc, rule):
    # test/runtests.jl:30-31
    for seed in range(64):
        rng = np.random.default_rng(128 + seed)
        draw = (lambda sh: rng.logistic(size=sh)) if rule == 0 else (lambda sh: rng.exponential(size=sh))
        W, h, b = np.ones((2, 3)), np.zeros(2), np.zeros(3)
        s, t = np.array([-1, 1], dtype=np.int8), np.array([-1, 1, -1], dtype=np.int8)
        T = 10.0 ** (-np.arange(1, 6.0))
        Fv, Fh = draw((5, 2)), draw((5, 3))
        s, t, _ = orc.bip_run(rule, W, h, b, s, t, 5, Fv, Fh, T)
        s, t, _ = orc.bip_run(rule, W, h, b, s, t, 1, draw((1, 2)), draw((1, 3)), T[-1:])
        assert s.tolist() in ([1, 1], [-1, -1])


def test_heaviside_tie_rule(orc):
    # src/SpinSystems.jl:163-171: H(0) = c = 1
    L = orc.lib()
    assert L.orc_heaviside(0.0) == 1.0 and L.orc_heaviside(-0.0) == 1.0
    assert L.orc_heaviside(1e-300) == 1.0 and L.orc_heaviside(-1e-300) == 0.0
    # an exact tie always gives +1 whatever the current spin (Glauber at T = 0 on an isolated spin)
    for s0 in (-1, 1):
        s, *_ = orc.ssf_run(1, np.zeros((1, 1)), np.zeros(1), np.array([s0], dtype=np.int8), 1, fluct=np.zeros(1), T=np.zeros(1))
        assert s[0] == 1


def test_hopfield_minus_h_quirk(orc):
    # src/SingleSpinFlip.jl:33-34: Hopfield uses J_i.s - h_i, Glauber/Metropolis J_i.s + h_i
    J, h = np.zeros((1, 1)), np.array([0.5])
    s, *_ = orc.ssf_run(0, J, h, np.array([1], dtype=np.int8), 1)
    assert s[0] == -1
    s, *_ = orc.ssf_run(1, J, h, np.array([-1], dtype=np.int8), 1, fluct=np.zeros(1), T=np.ones(1))
    assert s[0] == 1


def test_embedding_identity(orc, synth):
    # demo.jl:82-90: H(s) = H_bip(s, s) + q N / 2 with W = (J + qI)/2, biases h/2
    J = synth.sk_J(32, 5)
    h = synth.gaussian(6, 32)
    s = synth.spins(7, 1, 32)[0]
    q = 0.5 * np.linalg.eigvalsh(J)[-1]
    Hb = orc.bip_energy(0.5 * (J + q * np.eye(32)), 0.5 * h, 0.5 * h, s, s)
    assert abs(orc.energy(J, h, s) - (Hb + 0.5 * q * 32)) < 1e-10


# ---------------------------------------------------------------- C oracle == Python twin == golden
@pytest.mark.parametrize("name", SSF_GOLDEN)
def test_golden_ssf(orc, synth, name):
    g = load_golden(name)
    J = golden_J(name, g, synth)
    s, flips, E, M = orc.ssf_run(int(g["rule"]), J, g["h"], g["s0"], int(g["nsteps"]), nodes=nodes_of(g),
                                 fluct=g["fluct"], T=g["T"], steps_per_T=int(g["steps_per_T"]),
                                 trace_every=int(g["trace_every"]))
    assert np.array_equal(s, g["s_final"]) and flips == int(g["flips"])
    assert np.array_equal(E, g["E"]) and np.array_equal(M, g["M"])


@pytest.mark.parametrize("name", BIP_GOLDEN)
def test_golden_bip(orc, synth, name):
    g = load_golden(name)
    W = golden_W(name, g, synth)
    s, t, E = orc.bip_run(int(g["rule"]), W, g["h"], g["b"], g["s0"], g["t0"], int(g["nsteps"]), g["Fv"], g["Fh"],
                          g["T"], want_E=True)
    assert np.array_equal(s, g["s_final"]) and np.array_equal(t, g["t_final"]) and np.array_equal(E, g["E"])


@pytest.mark.parametrize("rule", [0, 1, 2])
def test_c_oracle_matches_python_twin(orc, synth, rule):
    from oracle import oracle_np
    J, h = synth.sk_J(24, 11), synth.gaussian(12, 24) * 0.3
    s0 = synth.spins(13, 1, 24)[0]
    nodes = synth.nodes(14, 24, 300)
    fl = synth.logistic(15, 300) if rule == 1 else synth.exponential(15, 300)
    T = synth.geometric_schedule(3.0, 0.01, 300)
    a, fa, _, _ = orc.ssf_run(rule, J, h, s0, 300, nodes=nodes, fluct=fl, T=T)
    b, fb = oracle_np.ssf_run(rule, J, h, s0, 300, nodes=nodes, fluct=fl, T=T)
    assert np.array_equal(a, b) and fa == fb
    assert orc.energy(J, h, a) == oracle_np.energy(J, h, b)


def test_batch_equals_single(orc, synth):
    J, h = synth.sk_J(40, 1), np.zeros(40)
    S0 = synth.spins(2, 5, 40)
    fl = synth.logistic(3, (5, 200))
    T = np.full(200, 0.7)
    Sb, fb = orc.ssf_run_batch(1, J, h, S0, 200, fluct=fl, fluct_per_replica=True, T=T, nthreads=3)
    tot = 0
    for r in range(5):
        s, f, _, _ = orc.ssf_run(1, J, h, S0[r], 200, fluct=fl[r], T=T)
        assert np.array_equal(s, Sb[r])
        tot += f
    assert tot == fb


# ---------------------------------------------------------------- Philox4x32-10 known answers
def test_philox_known_answers(orc):
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    from oracle import oracle_np
    for ctr, key, want in kat:
        assert tuple(int(x) for x in orc.philox4x32_10(ctr, key)) == want
        got = oracle_np.philox4x32_10(*[np.uint32(c) for c in ctr], key[0], key[1])
        assert tuple(int(x) for x in got) == want


# ---------------------------------------------------------------- the sampled distribution (RNG-independent)
def test_kaufman_matches_enumeration():
    from exact_ising import enumerate_mean_energy, mean_energy
    for L in (3, 4):
        assert abs(enumerate_mean_energy(L, 2.269) - mean_energy(L, 2.269)) < 1e-6


@pytest.mark.parametrize("rule", [1, 2])
def test_oracle_samples_the_boltzmann_distribution(orc, synth, rule):
    """Glauber (heat bath) and Metropolis single-spin dynamics of the oracle reproduce the exact mean energy of
    the 4 x 4 torus at T = 2.269 (random-site updates as the reference draws them)."""
    from exact_ising import mean_energy
    L, T, sweeps = 4, 2.269, 60000
    N = L * L
    J = synth.lattice_J(L)
    rng = np.random.default_rng(5)
    nsteps = sweeps * N
    nodes = rng.integers(0, N, nsteps).astype(np.int32)
    fl = rng.logistic(size=nsteps) if rule == 1 else rng.exponential(size=nsteps)
    _, _, E, _ = orc.ssf_run(rule, J, np.zeros(N), synth.spins(1, 1, N)[0], nsteps, nodes=nodes, fluct=fl,
                             T=np.array([T]), steps_per_T=nsteps, trace_every=N)
    E = E[1000:]
    blocks = E[: len(E) // 50 * 50].reshape(50, -1).mean(1)
    assert abs(blocks.mean() - mean_energy(L, T)) < 5 * blocks.std() / np.sqrt(50) + 0.02
