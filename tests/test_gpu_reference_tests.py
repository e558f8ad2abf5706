"""The reference's own test file (test/runtests.jl:1-32) re-expressed against the Python mirror of its API,
running on the GPU through the C ABI."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def runAnnealer(pkg, updatingAlgorithm):
    # test/runtests.jl:6-17
    rng = np.random.default_rng(128)
    sampler = pkg.SamplingHelper.makeSampler_(updatingAlgorithm, 5, annealingSchedule=lambda n: 10.0 ** (-n), rng=rng)
    result = None
    count = 0
    for result in sampler:
        count += 1
    assert count == 6  # n + 1 items (SamplingHelper.jl:44,48)
    pkg.SamplingHelper.update_(updatingAlgorithm)
    return result.spinSystem.spinConfiguration.tolist()


def test_single_spin_flip(pkg, ctx):
    import scipy.sparse as sp
    SS, SSF = pkg.SpinSystems, pkg.SingleSpinFlip
    ss = SS.SpinSystem([-1, +1], sp.csc_matrix(np.array([[0, 1], [1, 0]])), [0, 0])
    assert SS.calcEnergy(ss) == 1.0
    assert runAnnealer(pkg, SSF.AsynchronousHopfieldNetwork(copy.deepcopy(ss))) in ([1, 1], [-1, -1])
    assert runAnnealer(pkg, SSF.GlauberDynamics(copy.deepcopy(ss), 10.0)) in ([1, 1], [-1, -1])
    assert runAnnealer(pkg, SSF.MetropolisMethod(copy.deepcopy(ss), 10.0)) in ([1, 1], [-1, -1])


def test_on_bipartite_graph(pkg, ctx):
    import scipy.sparse as sp
    OBG = pkg.OnBipartiteGraph
    ss = OBG.SpinSystemOnBipartiteGraph([-1, +1], [-1, +1, -1], sp.csc_matrix(np.ones((2, 3))), [0, 0], [0, 0, 0])
    assert OBG.calcEnergy(ss) == 0.0
    assert runAnnealer(pkg, OBG.StochasticCellularAutomata(copy.deepcopy(ss), 10.0)) in ([1, 1], [-1, -1])
    assert runAnnealer(pkg, OBG.MomentumAnnealing(copy.deepcopy(ss), 10.0)) in ([1, 1], [-1, -1])


def test_three_argument_update_matches_oracle(pkg, ctx, orc, synth):
    """update!(ua, node, fluct) one call at a time (the parity boundary)."""
    N = 9
    J = synth.lattice_J(3, -1.0)  # demo.jl 3x3 antiferromagnet
    s0 = synth.spins(3, 1, N)[0]
    for cls, rule in ((pkg.SingleSpinFlip.GlauberDynamics, 1), (pkg.SingleSpinFlip.MetropolisMethod, 2)):
        ua = cls(pkg.SpinSystems.SpinSystem(s0.copy(), J, np.zeros(N)), 1.3)
        nodes = synth.nodes(5, N, 40)
        fl = synth.logistic(6, 40) if rule == 1 else synth.exponential(6, 40)
        s = s0.copy()
        for k in range(40):
            pkg.SingleSpinFlip.update_(ua, int(nodes[k]), float(fl[k]))
            s, *_ = orc.ssf_run(rule, J, np.zeros(N), s, 1, nodes=nodes[k:k + 1], fluct=fl[k:k + 1], T=np.array([1.3]))
            assert np.array_equal(ua.spinSystem.spinConfiguration, s)
        assert abs(pkg.SpinSystems.calcEnergy(ua) - orc.energy(J, np.zeros(N), s)) < 1e-12
        assert pkg.SpinSystems.calcLocalMagneticField(ua, 4) == orc.local_field(J, np.zeros(N), s)[4]


def test_multispinflip_embedding(pkg, ctx, orc, synth):
    """MultiSpinFlip SCA == bipartite SCA on W = (J + qI)/2 (demo.jl:82-90)."""
    N = 48
    J, h = synth.sk_J(N, 4), synth.gaussian(5, N) * 0.1
    s0 = synth.spins(6, 1, N)[0]
    ua = pkg.MultiSpinFlip.StochasticCellularAutomata(pkg.SpinSystems.SpinSystem(s0, J, h), 0.7)
    q = ua.pinningParameter
    assert abs(q - 0.5 * np.linalg.eigvalsh(J)[-1]) < 1e-12
    Fv, Fh = synth.logistic(7, (6, N), 1), synth.logistic(7, (6, N), 2)
    s, t = s0.copy(), s0.copy()
    W = 0.5 * (J + q * np.eye(N))
    for k in range(6):
        pkg.MultiSpinFlip.update_(ua, Fv[k], Fh[k])
        s, t, _ = orc.bip_run(0, W, 0.5 * h, 0.5 * h, s, t, 1, Fv[k:k + 1], Fh[k:k + 1], np.array([0.7]))
        assert np.array_equal(ua.spinSystem.spinConfiguration, s)
    Hb = pkg.SpinSystems.calcEnergy(ua.bipartite)
    sig, tau = ua.bipartite.spinSystem.spinConfiguration, ua.bipartite.spinSystem.hiddenLayer
    assert abs(Hb - orc.bip_energy(W, 0.5 * h, 0.5 * h, sig, tau)) < 1e-9


def test_makesampler_replays_every_step(pkg, ctx, orc, synth):
    """makeSampler_ yields n + 1 states; each equals the oracle's state after that many steps, and calcEnergy on
    the yielded object (the demo's `map(calcEnergy, sampler)` idiom, demo.jl:108-115) matches at every step."""
    N, n = 9, 150
    J = synth.lattice_J(3, -1.0)
    s0 = synth.spins(4, 1, N)[0]
    sched = lambda k: 2.0 * 0.98 ** k  # noqa: E731
    for cls, rule in ((pkg.SingleSpinFlip.GlauberDynamics, 1), (pkg.SingleSpinFlip.MetropolisMethod, 2),
                      (pkg.SingleSpinFlip.AsynchronousHopfieldNetwork, 0)):
        ss = pkg.SpinSystems.SpinSystem(s0.copy(), J, np.zeros(N))
        ua = cls(ss) if rule == 0 else cls(ss, 2.0)
        rng = np.random.default_rng(3)
        states, energies, temps = [], [], []
        for item in pkg.SamplingHelper.makeSampler_(ua, n, annealingSchedule=sched, rng=rng, chunk=64):
            assert item is ua
            states.append(item.spinSystem.spinConfiguration.copy())
            energies.append(pkg.SpinSystems.calcEnergy(item))
            temps.append(getattr(item, "temperature", None))
        assert len(states) == n + 1
        # the same draws, in the reference's order (nodes first, then fluctuations)
        rng = np.random.default_rng(3)
        nodes = rng.integers(0, N, n).astype(np.int32)
        fl = ua.distribution.rand(rng, n)
        T = np.array([sched(k) for k in range(n + 1)])
        s = s0.copy()
        assert np.array_equal(states[0], s0)
        for k in range(n):
            s, *_ = orc.ssf_run(rule, J, np.zeros(N), s, 1, nodes=nodes[k:k + 1], fluct=fl[k:k + 1], T=T[k + 1:k + 2])
            assert np.array_equal(states[k + 1], s), k
            assert abs(energies[k + 1] - orc.energy(J, np.zeros(N), s)) < 1e-9
            if rule != 0:
                assert temps[k + 1] == T[k + 1]
        assert np.array_equal(ua.spinSystem.spinConfiguration, s)  # after the replay the live state is the final one
        assert abs(pkg.SpinSystems.calcLocalMagneticField(ua, 2) - orc.local_field(J, np.zeros(N), s)[2]) < 1e-12
    # strided sampler: n // stride + 1 (+1 for a remainder) items
    ss = pkg.SpinSystems.SpinSystem(s0.copy(), J, np.zeros(N))
    ua = pkg.SingleSpinFlip.GlauberDynamics(ss, 1.0)
    items = sum(1 for _ in pkg.SamplingHelper.makeSampler_(ua, 103, rng=np.random.default_rng(1), stride=10, chunk=40))
    assert items == 1 + 10 + 1


def test_makesampler_bipartite_replays_every_step(pkg, ctx, orc, synth):
    """The bipartite sampler (src/SamplingHelper.jl:110-133): n + 1 yields, both layers and the energy after every
    step equal the oracle's, with the reference's draw order (all visible fluctuations first, then all hidden)."""
    nv, nh, n = 6, 4, 40
    W, h, b = synth.bipartite_W(nv, nh, 3, 0.8)
    s0, t0 = synth.spins(4, 1, nv)[0], synth.spins(5, 1, nh)[0]
    sched = lambda k: 1.5 * 0.95 ** k  # noqa: E731
    for cls, rule in ((pkg.OnBipartiteGraph.StochasticCellularAutomata, 0), (pkg.OnBipartiteGraph.MomentumAnnealing, 1)):
        ss = pkg.OnBipartiteGraph.SpinSystemOnBipartiteGraph(s0.copy(), t0.copy(), W, h, b)
        ua = cls(ss, 1.5)
        rng = np.random.default_rng(11)
        states = [(u.spinSystem.spinConfiguration.copy(), u.spinSystem.hiddenLayer.copy(), pkg.SpinSystems.calcEnergy(u))
                  for u in pkg.SamplingHelper.makeSampler_(ua, n, annealingSchedule=sched, rng=rng, chunk=16)]
        assert len(states) == n + 1
        rng = np.random.default_rng(11)
        Fv = ua.distribution.rand(rng, (n, nv))
        Fh = ua.distribution.rand(rng, (n, nh))
        s, t = s0.copy(), t0.copy()
        for k in range(n):
            s, t, _ = orc.bip_run(rule, W, h, b, s, t, 1, Fv[k:k + 1], Fh[k:k + 1], np.array([sched(k + 1)]))
            assert np.array_equal(states[k + 1][0], s) and np.array_equal(states[k + 1][1], t), k
            assert abs(states[k + 1][2] - orc.bip_energy(W, h, b, s, t)) < 1e-9
        assert np.array_equal(ss.spinConfiguration, s) and np.array_equal(ss.hiddenLayer, t)


def test_makesampler_early_close_and_consumer_assignment(pkg, ctx, orc, synth):
    """The sampler reads ahead (one library call per chunk); what a consumer can observe must still be the reference's
    lazy, unbuffered Channel (src/SamplingHelper.jl:42-50): stopping early leaves the system in the last yielded state
    (a later update_ continues from it), and spins assigned between two items are what the next step starts from."""
    import itertools
    N, n = 9, 60
    J = synth.lattice_J(3, -1.0)
    s0 = synth.spins(4, 1, N)[0]
    sched = lambda k: 2.0 * 0.97 ** k  # noqa: E731
    rng = np.random.default_rng(5)
    nodes = rng.integers(0, N, n).astype(np.int32)
    ssd = pkg.SingleSpinFlip.GlauberDynamics(pkg.SpinSystems.SpinSystem(s0.copy(), J, np.zeros(N)), 2.0)
    fl = ssd.distribution.rand(rng, n)
    T = np.array([sched(k) for k in range(n + 1)])

    def oracle_after(k, s=None, k0=0):
        s = s0.copy() if s is None else s.copy()
        for j in range(k0, k):
            s, *_ = orc.ssf_run(1, J, np.zeros(N), s, 1, nodes=nodes[j:j + 1], fluct=fl[j:j + 1], T=T[j + 1:j + 2])
        return s

    # (a) break after 7 of 60 steps (the device had run the whole chunk ahead)
    ua = pkg.SingleSpinFlip.GlauberDynamics(pkg.SpinSystems.SpinSystem(s0.copy(), J, np.zeros(N)), 2.0)
    for item in itertools.islice(pkg.SamplingHelper.makeSampler_(ua, n, annealingSchedule=sched, rng=np.random.default_rng(5)), 8):
        pass
    s7 = oracle_after(7)
    assert np.array_equal(ua.spinSystem.spinConfiguration, s7)
    assert abs(pkg.SpinSystems.calcEnergy(ua) - orc.energy(J, np.zeros(N), s7)) < 1e-9
    assert ua.temperature == T[7]
    pkg.SingleSpinFlip.update_(ua, 4, 0.3)           # continues from the state the consumer saw, not from step 60
    s8, *_ = orc.ssf_run(1, J, np.zeros(N), s7, 1, nodes=np.array([4], dtype=np.int32), fluct=np.array([0.3]), T=T[7:8])
    assert np.array_equal(ua.spinSystem.spinConfiguration, s8)

    # (b) the consumer flips all spins after item 10: the remaining 50 steps start from its configuration
    ua = pkg.SingleSpinFlip.GlauberDynamics(pkg.SpinSystems.SpinSystem(s0.copy(), J, np.zeros(N)), 2.0)
    count = 0
    for k, item in enumerate(pkg.SamplingHelper.makeSampler_(ua, n, annealingSchedule=sched, rng=np.random.default_rng(5))):
        count += 1
        if k == 10:
            assert np.array_equal(item.spinSystem.spinConfiguration, oracle_after(10))
            pkg.SpinSystems.setSpinConfiguration(item, -item.spinSystem.spinConfiguration)
    assert count == n + 1
    assert np.array_equal(ua.spinSystem.spinConfiguration, oracle_after(n, -oracle_after(10), 10))

    # (c) a rejected assignment (not +-1) changes nothing, on the host or on the device
    before = ua.spinSystem.spinConfiguration.copy()
    with pytest.raises(ValueError):
        ua.spinSystem.spinConfiguration = np.zeros(N)
    assert np.array_equal(ua.spinSystem.spinConfiguration, before)
    assert abs(pkg.SpinSystems.calcEnergy(ua) - orc.energy(J, np.zeros(N), before)) < 1e-9


def test_multispinflip_sees_changes_of_its_spin_system(pkg, ctx, orc, synth):
    """setSpinConfiguration / setCouplingCoefficients on the general-graph system of a MultiSpinFlip algorithm
    (src/SpinSystems.jl:61-66) reach the embedded bipartite system before the next step."""
    N = 32
    J, h = synth.sk_J(N, 14), np.zeros(N)
    s0, s1 = synth.spins(15, 1, N)[0], synth.spins(16, 1, N)[0]
    ua = pkg.MultiSpinFlip.StochasticCellularAutomata(pkg.SpinSystems.SpinSystem(s0, J, h), 0.5, pinningParameter=1.0)
    Fv, Fh = synth.logistic(17, (2, N), 1), synth.logistic(17, (2, N), 2)
    pkg.SpinSystems.setSpinConfiguration(ua, s1)
    pkg.MultiSpinFlip.update_(ua, Fv[0], Fh[0])
    W = 0.5 * (J + np.eye(N))
    s, t, _ = orc.bip_run(0, W, h, h, s1, s1, 1, Fv[:1], Fh[:1], np.array([0.5]))
    assert np.array_equal(ua.spinSystem.spinConfiguration, s)
    J2 = synth.sk_J(N, 18)
    pkg.SpinSystems.setCouplingCoefficients(ua, J2)
    pkg.MultiSpinFlip.update_(ua, Fv[1], Fh[1])
    s2, _, _ = orc.bip_run(0, 0.5 * (J2 + np.eye(N)), h, h, s, s, 1, Fv[1:], Fh[1:], np.array([0.5]))
    assert np.array_equal(ua.spinSystem.spinConfiguration, s2)
