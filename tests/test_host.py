"""CPU-only: host-side logic of the Python mirror (constructor validation mirrors src/SpinSystems.jl:19-51,
97-118; schedules; synthetic generators; replica partition)."""
import warnings

import numpy as np
import pytest


def test_spinsystem_validation(pkg):
    SS = pkg.SpinSystems
    with pytest.raises(ValueError, match="not a square matrix"):
        SS.SpinSystem([1, -1], np.zeros((2, 3)), np.zeros(2))
    with pytest.raises(ValueError, match="external-magnetic-field"):
        SS.SpinSystem([1, -1], np.zeros((2, 2)), np.zeros(3))
    with pytest.warns(UserWarning, match="should be symmetric"):
        ss = SS.SpinSystem([1, -1, 1], np.array([[0, 1, 2], [5, 0, 3], [6, 7, 0.0]]), np.zeros(3))
    assert np.array_equal(ss.couplingCoefficients, np.array([[0, 1, 2], [1, 0, 3], [2, 3, 0.0]]))  # Symmetric(J, :U)
    with pytest.warns(UserWarning, match="diagonal"):
        ss = SS.SpinSystem([1, -1], np.array([[4.0, 1], [1, 9.0]]), np.zeros(2))
    assert np.array_equal(np.diag(ss.couplingCoefficients), [0, 0])
    with pytest.warns(UserWarning, match="too smaller"):
        ss = SS.SpinSystem([1, -1], np.zeros((3, 3)), np.zeros(3))
    assert ss.couplingCoefficients.shape == (2, 2)
    with pytest.raises(ValueError):
        SS.SpinSystem([1, 0], np.zeros((2, 2)), np.zeros(2))  # spins must be +-1
    ens = SS.SpinSystem(np.ones((5, 4), dtype=int), np.zeros((4, 4)), np.zeros(4))
    assert ens.replicas == 5 and ens.spinConfiguration.shape == (5, 4)


def test_bipartite_validation(pkg):
    SS = pkg.SpinSystems
    with pytest.raises(ValueError, match="rows"):
        SS.SpinSystemOnBipartiteGraph([1, -1], [1, 1, 1], np.ones((3, 3)), np.zeros(2), np.zeros(3))
    with pytest.raises(ValueError, match="columns"):
        SS.SpinSystemOnBipartiteGraph([1, -1], [1, 1, 1], np.ones((2, 2)), np.zeros(2), np.zeros(3))
    with pytest.raises(ValueError, match="visible nodes"):
        SS.SpinSystemOnBipartiteGraph([1, -1], [1, 1, 1], np.ones((2, 3)), np.zeros(3), np.zeros(3))
    with pytest.raises(ValueError, match="hidden nodes"):
        SS.SpinSystemOnBipartiteGraph([1, -1], [1, 1, 1], np.ones((2, 3)), np.zeros(2), np.zeros(2))
    ss = SS.SpinSystemOnBipartiteGraph([1, -1], [1, 1, 1], np.ones((2, 3)), np.zeros(2), np.zeros(3))
    assert ss.hiddenLayer.tolist() == [1, 1, 1]


def test_algorithm_objects(pkg):
    ss = pkg.SpinSystems.SpinSystem([1, -1], np.array([[0.0, 1], [1, 0]]), np.zeros(2))
    h = pkg.SingleSpinFlip.AsynchronousHopfieldNetwork(ss)
    g = pkg.SingleSpinFlip.GlauberDynamics(ss, 10.0)
    m = pkg.SingleSpinFlip.MetropolisMethod(ss, 10.0)
    assert not hasattr(h, "temperature") and g.temperature == 10.0
    assert (h.distribution.name, g.distribution.name, m.distribution.name) == ("Uniform", "Logistic", "Exponential")
    assert pkg.SpinSystems.getSpinConfiguration(g).tolist() == [1, -1]
    assert pkg.SpinSystems.getCouplingCoefficients(g)[0, 1] == 1.0


def test_schedule_timing(pkg):
    """T_k = schedule(k) is applied before step k; index 0 only before the first yield (SamplingHelper.jl:43-46)."""
    from isingmodel_jl_b200.SamplingHelper import _schedule
    ss = pkg.SpinSystems.SpinSystem([1, -1], np.array([[0.0, 1], [1, 0]]), np.zeros(2))
    g = pkg.SingleSpinFlip.GlauberDynamics(ss, 10.0)
    T = _schedule(g, 5, lambda n: 10.0 ** (-n))
    assert np.allclose(T, [1, .1, .01, .001, 1e-4, 1e-5])
    assert _schedule(pkg.SingleSpinFlip.AsynchronousHopfieldNetwork(ss), 5, lambda n: 1.0) is None
    assert np.array_equal(_schedule(g, 3, None), [10.0] * 4)


def test_synth_is_deterministic(synth):
    assert synth.splitmix64(0, 2).tolist() == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4]  # SplitMix64 reference values
    J = synth.sk_J(16, 3)
    assert np.array_equal(J, J.T) and not np.diag(J).any()
    L = synth.lattice_J(4)
    assert (L.sum(1) == 4).all() and np.array_equal(L, L.T)
    n = synth.nodes(1, 7, 1000)
    assert n.min() >= 0 and n.max() < 7
    s = synth.spins(1, 3, 9)
    assert set(np.unique(s)) == {-1, 1}


def test_replica_range(pkg):
    rr = pkg.sharding.replica_range
    for R, W in [(4096, 8), (10, 4), (3, 8), (16384, 2)]:
        parts = [rr(R, r, W) for r in range(W)]
        assert parts[0][0] == 0 and parts[-1][1] == R
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1


def test_sparse_input_validation(pkg):
    import scipy.sparse as sp
    SS = pkg.SpinSystems
    J = sp.csc_matrix(np.array([[0, 1, 0], [1, 0, 2.0], [0, 2.0, 0]]))
    ss = SS.SpinSystem([1, -1, 1], J, np.zeros(3))
    assert ss._sparse and ss.couplingCoefficients.nnz == 4
    with pytest.warns(UserWarning, match="should be symmetric"):
        ss = SS.SpinSystem([1, -1, 1], sp.csc_matrix(np.array([[0, 1, 2.0], [5, 0, 3], [6, 7, 0]])), np.zeros(3))
    assert np.array_equal(ss.couplingCoefficients.toarray(), np.array([[0, 1, 2], [1, 0, 3], [2, 3, 0.0]]))
    with pytest.warns(UserWarning, match="diagonal"):
        ss = SS.SpinSystem([1, -1], sp.csc_matrix(np.array([[4.0, 1], [1, 9.0]])), np.zeros(2))
    assert not ss.couplingCoefficients.diagonal().any()


def test_spin_validation_fast_path(pkg):
    """The int8 fast path of the spin check (one pass over the bytes: +1 = 0x01 and -1 = 0xFF are the only bytes x with
    (x + 1) & 0xFD == 0) accepts and rejects exactly what the general Float64 path does."""
    chk = pkg.SpinSystems._checked_spins
    good = np.array([[1, -1, -1, 1]], dtype=np.int8)
    assert chk(good, (1, 4), "spins").dtype == np.int8
    for bad in (0, 2, -2, 3, -3, 127, -128, 64, -127, 126):
        a = good.copy()
        a[0, 2] = bad
        with pytest.raises(ValueError, match=r"\+1 / -1"):
            chk(a, (1, 4), "spins")
        with pytest.raises(ValueError, match=r"\+1 / -1"):
            chk(a.astype(np.float64), (1, 4), "spins")
    assert np.array_equal(chk(np.array([1.0, -1.0]), (1, 2), "spins"), [[1, -1]])
    with pytest.raises(ValueError, match="wrong shape"):
        chk(good, (1, 5), "spins")


def test_checkerboard_site_list(synth):
    """ISB_ORDER_CHECKERBOARD's site list: a permutation of the sites, the (x + y) even ones first, ascending within a
    colour; no two sites of a colour are lattice neighbours (which is why the kernel may decide them concurrently)."""
    for L in (4, 32, 64):
        nodes = synth.checkerboard_nodes(L)
        n = L * L
        assert sorted(nodes.tolist()) == list(range(n))
        x, y = nodes % L, nodes // L
        colour = (x + y) & 1
        assert np.all(colour[: n // 2] == 0) and np.all(colour[n // 2:] == 1)
        assert np.all(np.diff(nodes[: n // 2]) > 0) and np.all(np.diff(nodes[n // 2:]) > 0)
        first = set(nodes[: n // 2].tolist())
        for i in nodes[: n // 2][:: max(1, n // 64)]:
            xi, yi = i % L, i // L
            for nb in (((xi + 1) % L) + yi * L, ((xi - 1) % L) + yi * L, xi + ((yi + 1) % L) * L, xi + ((yi - 1) % L) * L):
                assert nb not in first
