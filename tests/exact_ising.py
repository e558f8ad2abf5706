"""Exact results for the periodic L x L Ising ferromagnet (Kaufman 1949), used as an RNG-independent check of the
sampled energy distribution (north_star: 'agreement of the sampled energy distribution').  The formula was
validated against brute-force enumeration for L = 3 and 4 (see test_oracle.py::test_kaufman_matches_enumeration)."""
import numpy as np


def ln_partition(m, n, K):
    l = np.arange(2 * n)
    c = np.cosh(2 * K) / np.tanh(2 * K) - np.cos(l * np.pi / n)
    g = np.arccosh(np.maximum(c, 1.0))
    g[0] = 2 * K + np.log(np.tanh(K))
    odd, even = g[1::2], g[0::2]
    a1 = np.sum(np.log(2 * np.cosh(m * odd / 2)))
    v2 = 2 * np.sinh(m * odd / 2)
    a2, s2 = np.sum(np.log(np.abs(v2))), np.prod(np.sign(v2))
    a3 = np.sum(np.log(2 * np.cosh(m * even / 2)))
    v4 = 2 * np.sinh(m * even / 2)
    a4, s4 = np.sum(np.log(np.abs(v4))), np.prod(np.sign(v4))
    mx = max(a1, a2, a3, a4)
    tot = np.exp(a1 - mx) + s2 * np.exp(a2 - mx) + np.exp(a3 - mx) + s4 * np.exp(a4 - mx)
    return -np.log(2) + 0.5 * m * n * np.log(2 * np.sinh(2 * K)) + mx + np.log(tot)


def mean_energy(L, T, h=1e-6):
    """<E> of the L x L torus with J = 1 (E = -sum over bonds s_i s_j) at temperature T."""
    b = 1.0 / T
    return -(ln_partition(L, L, b + h) - ln_partition(L, L, b - h)) / (2 * h)


def enumerate_mean_energy(L, T):
    N = L * L
    idx = np.arange(N)
    x, y = idx % L, idx // L
    right, down = ((x + 1) % L) + y * L, x + ((y + 1) % L) * L
    states = ((np.arange(2 ** N)[:, None] >> idx[None, :]) & 1) * 2 - 1
    E = -(states * states[:, right]).sum(1) - (states * states[:, down]).sum(1)
    w = np.exp(-(E - E.min()) / T)
    return float((E * w).sum() / w.sum())
