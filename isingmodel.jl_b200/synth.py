"""Deterministic synthetic workloads (BASELINE.json configs) — host-side input generation only.

Everything is built from SplitMix64 with integer arithmetic and exactly-rounded float operations, so the
same seed gives bit-identical inputs on every machine (numpy's Generator streams and libm's log are not
guaranteed stable across versions / CPUs, and the golden fixtures must not depend on them)."""
from __future__ import annotations

import numpy as np

_U64 = np.uint64


def splitmix64(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n 64-bit words of SplitMix64 started at ``seed`` (stream s offsets the state by s * 2^40 steps)."""
    with np.errstate(over="ignore"):
        i = np.arange(1, n + 1, dtype=_U64) + _U64((stream << 40) & 0xFFFFFFFFFFFFFFFF)
        z = _U64(seed & 0xFFFFFFFFFFFFFFFF) + i * _U64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _U64(27))) * _U64(0x94D049BB133111EB)
        return z ^ (z >> _U64(31))


def uniform01(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """Doubles in (0, 1): (top 52 bits + 1/2) * 2^-52 — exact."""
    return ((splitmix64(seed, n, stream) >> _U64(12)).astype(np.float64) + 0.5) * (1.0 / 4503599627370496.0)


def gaussian(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """Approximately N(0, 1): Irwin-Hall sum of 12 uniforms built from 32-bit integers (exact arithmetic)."""
    acc = np.zeros(n, dtype=np.int64)
    for k in range(6):
        w = splitmix64(seed, n, stream * 8 + k + 1)
        acc += (w >> _U64(32)).astype(np.int64) + (w & _U64(0xFFFFFFFF)).astype(np.int64)
    return acc.astype(np.float64) * (1.0 / 4294967296.0) - 6.0


def spins(seed: int, R: int, N: int) -> np.ndarray:
    """(R, N) int8 spins, +1 / -1 with probability 1/2."""
    w = splitmix64(seed, R * N, 77)
    return (2 * ((w >> _U64(63)).astype(np.int8)) - 1).reshape(R, N)


def lattice_J(L: int, coupling: float = 1.0) -> np.ndarray:
    """Periodic L x L square lattice, J_ij = coupling on the 4 nearest neighbours (C1; ferromagnet for
    coupling > 0 under the reference's sign convention E = -1/2 s'Js, src/SpinSystems.jl:68-71)."""
    N = L * L
    J = np.zeros((N, N), dtype=np.float64)
    idx = np.arange(N)
    x, y = idx % L, idx // L
    for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
        j = ((x + dx) % L) + ((y + dy) % L) * L
        J[idx, j] = coupling
    if L <= 2:
        np.fill_diagonal(J, 0.0)
    return J


def sk_J(N: int, seed: int) -> np.ndarray:
    """Sherrington-Kirkpatrick couplings J_ij = J_ji ~ N(0, 1/N), zero diagonal (C2, C3)."""
    g = gaussian(seed, N * N).reshape(N, N) / np.sqrt(float(N))
    J = np.triu(g, 1)
    return J + J.T


def bipartite_W(nv: int, nh: int, seed: int, sigma: float = 0.1):
    """W ~ N(0, sigma^2) (nv x nh), h, b ~ N(0, sigma^2) (C4: sigma = 0.1)."""
    W = gaussian(seed, nv * nh).reshape(nv, nh) * sigma
    h = gaussian(seed, nv, stream=3) * sigma
    b = gaussian(seed, nh, stream=5) * sigma
    return W, h, b


def logistic(seed: int, shape, stream: int = 0) -> np.ndarray:
    """Logistic(0,1) by inversion (Distributions.jl: quantile(Logistic(), u) = log(u / (1 - u)))."""
    n = int(np.prod(shape))
    u = uniform01(seed, n, stream + 100)
    return np.log(u / (1.0 - u)).reshape(shape)


def exponential(seed: int, shape, stream: int = 0) -> np.ndarray:
    """Exponential(1) by inversion."""
    n = int(np.prod(shape))
    return (-np.log(uniform01(seed, n, stream + 200))).reshape(shape)


def nodes(seed: int, n_sites: int, nsteps: int) -> np.ndarray:
    """Uniform site indices in [0, n_sites) (0-based)."""
    return ((splitmix64(seed, nsteps, 300) >> _U64(32)) * _U64(n_sites) >> _U64(32)).astype(np.int32)


def geometric_schedule(T0: float, Tf: float, n: int) -> np.ndarray:
    """T_k = T0 (Tf/T0)^(k/(n-1)), k = 0..n-1 (C2: one temperature per sweep)."""
    if n == 1:
        return np.array([T0])
    return T0 * (Tf / T0) ** (np.arange(n) / (n - 1.0))


def checkerboard_nodes(L: int) -> np.ndarray:
    """The site list of one ISB_ORDER_CHECKERBOARD sweep of an L x L lattice (site i = x + L y, 0-based): the sites with
    (x + y) even in ascending index, then those with (x + y) odd."""
    i = np.arange(L * L)
    colour = ((i % L) + (i // L)) & 1
    return np.concatenate([i[colour == 0], i[colour == 1]]).astype(np.int32)
