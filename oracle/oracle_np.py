"""numpy / pure-Python twin of ``ising_oracle.c`` — TEST INFRASTRUCTURE ONLY.

An independent restatement of the same reference lines, written with Python floats (IEEE double,
one rounding per operation, no FMA) and explicit sequential loops, so that the C oracle can be
cross-checked on small cases.  Citations are paths under /root/reference.
"""
from __future__ import annotations

import numpy as np

HOPFIELD, GLAUBER, METROPOLIS = 0, 1, 2
SCA, MA = 0, 1


def heaviside(x: float) -> float:
    """src/SpinSystems.jl:163-171 (c = 1 at x == 0)."""
    if x > 0.0:
        return 1.0
    if x < 0.0:
        return 0.0
    return 1.0


def _rowdot(row, s) -> float:
    acc = 0.0
    for a, b in zip(row, s):
        acc += float(a) * float(b)
    return acc


def local_field_site(J, h, s, i: int) -> float:
    """src/SpinSystems.jl:80-83."""
    return _rowdot(J[i, :], s) + float(h[i])


def energy(J, h, s) -> float:
    """src/SpinSystems.jl:68-71."""
    quad = 0.0
    lin = 0.0
    for i in range(len(s)):
        quad += float(s[i]) * _rowdot(J[i, :], s)
        lin += float(h[i]) * float(s[i])
    return -0.5 * quad - lin


def ssf_update(rule: int, J, h, s, node: int, fluct: float, T: float) -> int:
    """src/SingleSpinFlip.jl:31-36, 46-55, 65-74 (in place)."""
    if rule == HOPFIELD:
        x = _rowdot(J[node, :], s) - float(h[node])
    else:
        two_h = 2.0 * local_field_site(J, h, s, node)
        ft = float(fluct) * float(T)
        if rule == METROPOLIS:
            ft = ft * float(s[node])
        x = two_h - ft
    v = int(2.0 * heaviside(x) - 1.0)
    s[node] = v
    return v


def ssf_run(rule, J, h, s, nsteps, nodes=None, start=0, fluct=None, T=None, steps_per_T=1):
    """src/SamplingHelper.jl:45-49."""
    J = np.asarray(J, dtype=np.float64)
    s = np.array(s, dtype=np.int64, copy=True)
    n = len(s)
    flips = 0
    for k in range(nsteps):
        node = int(nodes[k]) if nodes is not None else (start + k) % n
        t = float(T[k // steps_per_T]) if T is not None else 0.0
        f = float(fluct[k]) if fluct is not None else 0.0
        old = int(s[node])
        if ssf_update(rule, J, h, s, node, f, t) != old:
            flips += 1
    return s.astype(np.int8), flips


def bip_energy(W, h, b, sigma, tau) -> float:
    """src/SpinSystems.jl:139-143."""
    quad = 0.0
    lv = 0.0
    lh = 0.0
    for i in range(len(sigma)):
        quad += float(sigma[i]) * _rowdot(W[i, :], tau)
        lv += float(h[i]) * float(sigma[i])
    for j in range(len(tau)):
        lh += float(b[j]) * float(tau[j])
    return -quad - lv - lh


def bip_update(rule, W, h, b, sigma, tau, Fv, Fh, T):
    """src/OnBipartiteGraph.jl:30-43 / :53-66 (in place)."""
    nv, nh = W.shape
    aux = [_rowdot(W[:, j], sigma) + float(b[j]) for j in range(nh)]
    for j in range(nh):
        ft = float(Fh[j]) * float(T)
        if rule == MA:
            ft = ft * float(tau[j])
        tau[j] = int(2.0 * heaviside(2.0 * aux[j] - ft) - 1.0)
    fld = [_rowdot(W[i, :], tau) + float(h[i]) for i in range(nv)]
    for i in range(nv):
        ft = float(Fv[i]) * float(T)
        if rule == MA:
            ft = ft * float(sigma[i])
        sigma[i] = int(2.0 * heaviside(2.0 * fld[i] - ft) - 1.0)


def bip_run(rule, W, h, b, sigma, tau, nsteps, Fv, Fh, T, steps_per_T=1):
    """src/SamplingHelper.jl:127-131; Fv [nsteps][nv], Fh [nsteps][nh]."""
    W = np.asarray(W, dtype=np.float64)
    sigma = np.array(sigma, dtype=np.int64, copy=True)
    tau = np.array(tau, dtype=np.int64, copy=True)
    for k in range(nsteps):
        bip_update(rule, W, h, b, sigma, tau, Fv[k], Fh[k], float(T[k // steps_per_T]))
    return sigma.astype(np.int8), tau.astype(np.int8)


# ------------------------------------------------------------------ Philox4x32-10 (vectorised)

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11). All inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n1 = p1.astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            n3 = p0.astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def philox_word(seed: int, domain: int, replica, step):
    """The 32-bit word the product's counter RNG assigns to (domain, replica, step):
    counter = (lo32(step>>2), hi32(step>>2), replica, domain), key = (lo32(seed), hi32(seed)),
    word index step & 3.  See isingmodel.jl_b200/csrc/philox.cuh."""
    step = np.asarray(step, dtype=np.uint64)
    replica = np.asarray(replica, dtype=np.uint32)
    q = step >> np.uint64(2)
    w = (step & np.uint64(3)).astype(np.int64)
    out = philox4x32_10((q & np.uint64(0xFFFFFFFF)).astype(np.uint32),
                        (q >> np.uint64(32)).astype(np.uint32), replica, np.uint32(domain),
                        seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    stacked = np.stack(np.broadcast_arrays(*out), axis=0)
    return np.take_along_axis(stacked, np.broadcast_to(w, stacked.shape[1:])[None, ...], axis=0)[0]


def uniform_from_word(w):
    """u = (w + 0.5) * 2^-32, exactly representable in double, never 0 or 1."""
    return (np.asarray(w, dtype=np.float64) + 0.5) * (1.0 / 4294967296.0)
