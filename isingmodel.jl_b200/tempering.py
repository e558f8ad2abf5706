"""Replica axis utilities: temperature ladders and replica exchange (parallel tempering) on
top of an ensemble, using per-replica temperature factors (isb_ens_set_temperature_scale).

The reference has one chain per object and a scalar ``temperature``; R reference objects that differ only in
their temperature are one ensemble here.  A swap exchanges the *temperatures* of two replicas (the spins stay
where they are), accepted with probability min(1, exp((1/T_a - 1/T_b)(E_a - E_b))).
"""
from __future__ import annotations

import numpy as np

from . import SamplingHelper, _lib
from .SpinSystems import calcEnergy


def set_temperatures(ua, temperatures):
    """Give replica r the temperature ``temperatures[r]`` for all later runs: the algorithm's scalar temperature
    (or schedule) becomes a common factor, ``temperatures / ua.temperature`` the per-replica scale."""
    ss = ua.spinSystem
    t = np.asarray(temperatures, dtype=np.float64)
    if t.shape != (ss.replicas,):
        raise ValueError("one temperature per replica is required")
    ua.temperatureScale = t / float(ua.temperature)
    ss._ensemble().set_temperature_scale(ua.temperatureScale)


def configurationIndex(spins):
    """The demo's labelling of a configuration (demo.jl:161-165): the integer whose binary digits are (1 - s_i)/2,
    first site most significant.  ``spins``: [..., N] of +-1."""
    s = np.asarray(spins)
    bits = ((1 - s) // 2).astype(np.int64)
    w = 1 << np.arange(s.shape[-1] - 1, -1, -1, dtype=np.int64)
    return bits @ w


def configurationHistogram(ua, maxMCSteps, *, stride=1, burn_in=0, seed=0, order="random", chunk=1 << 16):
    """Counts of the configurations visited by every replica (the histogram cell of demo.jl:159-168, which maps
    all states of ``makeSampler!(GlauberDynamics(ss, T), 50000)`` through ``configurationIndex``), taken every
    ``stride`` steps after ``burn_in`` steps, accumulated on the device (isb_ssf_run_hist).  Returns int64[2^N]."""
    n = ua.spinSystem._host_spins.shape[1]
    hist = np.zeros(1 << n, dtype=np.int64)
    off = 0
    if burn_in:
        SamplingHelper.run_(ua, burn_in, seed=seed, order=order, temperatures=np.array([ua.temperature]),
                            steps_per_T=burn_in)
        off = burn_in
    chunk = max(stride, chunk // stride * stride)
    done = 0
    while done < maxMCSteps:
        m = min(chunk, maxMCSteps - done)
        SamplingHelper.run_(ua, m, seed=seed, step_offset=off + done, order=order, start=(done % n) if order == "sequential" else 0,
                            temperatures=np.array([ua.temperature]), steps_per_T=m, trace_every=stride, hist=hist)
        done += m
    return hist


class ParallelTempering:
    """``ladders`` independent temperature ladders of ``len(levels)`` replicas each (replica index =
    ladder * len(levels) + slot).  ``run`` alternates ``sweeps`` sequential sweeps of every replica at its current
    temperature with one round of neighbour swaps inside each ladder."""

    def __init__(self, ua, levels, *, seed=0):
        self.ua, self.ss = ua, ua.spinSystem
        self.levels = np.asarray(levels, dtype=np.float64)
        L = len(self.levels)
        if self.ss.replicas % L:
            raise ValueError("the number of replicas must be a multiple of the number of temperature levels")
        self.nlad = self.ss.replicas // L
        # level[l, s] = index into `levels` currently held by slot s of ladder l
        self.level = np.tile(np.arange(L), (self.nlad, 1))
        self.rng = np.random.default_rng(seed)
        self.seed, self.offset = int(seed), 0
        self.accepted = np.zeros(L - 1, dtype=np.int64)
        self.proposed = np.zeros(L - 1, dtype=np.int64)
        ua.temperature = 1.0
        self._push()

    def _push(self):
        self.ua.temperatureScale = self.levels[self.level].reshape(-1)
        self.ss._ensemble().set_temperature_scale(self.ua.temperatureScale)

    def run(self, rounds, sweeps=1):
        """Returns E[rounds][levels][ladders]: the energy found at each temperature level after each round."""
        n = self.ss._host_spins.shape[1]
        L = len(self.levels)
        out = np.zeros((rounds, L, self.nlad))
        for k in range(rounds):
            SamplingHelper.run_(self.ua, sweeps * n, order="sequential", seed=self.seed, step_offset=self.offset,
                                temperatures=np.array([1.0]), steps_per_T=sweeps * n)
            self.offset += sweeps * n
            E = np.asarray(calcEnergy(self.ua)).reshape(self.nlad, L)
            # the state sampled at each level during this round = the slot that held the level BEFORE the swaps
            slot_of = np.argsort(self.level, axis=1)            # slot_of[l, level] = slot holding that level
            lad = np.arange(self.nlad)
            out[k] = E[lad[None, :], slot_of.T]
            # neighbour swaps: even pairs on even rounds, odd pairs on odd rounds
            for lo in range(k & 1, L - 1, 2):
                a, b = slot_of[:, lo], slot_of[:, lo + 1]
                d = (1.0 / self.levels[lo] - 1.0 / self.levels[lo + 1]) * (E[lad, a] - E[lad, b])
                acc = self.rng.random(self.nlad) < np.exp(np.minimum(0.0, d))
                self.proposed[lo] += self.nlad
                self.accepted[lo] += int(acc.sum())
                la = lad[acc]
                self.level[la, a[acc]], self.level[la, b[acc]] = lo + 1, lo
            self._push()
        return out
