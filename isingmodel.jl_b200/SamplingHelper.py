"""Host-side mirror of ``SamplingHelper`` (reference: src/SamplingHelper.jl) over the C ABI.

``update_`` / ``makeSampler_`` keep the reference's contract (draw order, temperature applied before the
step, n+1 yielded states that are the SAME mutable object).  ``run_`` is the batched form the library is
built for: the whole step loop of ``makeSampler!`` executes inside one C-ABI call.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import MultiSpinFlip, OnBipartiteGraph, SingleSpinFlip, _lib
from .SingleSpinFlip import SingleSpinUpdatingAlgorithm
from .SpinSystems import UpdatingAlgorithmOnBipartiteGraph

__all__ = ["update_", "makeSampler_", "run_"]


def _rng(rng):
    return rng if rng is not None else np.random.default_rng()


def _bip(ua):
    return ua.bipartite if isinstance(ua, MultiSpinFlip.MultiSpinUpdatingAlgorithm) else ua


def update_(ua, rng=None):
    """``update!(ua; rng)`` — src/SamplingHelper.jl:22-26 (single spin: node first, then the fluctuation),
    :104-108 (bipartite: visible fluctuations first, then hidden)."""
    rng = _rng(rng)
    if isinstance(ua, SingleSpinUpdatingAlgorithm):
        n = ua.spinSystem._host_spins.shape[1]
        updatedNode = int(rng.integers(0, n))
        fluctuation = float(ua.distribution.rand(rng))
        return SingleSpinFlip.update_(ua, updatedNode, fluctuation)
    b = _bip(ua)
    ss = b.spinSystem
    fv = b.distribution.rand(rng, ss._host_s.shape[1])
    fh = b.distribution.rand(rng, ss._host_t.shape[1])
    if b is ua:
        return OnBipartiteGraph.update_(ua, fv, fh)
    return MultiSpinFlip.update_(ua, fv, fh)


def _schedule(ua, maxMCSteps, annealingSchedule):
    """T[k] for k = 0..n: the reference sets T <- schedule(k) BEFORE step k (SamplingHelper.jl:43-46,128)."""
    if not hasattr(ua, "temperature"):  # Hopfield: hasproperty(ua, :temperature) is false (:33)
        return None
    if annealingSchedule is None:
        T0 = ua.temperature
        return np.full(maxMCSteps + 1, T0, dtype=np.float64)
    return np.array([float(annealingSchedule(k)) for k in range(maxMCSteps + 1)], dtype=np.float64)


def run_(ua, maxMCSteps, annealingSchedule=None, rng=None, *, seed=0, step_offset=0, order="random",
         trace_every=0, per_replica_noise=True, temperatures=None, steps_per_T=1, start=0, hist=None):
    """The whole ``makeSampler!`` loop in one library call.  Returns a dict of traces.

    rng given  -> site list and fluctuations are drawn on the host in the reference's order
                  (SamplingHelper.jl:39-40 / :121-122) and shipped to the GPU;
    rng None   -> drawn on the GPU by the counter RNG (Philox4x32-10, ``seed``/``step_offset``).
    order      -> "random" (the reference's uniformly random site per step) or "sequential" (sweeps,
                  first site ``start``).
    temperatures / steps_per_T -> the schedule already evaluated on the host: entry k // steps_per_T is the
                  temperature of step k (0-based); replaces ``annealingSchedule`` (e.g. one entry per sweep).
    hist       -> int64[2^N] (single-spin algorithms, N <= 24): the configuration of every replica at every
                  ``trace_every``-th step is counted into it on the device (see ``tempering.configurationHistogram``).
    """
    if maxMCSteps < 0:
        warnings.warn(f"{maxMCSteps} is negative.")  # SamplingHelper.jl:29-31
        maxMCSteps = 0
    if temperatures is not None and hasattr(ua, "temperature"):
        Tsteps = np.ascontiguousarray(temperatures, dtype=np.float64)
        T = Tsteps[[0, min(len(Tsteps) - 1, max(0, maxMCSteps - 1) // steps_per_T)]]
    else:
        T = _schedule(ua, maxMCSteps, annealingSchedule)
        Tsteps = None if T is None else T[1:] if maxMCSteps > 0 else T[:1]
        steps_per_T = 1
    if isinstance(ua, SingleSpinUpdatingAlgorithm):
        ss = ua.spinSystem
        ens = ss._ensemble()
        if getattr(ua, "temperatureScale", None) is not None:
            ens.set_temperature_scale(ua.temperatureScale)   # per-replica temperatures (tempering.set_temperatures)
        R, n = ss._host_spins.shape
        nodes = fluct = None
        per_rep = False
        if rng is not None:
            if order == "random":
                nodes = rng.integers(0, n, maxMCSteps)
            if ua._rule != _lib.RULE_HOPFIELD:
                per_rep = per_replica_noise and R > 1
                fluct = ua.distribution.rand(rng, (R, maxMCSteps) if per_rep else maxMCSteps)
        o = _lib.ORDER_SEQUENTIAL if order == "sequential" else (_lib.ORDER_LIST if nodes is not None else _lib.ORDER_RANDOM)
        out = ens.ssf_run(ua._rule, maxMCSteps, order=o, nodes=nodes, start=start, fluct=fluct,
                          fluct_per_replica=per_rep, seed=seed, step_offset=step_offset, T=Tsteps,
                          steps_per_T=steps_per_T, trace_every=trace_every, hist=hist)
        ss._dev_newer = True
        if T is not None and maxMCSteps > 0:
            ua.temperature = float(T[-1])
        return out
    b = _bip(ua)
    ss = b.spinSystem
    ens = ss._ensemble()
    if getattr(ua, "temperatureScale", None) is not None:
        ens.set_temperature_scale(ua.temperatureScale)
    Fv = Fh = None
    if rng is not None:
        # reference layout: (units, steps) column-major == [steps][units] row-major
        Fv = b.distribution.rand(rng, (maxMCSteps, ens.nv))
        Fh = b.distribution.rand(rng, (maxMCSteps, ens.nh))
    E = ens.bip_run(b._rule, maxMCSteps, Fv=Fv, Fh=Fh, seed=seed, step_offset=step_offset, T=Tsteps,
                    steps_per_T=steps_per_T, trace_every=trace_every)
    ss._dev_newer = True
    if maxMCSteps > 0:
        b.temperature = float(T[-1])
    if b is not ua:
        ua._sync_back()
    return {"E": E}


def makeSampler_(updatingAlgorithm, maxMCSteps, annealingSchedule=None, rng=None, *, stride=1, chunk=4096):
    """``makeSampler!(ua, n; annealingSchedule, rng)`` — src/SamplingHelper.jl:28-51, 69-91, 110-133.

    A generator standing in for the Julia ``Channel``: yields ``updatingAlgorithm`` (the same mutable
    object every time) once before the first step and then after every ``stride`` steps (reference:
    stride = 1, n + 1 items).  All randomness is drawn up front in the reference's order.  For single-spin
    algorithms the device runs `chunk` steps per library call and records the state after every `stride` steps
    (isb_ssf_run_snap); the generator replays those snapshots, so `map(calcEnergy, sampler)` (demo.jl:108-115)
    costs no device round trip per spin.
    """
    ua = updatingAlgorithm
    if maxMCSteps < 0:
        warnings.warn(f"{maxMCSteps} is negative.")
        maxMCSteps = 0
    rng = _rng(rng)
    T = _schedule(ua, maxMCSteps, annealingSchedule)
    single = isinstance(ua, SingleSpinUpdatingAlgorithm)
    if single:
        n = ua.spinSystem._host_spins.shape[1]
        updatedNodes = rng.integers(0, n, maxMCSteps)              # :39
        fluctuations = ua.distribution.rand(rng, maxMCSteps)       # :40
    else:
        b = _bip(ua)
        ens = b.spinSystem._ensemble()
        Fv = b.distribution.rand(rng, (maxMCSteps, ens.nv))        # :121
        Fh = b.distribution.rand(rng, (maxMCSteps, ens.nh))        # :122

    def gen():
        if T is not None:
            ua.temperature = float(T[0])                           # :43
        yield ua                                                   # :44
        k = 0
        while k < maxMCSteps:
            if single:
                # one library call per chunk; the kernel records the spins (and energies) after every `stride`
                # steps and the Channel contract is replayed from those snapshots
                ss = ua.spinSystem
                per_yield = max(1, stride)
                m = min(maxMCSteps - k, max(per_yield, (chunk // per_yield) * per_yield))
                full = (m // per_yield) * per_yield
                if full == 0:
                    full = m                                       # the last, shorter piece: one yield at its end
                    per_yield = m
                out = ss._ensemble().ssf_run(ua._rule, full, nodes=updatedNodes[k:k + full],
                                             fluct=None if ua._rule == _lib.RULE_HOPFIELD else fluctuations[k:k + full],
                                             T=None if T is None else T[k + 1:k + 1 + full], trace_every=per_yield,
                                             want_M=False, want_S=True)
                ss._dev_newer = True
                for j in range(full // per_yield):
                    ss._set_snapshot(out["S"][j], out["E"][j])
                    if T is not None:
                        ua.temperature = float(T[k + (j + 1) * per_yield])   # :46
                    yield ua                                       # :48
                ss._set_snapshot(None)
                k += full
                continue
            # bipartite / multi-spin: the same replay from kernel-recorded snapshots of both layers
            b = _bip(ua)
            ss = b.spinSystem
            per_yield = max(1, stride)
            m = min(maxMCSteps - k, max(per_yield, (chunk // per_yield) * per_yield))
            full = (m // per_yield) * per_yield
            if full == 0:
                full = m
                per_yield = m
            E, Sv, Sh = ss._ensemble().bip_run(b._rule, full, Fv=Fv[k:k + full], Fh=Fh[k:k + full], T=T[k + 1:k + 1 + full],
                                               trace_every=per_yield, want_S=True)
            ss._dev_newer = True
            for j in range(full // per_yield):
                ss._set_snapshot(Sv[j], Sh[j], E[j])
                if b is not ua:
                    ua.spinSystem._set_snapshot(Sv[j])
                ua.temperature = float(T[k + (j + 1) * per_yield])     # :128
                yield ua                                               # :130
            ss._set_snapshot(None)
            if b is not ua:
                ua.spinSystem._set_snapshot(None)
                ua._sync_back()
            k += full

    return gen()
