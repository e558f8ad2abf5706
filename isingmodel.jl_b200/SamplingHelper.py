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
    if isinstance(ua, MultiSpinFlip.MultiSpinUpdatingAlgorithm):
        ua._sync_in()
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
        o = {"sequential": _lib.ORDER_SEQUENTIAL, "checkerboard": _lib.ORDER_CHECKERBOARD}.get(
            order, _lib.ORDER_LIST if nodes is not None else _lib.ORDER_RANDOM)
        out = ens.ssf_run(ua._rule, maxMCSteps, order=o, nodes=nodes, start=start, fluct=fluct,
                          fluct_per_replica=per_rep, seed=seed, step_offset=step_offset, T=Tsteps,
                          steps_per_T=steps_per_T, trace_every=trace_every, hist=hist)
        ss._dev_newer = True
        if T is not None and maxMCSteps > 0:
            ua.temperature = float(T[-1])
        return out
    if isinstance(ua, MultiSpinFlip.MultiSpinUpdatingAlgorithm):
        ua._sync_in()
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
    stride = 1, n + 1 items).  All randomness is drawn up front in the reference's order.

    Read-ahead: the device runs up to ``chunk`` steps per library call and records the state after every ``stride``
    steps (isb_ssf_run_snap / isb_bip_run_snap); the generator replays those snapshots, so
    ``map(calcEnergy, sampler)`` (demo.jl:108-115) costs no device round trip per spin.  The reference's Channel is
    unbuffered and steps lazily; the replay keeps what a consumer can observe of that:
      * every yielded state (spins, energy, fields, temperature) is the state after exactly that step;
      * if the consumer assigns spins between two items (or, with the default schedule, the temperature), the rest of
        the chunk is discarded and the run resumes from the consumer's state at the next step;
      * if the consumer stops early (``break``, ``islice``, ``close()``), the device is put back to the last yielded
        state, so later ``update_`` / ``run_`` calls continue from what the consumer saw (``chunk=stride`` gives the
        reference's strict laziness: nothing is computed ahead).
    """
    ua = updatingAlgorithm
    if maxMCSteps < 0:
        warnings.warn(f"{maxMCSteps} is negative.")
        maxMCSteps = 0
    rng = _rng(rng)
    T = _schedule(ua, maxMCSteps, annealingSchedule)
    single = isinstance(ua, SingleSpinUpdatingAlgorithm)
    multi = isinstance(ua, MultiSpinFlip.MultiSpinUpdatingAlgorithm)
    b = _bip(ua)
    if single:
        n = ua.spinSystem._host_spins.shape[1]
        updatedNodes = rng.integers(0, n, maxMCSteps)              # :39
        fluctuations = ua.distribution.rand(rng, maxMCSteps)       # :40
    else:
        if multi:
            ua._sync_in()
        ens = b.spinSystem._ensemble()
        Fv = b.distribution.rand(rng, (maxMCSteps, ens.nv))        # :121
        Fh = b.distribution.rand(rng, (maxMCSteps, ens.nh))        # :122
    per_yield = max(1, int(stride))

    def run_chunk(k, full, py):
        """Steps [k, k + full) on the device; returns the list of per-yield snapshots (tuples for ss._set_snapshot)."""
        if single:
            ss = ua.spinSystem
            out = ss._ensemble().ssf_run(ua._rule, full, nodes=updatedNodes[k:k + full],
                                         fluct=None if ua._rule == _lib.RULE_HOPFIELD else fluctuations[k:k + full],
                                         T=None if T is None else T[k + 1:k + 1 + full], trace_every=py,
                                         want_M=False, want_S=True)
            ss._dev_newer = True
            return [(out["S"][j], out["E"][j]) for j in range(full // py)]
        if multi:
            ua._sync_in()
        ss = b.spinSystem
        E, Sv, Sh = ss._ensemble().bip_run(b._rule, full, Fv=Fv[k:k + full], Fh=Fh[k:k + full], T=T[k + 1:k + 1 + full],
                                           trace_every=py, want_S=True)
        ss._dev_newer = True
        return [(Sv[j], Sh[j], E[j]) for j in range(full // py)]

    def gen():
        if T is not None:
            ua.temperature = float(T[0])                           # :43
        yield ua                                                   # :44
        ss = b.spinSystem if not single else ua.spinSystem
        k = 0
        while k < maxMCSteps:
            py = per_yield
            m = min(maxMCSteps - k, max(py, (chunk // py) * py))
            full = (m // py) * py
            if full == 0:
                full = py = m                                      # the last, shorter piece: one yield at its end
            snaps = run_chunk(k, full, py)
            done = 0
            try:
                for j, snap in enumerate(snaps):
                    ss._set_snapshot(*snap)
                    if multi:
                        ua.spinSystem._set_snapshot(snap[0])
                    kk = k + (j + 1) * py
                    if T is not None:
                        ua.temperature = float(T[kk])              # :46 / :128
                    yield ua                                       # :48 / :130
                    done = j + 1
                    touched = ss._snap is None or ss._snap[0] is not snap[0] or \
                        (multi and (ua.spinSystem._snap is None or ua.spinSystem._snap[0] is not snap[0]))
                    if T is not None and annealingSchedule is None and float(ua.temperature) != float(T[kk]):
                        T[kk:] = float(ua.temperature)             # default schedule n -> ua.temperature, read lazily
                        touched = True
                    if touched and done < len(snaps):
                        break
            finally:
                # the device holds the END of the chunk; the consumer saw snapshot `done` (1-based; at a close() or an
                # exception inside `yield` the one being shown).  Anything but a completed chunk: put the device back.
                shown = ss._snap
                complete = done == len(snaps) and shown is not None and shown[0] is snaps[-1][0]
                if shown is not None and not complete:
                    ss._restore_snapshot()
                else:
                    ss._set_snapshot(None)
                if multi:
                    assigned = ua.spinSystem._snap is None         # the consumer assigned the general-graph spins:
                    ua.spinSystem._set_snapshot(None)              # they reach the embedding at the next step
                    if not assigned:
                        ua._sync_back()
            k += done * py

    return gen()
