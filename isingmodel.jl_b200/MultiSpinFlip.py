"""``MultiSpinFlip`` — in the reference this module is an empty stub (src/MultiSpinFlip.jl:1-11: one abstract
type; SamplingHelper.jl:64-91 calls an ``update!`` that is never defined).  The only multi-spin algorithm the
reference contains is the stochastic cellular automaton on a bipartite graph, and its demo shows how a
general-graph problem is mapped onto it (demo.jl:82-90):

    W = (J + q I) / 2,   h_v = h_h = h / 2,   sigma = tau = s,     q = eigmax(J) / 2 (pinning parameter)

``StochasticCellularAutomata(SpinSystem, T)`` defined here is exactly that embedding, so it is pinned by the
bipartite oracle.  H(s) = H_bip(s, s) + q N / 2 (demo.jl:90).
"""
from __future__ import annotations

import numpy as np

from . import OnBipartiteGraph, _lib
from .SpinSystems import SpinSystem, SpinSystemOnBipartiteGraph, UpdatingAlgorithm

__all__ = ["update_", "StochasticCellularAutomata", "MultiSpinUpdatingAlgorithm"]


class MultiSpinUpdatingAlgorithm(UpdatingAlgorithm):
    """src/MultiSpinFlip.jl:9"""


class StochasticCellularAutomata(MultiSpinUpdatingAlgorithm):
    def __init__(self, spinSystem: SpinSystem, temperature: float, pinningParameter: float | None = None, *,
                 prec=_lib.PREC_F64):
        J = spinSystem.couplingCoefficients
        J = J.toarray() if hasattr(J, "toarray") else J
        n = J.shape[0]
        if pinningParameter is None:
            pinningParameter = 0.5 * float(np.linalg.eigvalsh(J)[-1])  # demo.jl:82
        self.pinningParameter = float(pinningParameter)
        self.spinSystem = spinSystem
        self._prec = prec
        self.bipartite = OnBipartiteGraph.StochasticCellularAutomata.__new__(OnBipartiteGraph.StochasticCellularAutomata)
        self.bipartite.temperature = float(temperature)
        self._embed()
        self.distribution = self.bipartite.distribution

    @property
    def temperature(self):
        return self.bipartite.temperature

    @temperature.setter
    def temperature(self, T):
        self.bipartite.temperature = float(T)

    def _embed(self):
        ss = self.spinSystem
        J = ss.couplingCoefficients
        J = J.toarray() if hasattr(J, "toarray") else J
        n = J.shape[0]
        s = np.array(ss.spinConfiguration, copy=True)
        self.bipartite = OnBipartiteGraph.StochasticCellularAutomata(
            SpinSystemOnBipartiteGraph(s, s.copy(), 0.5 * (J + self.pinningParameter * np.eye(n)),
                                       0.5 * ss.externalMagneticField, 0.5 * ss.externalMagneticField,
                                       device=ss._device, prec=self._prec), self.temperature)
        self._J_id, self._h_id = ss.couplingCoefficients, ss.externalMagneticField

    def _sync_in(self):
        """Changes made to the general-graph system since the last step (setSpinConfiguration, setCouplingCoefficients,
        setExternalMagneticField: src/SpinSystems.jl:61-66) reach the embedded bipartite system before it steps."""
        ss = self.spinSystem
        if ss.couplingCoefficients is not self._J_id or ss.externalMagneticField is not self._h_id:
            self._embed()
            return
        cur = np.atleast_2d(ss.spinConfiguration)
        b = self.bipartite.spinSystem
        if not np.array_equal(cur, np.atleast_2d(b.spinConfiguration)):
            b.spinConfiguration = cur.copy()
            b.hiddenLayer = cur.copy()

    def _sync_back(self):
        self.spinSystem.spinConfiguration = self.bipartite.spinSystem.spinConfiguration


def update_(ua: StochasticCellularAutomata, fluctuationForSpinConfiguration, fluctuationForHiddenLayer):
    """One synchronous (all-spin) SCA step of the embedded system; the visible layer is the spin configuration."""
    ua._sync_in()
    OnBipartiteGraph.update_(ua.bipartite, fluctuationForSpinConfiguration, fluctuationForHiddenLayer)
    ua._sync_back()
    return ua.spinSystem.spinConfiguration
