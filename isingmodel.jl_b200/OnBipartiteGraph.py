"""Host-side mirror of ``OnBipartiteGraph`` (reference: src/OnBipartiteGraph.jl) over the C ABI."""
from __future__ import annotations

import warnings

import numpy as np

from . import _lib
from ._dist import Exponential, Logistic
from .SpinSystems import SpinSystemOnBipartiteGraph, UpdatingAlgorithmOnBipartiteGraph, calcEnergy  # noqa: F401

__all__ = ["update_", "StochasticCellularAutomata", "MomentumAnnealing", "SpinSystemOnBipartiteGraph"]


class _BipAlgorithm(UpdatingAlgorithmOnBipartiteGraph):
    _rule = -1

    def __init__(self, spinSystem: SpinSystemOnBipartiteGraph, temperature: float):
        self.spinSystem = spinSystem
        self.temperature = float(temperature)

    def __deepcopy__(self, memo):
        import copy
        new = object.__new__(type(self))
        new.__dict__ = {k: copy.deepcopy(v, memo) for k, v in self.__dict__.items()}
        return new


class StochasticCellularAutomata(_BipAlgorithm):
    """src/OnBipartiteGraph.jl:10-16 (Logistic noise: block Gibbs)."""
    _rule = _lib.BIP_SCA

    def __init__(self, spinSystem, temperature):
        super().__init__(spinSystem, temperature)
        self.distribution = Logistic()


class MomentumAnnealing(_BipAlgorithm):
    """src/OnBipartiteGraph.jl:45-51 (Exponential noise multiplied by the unit's previous value)."""
    _rule = _lib.BIP_MA

    def __init__(self, spinSystem, temperature):
        super().__init__(spinSystem, temperature)
        self.distribution = Exponential()


def update_(ua: _BipAlgorithm, fluctuationForSpinConfiguration, fluctuationForHiddenLayer):
    """``update!(ua, Fv, Fh)`` — src/OnBipartiteGraph.jl:30-43 (SCA), :53-66 (MomentumAnnealing).

    Hidden layer first (from the old visible layer), then the visible layer (from the new hidden one).
    ``Fv`` is a vector of length Nv (shared by all replicas) or an ``(R, Nv)`` array; likewise ``Fh``.
    """
    if ua.temperature < 0:  # :31-33
        warnings.warn(f"{ua.temperature} is negative.")
    ss = ua.spinSystem
    ens = ss._ensemble()
    fv = np.asarray(fluctuationForSpinConfiguration, dtype=np.float64)
    fh = np.asarray(fluctuationForHiddenLayer, dtype=np.float64)
    if fv.shape[-1] != ens.nv or fh.shape[-1] != ens.nh:
        raise ValueError("fluctuation vectors must have the sizes of the visible and the hidden layer")
    per_rep = fv.ndim == 2
    ens.bip_run(ua._rule, 1, Fv=fv, Fh=fh, fluct_per_replica=per_rep, T=[ua.temperature])
    ss._dev_newer = True
    return ss.spinConfiguration
