"""Host-side mirror of ``SingleSpinFlip`` (reference: src/SingleSpinFlip.jl) over the C ABI."""
from __future__ import annotations

import warnings

import numpy as np

from . import _lib
from ._dist import Exponential, Logistic, Uniform
from .SpinSystems import SpinSystem, UpdatingAlgorithm

__all__ = ["update_", "AsynchronousHopfieldNetwork", "GlauberDynamics", "MetropolisMethod"]


class SingleSpinUpdatingAlgorithm(UpdatingAlgorithm):
    """src/SingleSpinFlip.jl:10"""
    _rule = -1

    def __deepcopy__(self, memo):
        import copy
        new = object.__new__(type(self))
        new.__dict__ = {k: copy.deepcopy(v, memo) for k, v in self.__dict__.items()}
        return new


class AsynchronousHopfieldNetwork(SingleSpinUpdatingAlgorithm):
    """src/SingleSpinFlip.jl:12-17 (no temperature field; ``distribution`` is a dummy Uniform)."""
    _rule = _lib.RULE_HOPFIELD

    def __init__(self, spinSystem: SpinSystem):
        self.spinSystem = spinSystem
        self.distribution = Uniform()


class GlauberDynamics(SingleSpinUpdatingAlgorithm):
    """src/SingleSpinFlip.jl:38-44 (Logistic noise == heat bath)."""
    _rule = _lib.RULE_GLAUBER

    def __init__(self, spinSystem: SpinSystem, temperature: float):
        self.spinSystem = spinSystem
        self.temperature = float(temperature)
        self.distribution = Logistic()


class MetropolisMethod(SingleSpinUpdatingAlgorithm):
    """src/SingleSpinFlip.jl:57-63 (Exponential noise)."""
    _rule = _lib.RULE_METROPOLIS

    def __init__(self, spinSystem: SpinSystem, temperature: float):
        self.spinSystem = spinSystem
        self.temperature = float(temperature)
        self.distribution = Exponential()


def _warn_negative(ua):
    T = getattr(ua, "temperature", 0.0)
    if T < 0:  # src/SingleSpinFlip.jl:47-49, 66-68
        warnings.warn(f"{T} is negative.")


def update_(ua: SingleSpinUpdatingAlgorithm, updatedNode: int, fluctuation=0.0):
    """``update!(ua, updatedNode, fluctuation)`` — src/SingleSpinFlip.jl:31-36, 46-55, 65-74.

    One single-spin update of every replica at site ``updatedNode`` (0-based).  ``fluctuation`` is a
    scalar (shared by all replicas) or an array with one entry per replica.
    """
    _warn_negative(ua)
    ss = ua.spinSystem
    ens = ss._ensemble()
    f = np.atleast_1d(np.asarray(fluctuation, dtype=np.float64))
    per_rep = f.size != 1
    ens.ssf_run(ua._rule, 1, nodes=[int(updatedNode)], fluct=f, fluct_per_replica=per_rep,
                T=[getattr(ua, "temperature", 0.0)])
    ss._dev_newer = True
    s = ss._ens.get_spins()
    ss._host_spins, ss._dev_newer = s, False
    v = s[:, int(updatedNode)]
    return int(v[0]) if ss._single else v
