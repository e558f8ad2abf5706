"""Stand-ins for the Distributions.jl objects stored in ``ua.distribution`` (reference:
src/SingleSpinFlip.jl:16,43,62; src/OnBipartiteGraph.jl:15,50).  They only *draw* host-side variates
for the ``rng=`` code path of SamplingHelper; the device path uses Philox (csrc/common.cuh)."""
from __future__ import annotations

import numpy as np


class _Dist:
    name = ""

    def __repr__(self):
        return f"{self.name}()"


class Uniform(_Dist):
    name = "Uniform"

    def rand(self, rng: np.random.Generator, size=None):
        return rng.uniform(0.0, 1.0, size)


class Logistic(_Dist):
    name = "Logistic"

    def rand(self, rng: np.random.Generator, size=None):
        return rng.logistic(0.0, 1.0, size)


class Exponential(_Dist):
    name = "Exponential"

    def rand(self, rng: np.random.Generator, size=None):
        return rng.exponential(1.0, size)
