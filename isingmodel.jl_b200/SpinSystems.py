"""Host-side mirror of ``SpinSystems`` (reference: src/SpinSystems.jl) over the C ABI.

Same type / function names and argument meaning as the Julia module; the arithmetic runs in
``libising_b200.so`` on the GPU (no CPU path).  Extension over the reference: the spin
configuration may be an ``(R, N)`` array, in which case the object is an ensemble of R independent
chains sharing one J; a 1-D configuration is the reference's single chain (R = 1).

Python conventions: site indices are 0-based (the Julia shim in ``julia/IsingModelB200.jl`` keeps the
reference's 1-based indices), ``update!`` is spelled ``update_``.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _lib

__all__ = ["SpinSystem", "UpdatingAlgorithm", "getSpinConfiguration", "getCouplingCoefficients",
           "getExternalMagneticField", "calcEnergy", "calcLocalMagneticField", "SpinSystemOnBipartiteGraph",
           "UpdatingAlgorithmOnBipartiteGraph", "getHiddenLayer", "getAuxiliaryBias", "calcLocalAuxiliaryBias",
           "setSpinConfiguration", "setCouplingCoefficients", "setExternalMagneticField", "setHiddenLayer",
           "setAuxiliaryBias", "heaviside"]


def _as_spins(x, name):
    a = np.asarray(x)
    if a.ndim not in (1, 2):
        raise ValueError(f"{name} must be a vector (one chain) or an (R, N) array (R replicas)")
    return a


def _checked_spins(value, shape, name):
    a = np.atleast_2d(np.asarray(value))
    if a.shape != shape:
        raise ValueError(f"{name} has the wrong shape")
    if a.dtype == np.int8:
        # the common case (and the bench's 33 MB layers): one pass over the bytes instead of a Float64 copy —
        # +1 = 0x01 and -1 = 0xFF are the only bytes x with (x + 1) & 0xFD == 0
        ok = not np.any((a.view(np.uint8) + np.uint8(1)) & np.uint8(0xFD))
    else:
        ok = bool(np.all(np.abs(np.asarray(a, dtype=np.float64)) == 1.0))
    if not ok:
        raise ValueError("spins must be +1 / -1")
    return np.ascontiguousarray(a, dtype=np.int8)


def _device_validates(value, shape, ens):
    """int8 layers of the right shape go to the library unchecked by the host mirror: isb_ens_set_spins / _set_hidden scan
    for +-1 themselves and reject BEFORE anything changes (one pass over the bytes instead of two more here: the layers
    of config 3 are 33 MB each)."""
    a = np.asarray(value)
    return ens is not None and a.dtype == np.int8 and np.atleast_2d(a).shape == shape


def _set_on_device(setter, v):
    try:
        setter(v)
    except _lib.IsbError as e:
        if e.code == _lib.ERR_ARG:
            raise ValueError("spins must be +1 / -1") from e
        raise


def _squeeze(self, arr):
    """Reference objects hold one chain: return the vector / scalar when R == 1 and the input was 1-D."""
    return arr[0] if self._single else arr


class SpinSystem:
    """src/SpinSystems.jl:14-52.  ``SpinSystem(spinConfiguration, couplingCoefficients, externalMagneticField)``."""

    def __init__(self, spinConfiguration, couplingCoefficients, externalMagneticField, *, device=None,
                 prec=_lib.PREC_AUTO):
        s = _as_spins(spinConfiguration, "spinConfiguration")
        self._single = s.ndim == 1
        s = np.atleast_2d(s)
        sparse = hasattr(couplingCoefficients, "tocsc")  # scipy.sparse, as the reference tests pass `sparse(...)`
        J = couplingCoefficients.tocsc().astype(np.float64) if sparse else np.asarray(couplingCoefficients)
        h = np.asarray(externalMagneticField)
        numNodes = s.shape[1]
        if J.ndim != 2 or J.shape[0] != J.shape[1]:
            r, c = (tuple(J.shape) + (0, 0))[:2]
            raise ValueError(f"The coupling-coefficient matrix is not a square matrix: {r}rows ≠ {c}columns.")
        row = J.shape[0]
        if numNodes < row:      # :25-27
            warnings.warn("The size of the spin-configuration vector is too smaller than the size of the "
                          "coupling-coefficient matrix.  The incorresponding components of the "
                          "coupling-coefficient matrix are ignored.")
            J = J[:numNodes, :numNodes]
        elif numNodes > row:    # :28-30
            warnings.warn("The size of the spin-configuration vector is too bigger than the size of the "
                          "coupling-coefficient matrix.  The incorresponding components of the "
                          "spin-configuration vector are ignored.")
            s = s[:, :row]
        elif (sparse and (J != J.T).nnz != 0) or (not sparse and not np.array_equal(J, J.T)):  # :31-33
            warnings.warn("The coupling-coefficient matrix should be symmetric.  It is symmetrized by its "
                          "upper-triangular components automatically.")
            if sparse:
                import scipy.sparse as sp
                J = (sp.triu(J) + sp.triu(J, 1).T).tocsc()
            else:
                J = np.triu(J) + np.triu(J, 1).T
        if np.any(J.diagonal() != 0):  # :35-38
            warnings.warn("The diagonal components of the coupling-coefficient matrix should be zero.  Their "
                          "non-zero components are ignored.")
            if sparse:
                J = J.tolil()
                J.setdiag(0)
                J = J.tocsc()
                J.eliminate_zeros()
            else:
                J = J - np.diag(np.diag(J))
        numBias = h.shape[0]
        if row != numBias:      # :40-42
            raise ValueError("The size of the coupling-coefficient matrix does not match the size of the "
                             f"external-magnetic-field vector: {row} ≠ {numBias}.")
        elif numNodes < numBias:  # :43-45
            warnings.warn("The size of the spin-configuration vector is too smaller than the size of the "
                          "external-magnetic-field vector.  The incorresponding components of the "
                          "external-magnetic-field vector are ignored.")
            h = h[:numNodes]
        self._sparse = sparse
        self._J = J if sparse else np.asarray(J, dtype=np.float64)       # float.(...) :50
        self._h = np.asarray(h, dtype=np.float64)
        self._host_spins = np.ascontiguousarray(s, dtype=np.int8)
        if not np.all(np.abs(np.asarray(s, dtype=np.float64)) == 1.0):
            raise ValueError("spins must be +1 / -1")
        self._device, self._prec = device, prec
        self._model = self._ens = None
        self._dev_newer = False
        # streaming sampler: while a recorded trajectory is being replayed (SamplingHelper.makeSampler_), the
        # state a consumer sees is a snapshot that lags the device; queries on it use a scratch ensemble
        self._snap = None       # (spins2d, energies or None)
        self._query_ens = None
        self._query_loaded = None

    # ---- device plumbing
    @property
    def replicas(self):
        return self._host_spins.shape[0]

    def _set_snapshot(self, spins2d, energies=None):
        self._snap = None if spins2d is None else (spins2d, energies)

    def _restore_snapshot(self):
        """End a replay early: the state the consumer last saw becomes the state of the host and of the device."""
        if self._snap is None:
            return
        v = np.ascontiguousarray(self._snap[0], dtype=np.int8).copy()
        self._snap = None
        self._host_spins = v
        self._dev_newer = False
        if self._ens is not None:
            self._ens.set_spins(v)

    def _query_ensemble(self):
        """The ensemble that holds the state a caller currently sees (the snapshot during a replay)."""
        ens = self._ensemble()
        if self._snap is None:
            return ens
        if self._query_ens is None:
            self._query_ens = _lib.Ensemble(self._model, self.replicas)
        if self._query_loaded is not self._snap[0]:
            self._query_ens.set_spins(self._snap[0])
            self._query_loaded = self._snap[0]
        return self._query_ens

    def _ensemble(self):
        if self._ens is None:
            ctx = _lib.context(self._device)
            if self._sparse:  # sparse couplings stay sparse on the device (csc -> isb_model_sparse)
                self._model = _lib.Model.sparse(ctx, self._J, self._h)
            else:
                self._model = _lib.Model.dense(ctx, self._J, self._h, self._prec)
            self._ens = _lib.Ensemble(self._model, self.replicas)
            self._ens.set_spins(self._host_spins)
        return self._ens

    def _invalidate_model(self):
        s = self._spins2d()
        self._host_spins = s
        self._ens = self._model = self._query_ens = self._query_loaded = None
        self._dev_newer = False

    def _spins2d(self):
        if self._snap is not None:
            return self._snap[0]
        if self._dev_newer:
            self._host_spins = self._ens.get_spins()
            self._dev_newer = False
        return self._host_spins

    # ---- reference fields
    @property
    def spinConfiguration(self):
        return _squeeze(self, self._spins2d())

    @spinConfiguration.setter
    def spinConfiguration(self, value):
        if _device_validates(value, self._host_spins.shape, self._ens):
            v = np.ascontiguousarray(np.atleast_2d(np.asarray(value)))
            _set_on_device(self._ens.set_spins, v)           # validated there BEFORE anything changes
            self._snap = None
            self._host_spins = v
            self._dev_newer = False
            return
        v = _checked_spins(value, self._host_spins.shape, "spin configuration")   # validated BEFORE anything changes
        self._snap = None       # during a replay (makeSampler_) the consumer's state wins: the sampler resumes from it
        self._host_spins = v
        self._dev_newer = False
        if self._ens is not None:
            self._ens.set_spins(v)

    @property
    def couplingCoefficients(self):
        return self._J

    @couplingCoefficients.setter
    def couplingCoefficients(self, J):
        sparse = hasattr(J, "tocsc")
        J = J.tocsc().astype(np.float64) if sparse else np.asarray(J, dtype=np.float64)
        if J.shape != self._J.shape:
            raise ValueError("coupling-coefficient matrix has the wrong shape")
        self._invalidate_model()
        self._J, self._sparse = J, sparse

    @property
    def externalMagneticField(self):
        return self._h

    @externalMagneticField.setter
    def externalMagneticField(self, h):
        h = np.asarray(h, dtype=np.float64)
        if h.shape != self._h.shape:
            raise ValueError("external-magnetic-field vector has the wrong shape")
        self._invalidate_model()
        self._h = h

    def __deepcopy__(self, memo):
        """deepcopy(ss) (test/runtests.jl:22-24): an independent system; when this one is already on the device the
        copy shares the immutable model and gets a device-side clone of the ensemble (isb_ens_clone)."""
        new = object.__new__(SpinSystem)
        new.__dict__.update(self.__dict__)
        new._J, new._h = self._J.copy(), self._h.copy()
        new._host_spins = np.array(self._spins2d(), dtype=np.int8, copy=True)
        new._snap = new._query_ens = new._query_loaded = None
        new._dev_newer = False
        if self._ens is not None:
            self._restore_or_keep()
            new._ens = self._ens.clone()
        return new

    def _restore_or_keep(self):
        if self._snap is not None:      # copied in the middle of a replay: the copy starts from what is shown
            self._restore_snapshot()


class UpdatingAlgorithm:
    """src/SpinSystems.jl:54-59 — any subclass has a ``spinSystem`` field."""
    spinSystem: SpinSystem


class SpinSystemOnBipartiteGraph:
    """src/SpinSystems.jl:90-119."""

    def __init__(self, spinConfiguration, hiddenLayer, couplingCoefficients, externalMagneticField, auxiliaryBias,
                 *, device=None, prec=_lib.PREC_F64):
        s = _as_spins(spinConfiguration, "spinConfiguration")
        t = _as_spins(hiddenLayer, "hiddenLayer")
        self._single = s.ndim == 1
        s, t = np.atleast_2d(s), np.atleast_2d(t)
        W = couplingCoefficients.toarray() if hasattr(couplingCoefficients, "toarray") else np.asarray(couplingCoefficients)
        h, b = np.asarray(externalMagneticField), np.asarray(auxiliaryBias)
        nv, nh = s.shape[1], t.shape[1]
        row, column = W.shape
        if row != nv:
            raise ValueError("The size of the coupling-coefficient matrix does not match the number of visible "
                             f"and hidden nodes: {nv}nodes ≠ {row}rows.")
        elif column != nh:
            raise ValueError("The size of the coupling-coefficient matrix does not match the number of visible "
                             f"and hidden nodes: {nh}nodes ≠ {column}columns.")
        if h.shape[0] != nv:
            raise ValueError("The size of the external-magnetic-field vector does not match the number of "
                             f"visible nodes: {h.shape[0]} ≠ {nv}.")
        elif b.shape[0] != nh:
            raise ValueError("The size of the eauxiliary-bias vector does not match the number of hidden "
                             f"nodes: {b.shape[0]} ≠ {nh}.")
        if s.shape[0] != t.shape[0]:
            raise ValueError("visible and hidden layers must hold the same number of replicas")
        self._W = np.asarray(W, dtype=np.float64)
        self._h = np.asarray(h, dtype=np.float64)
        self._b = np.asarray(b, dtype=np.float64)
        if not (np.all(np.abs(np.asarray(s, dtype=np.float64)) == 1.0)
                and np.all(np.abs(np.asarray(t, dtype=np.float64)) == 1.0)):
            raise ValueError("spins must be +1 / -1")
        self._host_s = np.ascontiguousarray(s, dtype=np.int8)
        self._host_t = np.ascontiguousarray(t, dtype=np.int8)
        self._device, self._prec = device, prec
        self._model = self._ens = None
        self._dev_newer = False
        self._snap = None       # (visible2d, hidden2d, energies or None) while a recorded trajectory is replayed
        self._query_ens = None
        self._query_loaded = None

    @property
    def replicas(self):
        return self._host_s.shape[0]

    def _set_snapshot(self, visible2d, hidden2d=None, energies=None):
        self._snap = None if visible2d is None else (visible2d, hidden2d, energies)

    def _restore_snapshot(self):
        """End a replay early: the layers the consumer last saw become the state of the host and of the device."""
        if self._snap is None:
            return
        s = np.ascontiguousarray(self._snap[0], dtype=np.int8).copy()
        t = np.ascontiguousarray(self._snap[1], dtype=np.int8).copy()
        self._snap = None
        self._host_s, self._host_t = s, t
        self._dev_newer = False
        if self._ens is not None:
            self._ens.set_spins(s)
            self._ens.set_hidden(t)

    def _query_ensemble(self):
        ens = self._ensemble()
        if self._snap is None:
            return ens
        if self._query_ens is None:
            self._query_ens = _lib.Ensemble(self._model, self.replicas)
        if self._query_loaded is not self._snap[0]:
            self._query_ens.set_spins(self._snap[0])
            self._query_ens.set_hidden(self._snap[1])
            self._query_loaded = self._snap[0]
        return self._query_ens

    def _ensemble(self):
        if self._ens is None:
            ctx = _lib.context(self._device)
            self._model = _lib.Model.bipartite(ctx, self._W, self._h, self._b, self._prec)
            self._ens = _lib.Ensemble(self._model, self.replicas)
            self._ens.set_spins(self._host_s)
            self._ens.set_hidden(self._host_t)
        return self._ens

    def _pull(self, overwriting=None):
        """Bring the host copies up to date; `overwriting` = "s" / "t": that layer is about to be replaced, skip its
        download (33 MB per layer at config 3)."""
        if self._dev_newer:
            if overwriting != "s":
                self._host_s = self._ens.get_spins()
            if overwriting != "t":
                self._host_t = self._ens.get_hidden()
            self._dev_newer = False

    def _invalidate_model(self):
        self._pull()
        self._ens = self._model = self._query_ens = self._query_loaded = None

    @property
    def spinConfiguration(self):
        if self._snap is not None:
            return _squeeze(self, self._snap[0])
        self._pull()
        return _squeeze(self, self._host_s)

    @spinConfiguration.setter
    def spinConfiguration(self, value):
        on_dev = _device_validates(value, self._host_s.shape, self._ens)
        v = np.ascontiguousarray(np.atleast_2d(np.asarray(value))) if on_dev else \
            _checked_spins(value, self._host_s.shape, "spin configuration")
        self._restore_snapshot()    # during a replay the other layer keeps the shown state, the sampler resumes from here
        self._pull(overwriting="s")
        if self._ens is not None:
            _set_on_device(self._ens.set_spins, v)           # (an int8 layer is validated there, before it changes anything)
        self._host_s = v

    @property
    def hiddenLayer(self):
        if self._snap is not None:
            return _squeeze(self, self._snap[1])
        self._pull()
        return _squeeze(self, self._host_t)

    @hiddenLayer.setter
    def hiddenLayer(self, value):
        on_dev = _device_validates(value, self._host_t.shape, self._ens)
        v = np.ascontiguousarray(np.atleast_2d(np.asarray(value))) if on_dev else \
            _checked_spins(value, self._host_t.shape, "hidden layer")
        self._restore_snapshot()
        self._pull(overwriting="t")
        if self._ens is not None:
            _set_on_device(self._ens.set_hidden, v)
        self._host_t = v

    @property
    def couplingCoefficients(self):
        return self._W

    @couplingCoefficients.setter
    def couplingCoefficients(self, W):
        W = np.asarray(W, dtype=np.float64)
        if W.shape != self._W.shape:
            raise ValueError("coupling-coefficient matrix has the wrong shape")
        self._invalidate_model()
        self._W = W

    @property
    def externalMagneticField(self):
        return self._h

    @externalMagneticField.setter
    def externalMagneticField(self, h):
        h = np.asarray(h, dtype=np.float64)
        if h.shape != self._h.shape:
            raise ValueError("external-magnetic-field vector has the wrong shape")
        self._invalidate_model()
        self._h = h

    @property
    def auxiliaryBias(self):
        return self._b

    @auxiliaryBias.setter
    def auxiliaryBias(self, b):
        b = np.asarray(b, dtype=np.float64)
        if b.shape != self._b.shape:
            raise ValueError("auxiliary-bias vector has the wrong shape")
        self._invalidate_model()
        self._b = b

    def __deepcopy__(self, memo):
        """deepcopy(ss) (test/runtests.jl:30-31): shares the immutable device model, clones the ensemble."""
        self._restore_snapshot()
        self._pull()
        new = object.__new__(SpinSystemOnBipartiteGraph)
        new.__dict__.update(self.__dict__)
        new._W, new._h, new._b = self._W.copy(), self._h.copy(), self._b.copy()
        new._host_s, new._host_t = self._host_s.copy(), self._host_t.copy()
        new._snap = new._query_ens = new._query_loaded = None
        if self._ens is not None:
            new._ens = self._ens.clone()
        return new


class UpdatingAlgorithmOnBipartiteGraph:
    """src/SpinSystems.jl:121-126."""
    spinSystem: SpinSystemOnBipartiteGraph


def _ss(x):
    return x.spinSystem if hasattr(x, "spinSystem") else x


# getters / setters: src/SpinSystems.jl:61-66, 128-137
def getSpinConfiguration(ua):
    return _ss(ua).spinConfiguration


def setSpinConfiguration(ua, spinConfiguration):
    _ss(ua).spinConfiguration = spinConfiguration


def getCouplingCoefficients(ua):
    return _ss(ua).couplingCoefficients


def setCouplingCoefficients(ua, couplingCoefficients):
    _ss(ua).couplingCoefficients = couplingCoefficients


def getExternalMagneticField(ua):
    return _ss(ua).externalMagneticField


def setExternalMagneticField(ua, externalMagneticField):
    _ss(ua).externalMagneticField = externalMagneticField


def getHiddenLayer(ua):
    return _ss(ua).hiddenLayer


def setHiddenLayer(ua, hiddenLayer):
    _ss(ua).hiddenLayer = hiddenLayer


def getAuxiliaryBias(ua):
    return _ss(ua).auxiliaryBias


def setAuxiliaryBias(ua, auxiliaryBias):
    _ss(ua).auxiliaryBias = auxiliaryBias


def calcEnergy(x):
    """src/SpinSystems.jl:68-73 / :139-145 (on the GPU: isb_ens_energy)."""
    ss = _ss(x)
    snap = getattr(ss, "_snap", None)
    if snap is not None and snap[-1] is not None:
        E = snap[-1]  # recorded by the kernel at this trace point
    elif snap is not None:
        E = ss._query_ensemble().energy()
    else:
        E = ss._ensemble().energy()
    return float(E[0]) if ss._single else E


def calcLocalMagneticField(x, nodeIndex=None):
    """src/SpinSystems.jl:75-86 / :147-152 (isb_ens_local_field); ``nodeIndex`` is 0-based."""
    ss = _ss(x)
    F = (ss._query_ensemble() if getattr(ss, "_snap", None) is not None else ss._ensemble()).local_field()
    if nodeIndex is not None:
        if isinstance(ss, SpinSystemOnBipartiteGraph):
            raise TypeError("calcLocalMagneticField(ss, i) is defined for SpinSystem only")
        F = F[:, int(nodeIndex)]
        return float(F[0]) if ss._single else F
    return F[0] if ss._single else F


def calcLocalAuxiliaryBias(x):
    """src/SpinSystems.jl:154-159 (isb_ens_local_aux_bias)."""
    ss = _ss(x)
    A = (ss._query_ensemble() if getattr(ss, "_snap", None) is not None else ss._ensemble()).local_aux_bias()
    return A[0] if ss._single else A


def heaviside(x, c=1.0):
    """src/SpinSystems.jl:163-171 — host helper for documentation / tests; kernels implement it as !(x < 0)."""
    return 1.0 if x > 0 else (0.0 if x < 0 else c)
