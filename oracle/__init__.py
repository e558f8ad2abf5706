"""CPU oracle for the IsingModel.jl spin-update hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product (``isingmodel.jl_b200``) never does.

Two restatements of the same reference lines live here:
  * ``ising_oracle.c``  — C, Float64, compiled to ``liboracle.so`` (fast enough for the CPU baseline);
  * ``oracle_np.py``    — numpy / pure-Python twin used to cross-check the C code on small cases.
Parity status is described in ``ising_oracle.h``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

HOPFIELD, GLAUBER, METROPOLIS = 0, 1, 2
SCA, MA = 0, 1


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc (idempotent)."""
    src = os.path.join(_HERE, "ising_oracle.c")
    hdr = os.path.join(_HERE, "ising_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")


def _opt(arr, dtype):
    if arr is None:
        return None
    a = np.ascontiguousarray(arr, dtype=dtype)
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = ctypes.CDLL(_LIB_PATH)
    c = ctypes
    L.orc_heaviside.restype = c.c_double
    L.orc_heaviside.argtypes = [c.c_double]
    L.orc_energy.restype = c.c_double
    L.orc_energy.argtypes = [c.c_int, c.c_void_p, c.c_int64, c.c_void_p, c.c_void_p]
    L.orc_local_field_site.restype = c.c_double
    L.orc_local_field_site.argtypes = [c.c_int, c.c_void_p, c.c_int64, c.c_void_p, c.c_void_p, c.c_int]
    L.orc_local_field.restype = None
    L.orc_local_field.argtypes = [c.c_int, c.c_void_p, c.c_int64, c.c_void_p, c.c_void_p, c.c_void_p]
    L.orc_ssf_update.restype = c.c_int
    L.orc_ssf_update.argtypes = [c.c_int, c.c_int, c.c_void_p, c.c_int64, c.c_void_p, c.c_void_p,
                                 c.c_int, c.c_double, c.c_double]
    L.orc_ssf_run.restype = c.c_int64
    L.orc_ssf_run.argtypes = [c.c_int, c.c_int, c.c_void_p, c.c_int64, c.c_void_p, c.c_void_p,
                              c.c_int64, c.c_void_p, c.c_int, c.c_void_p, c.c_void_p, c.c_int64,
                              c.c_int64, c.c_void_p, c.c_void_p]
    L.orc_ssf_run_batch.restype = c.c_int64
    L.orc_ssf_run_batch.argtypes = [c.c_int, c.c_int, c.c_void_p, c.c_int64, c.c_void_p,
                                    c.c_int, c.c_void_p, c.c_int64, c.c_int64, c.c_void_p, c.c_int,
                                    c.c_void_p, c.c_int, c.c_void_p, c.c_int64, c.c_int]
    L.orc_bip_energy.restype = c.c_double
    L.orc_bip_energy.argtypes = [c.c_int, c.c_int, c.c_void_p, c.c_int64, c.c_void_p, c.c_void_p,
                                 c.c_void_p, c.c_void_p]
    L.orc_bip_local_field.restype = None
    L.orc_bip_local_field.argtypes = [c.c_int, c.c_int, c.c_void_p, c.c_int64, c.c_void_p,
                                      c.c_void_p, c.c_void_p]
    L.orc_bip_aux_bias.restype = None
    L.orc_bip_aux_bias.argtypes = [c.c_int, c.c_int, c.c_void_p, c.c_int64, c.c_void_p,
                                   c.c_void_p, c.c_void_p]
    L.orc_bip_update.restype = None
    L.orc_bip_update.argtypes = [c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_int64, c.c_void_p,
                                 c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p,
                                 c.c_double]
    L.orc_bip_run.restype = None
    L.orc_bip_run.argtypes = [c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_int64, c.c_void_p,
                              c.c_void_p, c.c_void_p, c.c_void_p, c.c_int64, c.c_void_p,
                              c.c_void_p, c.c_void_p, c.c_int64, c.c_void_p]
    L.orc_bip_run_batch.restype = None
    L.orc_bip_run_batch.argtypes = [c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_int64, c.c_void_p,
                                    c.c_void_p, c.c_int, c.c_void_p, c.c_int64, c.c_void_p,
                                    c.c_int64, c.c_int64, c.c_void_p, c.c_void_p, c.c_int,
                                    c.c_void_p, c.c_int64, c.c_int]
    L.orc_philox4x32_10.restype = None
    L.orc_philox4x32_10.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p]
    L.orc_num_threads.restype = c.c_int
    _lib = L
    return L


# ------------------------------------------------------------------ thin numpy-facing wrappers

def _colmajor(M):
    """Return (Fortran-ordered float64 copy, leading dimension)."""
    A = np.asfortranarray(np.asarray(M, dtype=np.float64))
    return A, A.shape[0]


def energy(J, h, s):
    A, ld = _colmajor(J)
    h = _opt(h, np.float64)
    s = _opt(s, np.int8)
    return lib().orc_energy(len(s), _ptr(A), ld, _ptr(h), _ptr(s))


def local_field(J, h, s):
    A, ld = _colmajor(J)
    h = _opt(h, np.float64)
    s = _opt(s, np.int8)
    out = np.empty(len(s), dtype=np.float64)
    lib().orc_local_field(len(s), _ptr(A), ld, _ptr(h), _ptr(s), _ptr(out))
    return out


def ssf_run(rule, J, h, s, nsteps, nodes=None, start=0, fluct=None, T=None, steps_per_T=1,
            trace_every=0):
    """Run one chain in place on a copy; returns (spins, flips, E_trace, M_trace)."""
    A, ld = _colmajor(J)
    h = _opt(h, np.float64)
    s = np.array(s, dtype=np.int8, copy=True)
    n = len(s)
    nodes = _opt(nodes, np.int32)
    fluct = _opt(fluct, np.float64)
    T = _opt(T if T is not None else np.zeros(max(1, nsteps)), np.float64)
    ntr = nsteps // trace_every if trace_every > 0 else 0
    E = np.zeros(ntr, dtype=np.float64)
    M = np.zeros(ntr, dtype=np.float64)
    flips = lib().orc_ssf_run(rule, n, _ptr(A), ld, _ptr(h), _ptr(s), nsteps, _ptr(nodes), start,
                              _ptr(fluct), _ptr(T), steps_per_T, trace_every, _ptr(E), _ptr(M))
    return s, flips, E, M


def ssf_run_batch(rule, J, h, S, nsteps, nodes=None, start=0, fluct=None, fluct_per_replica=False,
                  T=None, steps_per_T=1, nthreads=1):
    """S: [R][N] int8 (replica-major). Returns (new S, total flips)."""
    A, ld = _colmajor(J)
    h = _opt(h, np.float64)
    S = np.array(S, dtype=np.int8, copy=True, order="C")
    R, n = S.shape
    nodes = _opt(nodes, np.int32)
    fluct = _opt(fluct, np.float64)
    T = _opt(T if T is not None else np.zeros(max(1, nsteps)), np.float64)
    flips = lib().orc_ssf_run_batch(rule, n, _ptr(A), ld, _ptr(h), R, _ptr(S), n, nsteps,
                                    _ptr(nodes), start, _ptr(fluct), int(bool(fluct_per_replica)),
                                    _ptr(T), steps_per_T, nthreads)
    return S, flips


def bip_energy(W, h, b, sigma, tau):
    A, ld = _colmajor(W)
    nv, nh = A.shape
    return lib().orc_bip_energy(nv, nh, _ptr(A), ld, _ptr(_opt(h, np.float64)),
                                _ptr(_opt(b, np.float64)), _ptr(_opt(sigma, np.int8)),
                                _ptr(_opt(tau, np.int8)))


def bip_local_field(W, h, tau):
    A, ld = _colmajor(W)
    nv, nh = A.shape
    out = np.empty(nv)
    lib().orc_bip_local_field(nv, nh, _ptr(A), ld, _ptr(_opt(h, np.float64)),
                              _ptr(_opt(tau, np.int8)), _ptr(out))
    return out


def bip_aux_bias(W, b, sigma):
    A, ld = _colmajor(W)
    nv, nh = A.shape
    out = np.empty(nh)
    lib().orc_bip_aux_bias(nv, nh, _ptr(A), ld, _ptr(_opt(b, np.float64)),
                           _ptr(_opt(sigma, np.int8)), _ptr(out))
    return out


def bip_run(rule, W, h, b, sigma, tau, nsteps, Fv, Fh, T, steps_per_T=1, want_E=False):
    """Fv: [nsteps][nv], Fh: [nsteps][nh] (row k = step k). Returns (sigma, tau, E)."""
    A, ld = _colmajor(W)
    nv, nh = A.shape
    sigma = np.array(sigma, dtype=np.int8, copy=True)
    tau = np.array(tau, dtype=np.int8, copy=True)
    Fv = _opt(Fv, np.float64)
    Fh = _opt(Fh, np.float64)
    T = _opt(T, np.float64)
    E = np.zeros(nsteps) if want_E else None
    lib().orc_bip_run(rule, nv, nh, _ptr(A), ld, _ptr(_opt(h, np.float64)),
                      _ptr(_opt(b, np.float64)), _ptr(sigma), _ptr(tau), nsteps, _ptr(Fv),
                      _ptr(Fh), _ptr(T), steps_per_T, _ptr(E))
    return sigma, tau, E


def bip_run_batch(rule, W, h, b, Sigma, Tau, nsteps, Fv, Fh, T, fluct_per_replica=False,
                  steps_per_T=1, nthreads=1):
    """Sigma [R][nv], Tau [R][nh]; Fv [nsteps][nv] shared or [R][nsteps][nv] per replica."""
    A, ld = _colmajor(W)
    nv, nh = A.shape
    Sigma = np.array(Sigma, dtype=np.int8, copy=True, order="C")
    Tau = np.array(Tau, dtype=np.int8, copy=True, order="C")
    R = Sigma.shape[0]
    lib().orc_bip_run_batch(rule, nv, nh, _ptr(A), ld, _ptr(_opt(h, np.float64)),
                            _ptr(_opt(b, np.float64)), R, _ptr(Sigma), nv, _ptr(Tau), nh, nsteps,
                            _ptr(_opt(Fv, np.float64)), _ptr(_opt(Fh, np.float64)),
                            int(bool(fluct_per_replica)), _ptr(_opt(T, np.float64)), steps_per_T,
                            nthreads)
    return Sigma, Tau


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_ptr(c), _ptr(k), _ptr(out))
    return out


def num_threads():
    return lib().orc_num_threads()
