"""ctypes wrapper of oracle/mile_oracle.c (the C restatement; TEST INFRASTRUCTURE / CPU baseline only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / 'libmile_oracle.so'
_ACT = {'identity': 0, 'relu': 1, 'sigmoid': 2, 'tanh': 3, 'gelu': 4, 'leaky_relu': 5}


class Spec(C.Structure):
    _fields_ = [('n_features', C.c_int32), ('n_layers', C.c_int32), ('widths', C.c_int32 * 12),
                ('bias_off', C.c_int32 * 12), ('kernel_off', C.c_int32 * 12), ('activation', C.c_int32),
                ('task', C.c_int32), ('prior', C.c_int32), ('prior_loc', C.c_float), ('prior_scale', C.c_float),
                ('n_batches', C.c_float)]


_lib = None


def build():
    subprocess.run(['make', '-s', '-C', str(HERE)], check=True)


def available() -> bool:
    try:
        load()
        return True
    except Exception:
        return False


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB.exists() or LIB.stat().st_mtime < (HERE / 'mile_oracle.c').stat().st_mtime:
        build()
    lib = C.CDLL(str(LIB))
    vp, i32, i64, u64, f = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float
    sp = C.POINTER(Spec)
    lib.mo_logpost_value_and_grad.argtypes = [sp, vp, vp, vp, i64, vp]
    lib.mo_logpost_value_and_grad.restype = f
    lib.mo_logpost_batch.argtypes = [sp, vp, i32, vp, vp, i64, vp, vp, i32]
    lib.mo_init.argtypes = [sp, vp, vp, i64, i32, vp, vp, vp, vp, vp, u64, i32]
    lib.mo_run_sampling.argtypes = [sp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, i32, i64, i32, vp, u64, vp, vp, i32]
    lib.mo_philox_normal.argtypes = [u64, C.c_uint32, u64, C.c_uint32, C.c_uint32]
    lib.mo_philox_normal.restype = f
    lib.mo_max_threads.restype = C.c_int
    lib.mo_set_row_threads.argtypes = [C.c_int]
    _lib = lib
    return lib


def _spec(ospec) -> Spec:
    s = Spec()
    s.n_features, s.n_layers = ospec.n_features, ospec.n_layers
    b, k = ospec.offsets()
    for l, w in enumerate(ospec.widths):
        s.widths[l], s.bias_off[l], s.kernel_off[l] = w, b[l], k[l]
    s.activation = _ACT[ospec.activation]
    s.task = 0 if ospec.task == 'regr' else 1
    s.prior = 0 if ospec.prior == 'normal' else 1
    s.prior_loc, s.prior_scale, s.n_batches = ospec.prior_loc, ospec.prior_scale, ospec.n_batches
    return s


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _xy(ospec, X, y):
    X = np.ascontiguousarray(X, np.float32)
    y = np.ascontiguousarray(y, np.int32 if ospec.task == 'class' else np.float32)
    return X, y


def max_threads() -> int:
    return int(load().mo_max_threads())


def logpost_batch(ospec, thetas, X, y, threads=1):
    lib = load()
    X, y = _xy(ospec, X, y)
    thetas = np.ascontiguousarray(thetas, np.float32)
    Cn, d = thetas.shape
    lp, g = np.empty(Cn, np.float32), np.empty((Cn, d), np.float32)
    s = _spec(ospec)
    lib.mo_logpost_batch(C.byref(s), _p(thetas), Cn, _p(X), _p(y), X.shape[0], _p(lp), _p(g), threads)
    return lp, g


class Chains:
    """State of C chains driven by the C restatement."""

    def __init__(self, ospec, X, y, theta0, z0=None, seed=0, threads=1):
        self.lib, self.ospec, self.threads = load(), ospec, threads
        self.X, self.y = _xy(ospec, X, y)
        self.spec = _spec(ospec)
        self.theta = np.ascontiguousarray(theta0, np.float32).copy()
        self.C, self.d = self.theta.shape
        self.u = np.empty_like(self.theta)
        self.g = np.empty_like(self.theta)
        self.lp = np.empty(self.C, np.float32)
        z0 = None if z0 is None else np.ascontiguousarray(z0, np.float32)
        self.lib.mo_init(C.byref(self.spec), _p(self.X), _p(self.y), self.X.shape[0], self.C, _p(self.theta),
                         _p(self.u), _p(self.lp), _p(self.g), _p(z0), seed, threads)

    def sample(self, n_steps, eps, L, step_base=0, thin=1, z=None, seed=0, keep=True, info=False):
        eps = np.ascontiguousarray(np.broadcast_to(eps, (self.C,)), np.float32)
        L = np.ascontiguousarray(np.broadcast_to(L, (self.C,)), np.float32)
        first, last = -(-step_base // thin), (step_base + n_steps - 1) // thin
        samples = np.empty((max(0, last - first + 1), self.C, self.d), np.float32) if keep else None
        inf = np.empty((n_steps, self.C, 3), np.float32) if info else None
        z = None if z is None else np.ascontiguousarray(z, np.float32)
        self.lib.mo_run_sampling(C.byref(self.spec), _p(self.X), _p(self.y), self.X.shape[0], self.C, _p(self.theta),
                                 _p(self.u), _p(self.lp), _p(self.g), _p(eps), _p(L), n_steps, step_base, thin, _p(z),
                                 seed, _p(samples), _p(inf), self.threads)
        return samples, inf


def set_row_threads(n: int):
    """Threads that split the rows of one chain's evaluation (nested under the one-chain-per-thread loop)."""
    lib = load()
    lib.mo_max_threads()          # enables two active OpenMP levels
    lib.mo_set_row_threads(int(n))


def run_sampling_timed(ospec, X, y, theta0, n_steps, eps, L, threads=1, row_threads=1):
    """n_steps MCLMC steps for all chains with in-library Philox noise (used by bench.py's CPU legs).
    threads = chains in flight (one per OpenMP thread); row_threads = threads sharing the rows of each chain."""
    set_row_threads(row_threads)
    try:
        ch = Chains(ospec, X, y, theta0, seed=1, threads=threads)
        ch.sample(n_steps, eps, L, seed=2, keep=False)
    finally:
        set_row_threads(1)
    return ch
