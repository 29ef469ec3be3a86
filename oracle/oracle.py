"""ctypes wrapper of the CPU oracle — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from fmcmc_b200 import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfmcmc_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("fmcmc_oracle.c", "r_rng.c", "Makefile")]
    srcs.append(os.path.join(_HERE, "..", "include", "fmcmc_b200.h"))
    stale = force or not os.path.exists(_SO) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        L.fmcmc_oracle_run.restype = C.c_int
        L.fmcmc_oracle_run.argtypes = [C.POINTER(A.ModelDesc)] + A.RUN_ARGTYPES[1:]
        L.fmcmc_oracle_logpost.restype = C.c_double
        L.fmcmc_oracle_logpost.argtypes = [C.POINTER(A.ModelDesc), dp]
        L.fmcmc_oracle_nparams.restype = C.c_int32
        L.fmcmc_oracle_nparams.argtypes = [C.POINTER(A.ModelDesc)]
        L.fmcmc_oracle_gelman.restype = C.c_int
        L.fmcmc_oracle_gelman.argtypes = [C.c_int64, C.c_int64, C.c_int, dp, dp, dp]
        L.fmcmc_oracle_reflect.restype = None
        L.fmcmc_oracle_reflect.argtypes = [C.c_int, dp, dp, dp, C.POINTER(C.c_uint8)]
        L.fmcmc_oracle_cov_recursive.restype = None
        L.fmcmc_oracle_cov_recursive.argtypes = [C.c_int, C.c_int64, dp, dp, dp, C.c_double,
                                                 C.c_double, C.c_double, dp, dp, dp]
        L.fmcmc_oracle_philox_u2.restype = None
        L.fmcmc_oracle_philox_u2.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                             C.c_uint32, dp, dp]
        L.fmcmc_oracle_set_threads.argtypes = [C.c_int]
        L.fmcmc_oracle_pooled_invariant.restype = C.c_int
        L.fmcmc_oracle_pooled_invariant.argtypes = [C.c_int64, dp]
        # R RNG layer
        L.r_set_seed.argtypes = [C.c_uint32]
        L.r_unif_rand.restype = C.c_double
        L.r_norm_rand.restype = C.c_double
        L.r_qnorm.restype = C.c_double
        L.r_qnorm.argtypes = [C.c_double]
        L.r_mt_u32.restype = C.c_uint32
        L.r_get_state.argtypes = [C.POINTER(C.c_uint32)]
        L.r_put_state.argtypes = [C.POINTER(C.c_uint32)]
        L.r_rnorm.argtypes = [C.c_int64, C.c_double, C.c_double, dp]
        L.r_runif.argtypes = [C.c_int64, C.c_double, C.c_double, dp]
        L.r_norm_rand_vec.argtypes = [C.c_int64, dp]
        L.r_log_runif.argtypes = [C.c_int64, dp]
        L.r_rt.argtypes = [C.c_int64, C.c_double, dp]
        L.r_rgamma_vec.argtypes = [C.c_int64, C.c_double, C.c_double, dp]
        L.r_exp_rand_vec.argtypes = [C.c_int64, dp]
        L.r_exp_rand_q.argtypes = [dp]
        L.r_sd.restype = C.c_double
        L.r_sd.argtypes = [dp, C.c_int64]
        _lib = L
    return _lib


# ---------------------------------------------------------------- R RNG ----
class RRng:
    """R's default RNG (Mersenne-Twister + Inversion), global state like R."""

    @staticmethod
    def set_seed(seed: int):
        lib().r_set_seed(C.c_uint32(seed & 0xFFFFFFFF))

    @staticmethod
    def rnorm(n, mean=0.0, sd=1.0):
        out = np.empty(n)
        lib().r_rnorm(n, mean, sd, A.ptr(out))
        return out

    @staticmethod
    def runif(n, a=0.0, b=1.0):
        out = np.empty(n)
        lib().r_runif(n, a, b, A.ptr(out))
        return out

    @staticmethod
    def norm_rand(n):
        out = np.empty(n)
        lib().r_norm_rand_vec(n, A.ptr(out))
        return out

    @staticmethod
    def log_runif(n):
        out = np.empty(n)
        lib().r_log_runif(n, A.ptr(out))
        return out

    @staticmethod
    def rt(n, df):
        out = np.empty(n)
        lib().r_rt(n, float(df), A.ptr(out))
        return out

    @staticmethod
    def rgamma(n, shape, scale=1.0):
        out = np.empty(n)
        lib().r_rgamma_vec(n, float(shape), float(scale), A.ptr(out))
        return out

    @staticmethod
    def rexp(n):
        out = np.empty(n)
        lib().r_exp_rand_vec(n, A.ptr(out))
        return out

    @staticmethod
    def sd(x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return lib().r_sd(A.ptr(x), x.size)


# ------------------------------------------------------------- the loop ----
class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def logpost(model: A.Marshalled, theta) -> float:
    th = np.ascontiguousarray(theta, dtype=np.float64)
    return lib().fmcmc_oracle_logpost(model.byref(), A.ptr(th))


def run(model: A.Marshalled, kernel_spec: dict, initial, nsteps, nchains=1, burnin=0, thin=1,
        stream: A.Marshalled | None = None, istate=None, dstate=None, chain_offset=0, threads=1):
    """One MCMC_without_conv_checker call.  Returns dict(ans, draws, logpost, report)."""
    L = lib()
    k = kernel_spec["k"]
    initial = np.ascontiguousarray(np.broadcast_to(np.asarray(initial, dtype=np.float64),
                                                   (nchains, k)))
    ks = A.marshal_kernel(kernel_spec)
    fixed = np.broadcast_to(np.asarray(kernel_spec.get("fixed", False), dtype=bool), (k,))
    kf = int((~fixed).sum())
    dlen = A.state_len(kernel_spec["type"], k, kf)
    if istate is None:
        istate = np.zeros((nchains, A.ISTATE_LEN), dtype=np.int64)
    if dstate is None:
        dstate = np.zeros((nchains, max(dlen, 1)), dtype=np.float64)
    st = A.marshal_state(istate, dstate if dlen else None)
    rs = A.marshal_run(nsteps, nchains, initial, burnin, thin, 0, chain_offset)
    if stream is None:
        stream = A.marshal_stream()
    keep = A.rows_kept(nsteps, burnin, thin)
    ans = np.empty((nchains, keep, k))
    draws = np.empty((nchains, keep, k))
    lp = np.empty((nchains, keep))
    rep = A.RunReport()
    err = C.create_string_buffer(512)
    L.fmcmc_oracle_set_threads(threads)
    rc = L.fmcmc_oracle_run(model.byref(), rs.byref(), ks.byref(), st.byref(), stream.byref(),
                            A.ptr(ans), A.ptr(draws), A.ptr(lp), C.byref(rep), err, 512)
    if rc != 0:
        raise OracleError(rc, err.value.decode())
    return dict(ans=ans, draws=draws, logpost=lp, report=rep, istate=istate, dstate=dstate)


def gelman(x):
    """x: [m][N][p] already windowed.  Returns (psrf[p], mpsrf, rc)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    m, N, p = x.shape
    psrf = np.empty(p)
    mpsrf = C.c_double()
    rc = lib().fmcmc_oracle_gelman(m, N, p, A.ptr(x), A.ptr(psrf), C.byref(mpsrf))
    return psrf, mpsrf.value, rc


def reflect(x, lb, ub, which=None):
    x = np.array(x, dtype=np.float64)
    k = x.size
    lb = np.ascontiguousarray(np.broadcast_to(np.asarray(lb, dtype=np.float64), (k,)))
    ub = np.ascontiguousarray(np.broadcast_to(np.asarray(ub, dtype=np.float64), (k,)))
    w = None if which is None else np.ascontiguousarray(which, dtype=np.uint8)
    lib().fmcmc_oracle_reflect(k, A.ptr(x), A.ptr(lb), A.ptr(ub),
                               A.ptr(w, C.POINTER(C.c_uint8)) if w is not None else None)
    return x


def cov_recursive(X, mean_prev, cov_prev, t, eps=0.0, Sd=1.0, Ik=None):
    X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
    rows, k = X.shape
    mean_prev = np.ascontiguousarray(mean_prev, dtype=np.float64)
    cov_prev = np.asfortranarray(cov_prev, dtype=np.float64)
    Ikp = np.asfortranarray(Ik, dtype=np.float64) if Ik is not None else None
    mean_out = np.empty(k)
    cov_out = np.empty((k, k), order="F")
    lib().fmcmc_oracle_cov_recursive(k, rows, A.ptr(X), A.ptr(mean_prev), A.ptr(cov_prev),
                                     float(t), float(eps), float(Sd),
                                     A.ptr(Ikp) if Ikp is not None else None,
                                     A.ptr(mean_out), A.ptr(cov_out))
    return mean_out, cov_out


def philox_u2(seed, chain, run, row, slot):
    u0, u1 = C.c_double(), C.c_double()
    lib().fmcmc_oracle_philox_u2(seed, chain, run, row, slot, C.byref(u0), C.byref(u1))
    return u0.value, u1.value
