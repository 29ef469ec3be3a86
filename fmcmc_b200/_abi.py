"""ctypes mirror of include/fmcmc_b200.h (the C ABI of libfmcmcb200.so).

This is the Python twin of the R `.Call` shim shown in INTEGRATION.md: plain
PODs, host pointers and sizes only.  Nothing in here computes anything.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

ABI_VERSION = 3

# status codes
OK, EINVAL, ENAN, ECUDA, ENOMEM, ENOTPD, EUNSUP, ENANRATIO = range(8)

# families
FAMILY_GAUSSIAN_LM, FAMILY_LOGISTIC, FAMILY_HIER_NORMAL = 1, 2, 3
MODEL_INTERCEPT, MODEL_GUARD, MODEL_SCALES = 1, 2, 4

# kernels
(KERNEL_NORMAL, KERNEL_NORMAL_REFLECTIVE, KERNEL_UNIF, KERNEL_UNIF_REFLECTIVE,
 KERNEL_ADAPT, KERNEL_RAM, KERNEL_NMIRROR, KERNEL_UMIRROR) = range(1, 9)
SCHEME_JOINT, SCHEME_ORDERED, SCHEME_RANDOM, SCHEME_EXPLICIT = range(4)
MVN_CHOLESKY, MVN_EIGEN = 0, 1

ISTATE_LEN = 4
STATE_HAS_MEAN, STATE_INIT, STATE_OBS_SHIFT = 1, 2, 2

STREAM_PHILOX, STREAM_FED = 0, 1

RUN_NO_DRAWS, RUN_COLMAJOR, RUN_APPEND, RUN_NO_OUTPUT, RUN_DEVICE_STATE, RUN_KEEP_STATE = 1, 2, 4, 8, 16, 32

DBL_MAX = float(np.finfo(np.float64).max)

_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)


class ModelDesc(C.Structure):
    _fields_ = [
        ("family", C.c_int32), ("flags", C.c_uint32), ("n", C.c_int64),
        ("p_x", C.c_int32), ("n_groups", C.c_int32),
        ("X", C.c_void_p), ("y", C.c_void_p), ("group", C.c_void_p),
        ("hyper", C.c_double * 4),
    ]


class KernelSpec(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("k", C.c_int32), ("scheme", C.c_int32), ("order_len", C.c_int32),
        ("order", _i32p), ("seq", _i32p), ("seq_len", C.c_int64),
        ("mu", _dp), ("scale", _dp), ("min_", _dp), ("max_", _dp), ("lb", _dp), ("ub", _dp),
        ("fixed", _u8p),
        ("warmup", C.c_int64), ("freq", C.c_int64), ("bw", C.c_int64),
        ("until", C.c_double), ("eps", C.c_double), ("Sd", C.c_double), ("arate", C.c_double),
        ("nadapt", _i64p), ("nadapt_len", C.c_int32), ("mvn_method", C.c_int32),
        ("constr", _dp),
    ]


class KernelState(C.Structure):
    _fields_ = [("istate", _i64p), ("dstate", _dp)]


class StreamSpec(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("kdraw", C.c_int32), ("seed", C.c_uint64),
        ("run_index", C.c_uint64), ("logu", _dp), ("z", _dp),
    ]


class RunSpec(C.Structure):
    _fields_ = [
        ("nsteps", C.c_int64), ("burnin", C.c_int64), ("thin", C.c_int64),
        ("nchains", C.c_int32), ("flags", C.c_uint32), ("chain_offset", C.c_int64),
        ("initial", _dp), ("nchains_total", C.c_int64), ("out_rows_total", C.c_int64), ("out_row_offset", C.c_int64),
    ]


class RunReport(C.Structure):
    _fields_ = [
        ("rows_kept", C.c_int64), ("first_iter", C.c_int64), ("last_iter", C.c_int64),
        ("nan_chain", C.c_int64), ("nan_step", C.c_int64), ("n_accept", C.c_int64),
        ("n_launches", C.c_int64), ("device_ms", C.c_double), ("path", C.c_int32),
        ("reserved", C.c_int32), ("hot_ms", C.c_double), ("hot_launches", C.c_int64),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
    ]


class ShardHandles(C.Structure):
    """fmcmc_shard_handles (observation sharding across GPUs)."""
    _fields_ = [
        ("partial", C.c_ubyte * 64), ("flags", C.c_ubyte * 64),
        ("partial_ptr", C.c_void_p), ("flags_ptr", C.c_void_p),
        ("device", C.c_int32), ("pid", C.c_int32),
    ]


def _as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def ptr(a, typ=_dp):
    return a.ctypes.data_as(typ) if a is not None else typ()


def rows_kept(nsteps: int, burnin: int, thin: int) -> int:
    """R/mcmc.R:786-813: which((1:(nsteps-burnin) %% thin) == 0)."""
    return max(nsteps - burnin, 0) // max(thin, 1)


def state_len(ktype: int, k: int, kf: int) -> int:
    if ktype == KERNEL_ADAPT:
        return kf * kf + kf
    if ktype == KERNEL_RAM:
        return kf * kf
    if ktype in (KERNEL_NMIRROR, KERNEL_UMIRROR):
        return 3 * k
    return 0


class Marshalled:
    """Keeps the numpy buffers behind a ctypes struct alive."""

    def __init__(self, struct, keep):
        self.struct = struct
        self.keep = keep

    def byref(self):
        return C.byref(self.struct)


def marshal_model(family, n, p_x=0, n_groups=0, X=None, y=None, group=None, flags=0,
                  hyper=(0.0, 0.0, 0.0, 0.0), device_ptrs=None) -> Marshalled:
    d = ModelDesc()
    d.family, d.flags, d.n, d.p_x, d.n_groups = family, flags, n, p_x, n_groups
    keep = []
    if device_ptrs is not None:
        d.X, d.y, d.group = device_ptrs
    else:
        if X is not None:
            X = np.asfortranarray(X, dtype=np.float64)  # column-major, R layout
            if X.ndim == 1:
                X = X.reshape(-1, 1, order="F")
            keep.append(X)
            d.X = X.ctypes.data
        if y is not None:
            y = _as_f64(y)
            keep.append(y)
            d.y = y.ctypes.data
        if group is not None:
            group = np.ascontiguousarray(group, dtype=np.int32)
            keep.append(group)
            d.group = group.ctypes.data
    for i, h in enumerate(hyper):
        d.hyper[i] = h
    return Marshalled(d, keep)


def marshal_kernel(spec: dict) -> Marshalled:
    """spec: plain dict produced by fmcmc_b200.kernels.FmcmcKernel.to_spec(k)."""
    ks = KernelSpec()
    keep = []
    k = spec["k"]
    ks.type, ks.k, ks.scheme = spec["type"], k, spec.get("scheme", SCHEME_JOINT)

    def vec(name, default):
        v = spec.get(name)
        if v is None:
            v = default
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (k,)))
        keep.append(a)
        return ptr(a)

    ks.mu = vec("mu", 0.0)
    ks.scale = vec("scale", 1.0)
    ks.min_ = vec("min_", -1.0)
    ks.max_ = vec("max_", 1.0)
    ks.lb = vec("lb", -DBL_MAX)
    ks.ub = vec("ub", DBL_MAX)
    fixed = np.ascontiguousarray(
        np.broadcast_to(np.asarray(spec.get("fixed", False), dtype=bool), (k,))).astype(np.uint8)
    keep.append(fixed)
    ks.fixed = ptr(fixed, _u8p)
    order = spec.get("order")
    if order is not None:
        order = np.ascontiguousarray(order, dtype=np.int32)
        keep.append(order)
        ks.order = ptr(order, _i32p)
        ks.order_len = order.size
    seq = spec.get("seq")
    if seq is not None:
        seq = np.ascontiguousarray(seq, dtype=np.int32)
        keep.append(seq)
        ks.seq = ptr(seq, _i32p)
        ks.seq_len = seq.shape[-1]
    ks.warmup = int(spec.get("warmup", 0))
    ks.freq = int(spec.get("freq", 1))
    ks.bw = int(spec.get("bw", 0))
    until = spec.get("until", math.inf)
    ks.until = float(until)
    ks.eps = float(spec.get("eps", 1e-4))
    sd = spec.get("Sd")
    ks.Sd = float(sd) if sd is not None else 0.0
    ks.arate = float(spec.get("arate", 0.234))
    nadapt = spec.get("nadapt")
    if nadapt is not None:
        nadapt = np.ascontiguousarray(nadapt, dtype=np.int64)
        keep.append(nadapt)
        ks.nadapt = ptr(nadapt, _i64p)
        ks.nadapt_len = nadapt.size
    ks.mvn_method = int(spec.get("mvn_method", MVN_CHOLESKY))
    constr = spec.get("constr")
    if constr is not None:
        constr = np.asfortranarray(constr, dtype=np.float64)
        keep.append(constr)
        ks.constr = ptr(constr)
    return Marshalled(ks, keep)


def marshal_state(istate: np.ndarray, dstate: np.ndarray | None) -> Marshalled:
    st = KernelState()
    st.istate = ptr(istate, _i64p)
    st.dstate = ptr(dstate) if dstate is not None and dstate.size else _dp()
    return Marshalled(st, [istate, dstate])


def marshal_stream(mode=STREAM_PHILOX, seed=0, run_index=0, logu=None, z=None, kdraw=0) -> Marshalled:
    ss = StreamSpec()
    keep = []
    ss.mode, ss.seed, ss.run_index, ss.kdraw = mode, seed & 0xFFFFFFFFFFFFFFFF, run_index, kdraw
    if logu is not None:
        logu = _as_f64(logu)
        keep.append(logu)
        ss.logu = ptr(logu)
    if z is not None:
        z = _as_f64(z)
        keep.append(z)
        ss.z = ptr(z)
        ss.kdraw = z.shape[-1] if z.ndim == 3 else kdraw
    return Marshalled(ss, keep)


def marshal_run(nsteps, nchains, initial=None, burnin=0, thin=1, flags=0, chain_offset=0, nchains_total=0,
                out_rows_total=0, out_row_offset=0) -> Marshalled:
    rs = RunSpec()
    keep = []
    rs.nsteps, rs.burnin, rs.thin = nsteps, burnin, thin
    rs.nchains, rs.flags, rs.chain_offset = nchains, flags, chain_offset
    rs.nchains_total = nchains_total
    rs.out_rows_total, rs.out_row_offset = out_rows_total, out_row_offset
    if initial is not None:
        initial = _as_f64(initial)
        keep.append(initial)
        rs.initial = ptr(initial)
    return Marshalled(rs, keep)


RUN_ARGTYPES = [C.c_void_p, C.POINTER(RunSpec), C.POINTER(KernelSpec), C.POINTER(KernelState),
                C.POINTER(StreamSpec), _dp, _dp, _dp, C.POINTER(RunReport), C.c_char_p, C.c_size_t]
