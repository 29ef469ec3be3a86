"""Thin object wrapper over the C ABI handles (what the R glue keeps as an external pointer)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as A
from . import _lib


class DeviceModel:
    """X / y uploaded once to HBM (fmcmc_model_create) + the run / store / Gelman entry points."""

    def __init__(self, family, device: int = 0, device_ptrs=None):
        L = _lib.lib()
        self.family = family
        self.k = family.k
        self._h = C.c_void_p()
        err = _lib.errbuf()
        if device_ptrs is None:
            m = family.marshal()
            rc = L.fmcmc_model_create(m.byref(), device, C.byref(self._h), err, len(err))
        else:
            m = A.marshal_model(family.family, family.n, family.p_x, family.n_groups, flags=family.flags,
                                hyper=family.hyper, device_ptrs=device_ptrs)
            rc = L.fmcmc_model_create_device(m.byref(), device, C.byref(self._h), err, len(err))
        _lib.check(rc, err)
        self.device = device

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().fmcmc_model_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_path(self, path: int):
        rc = _lib.lib().fmcmc_set_path(self._h, path)
        if rc:
            raise _lib.FmcmcError(rc, "bad path")

    def mark(self, slot: int):
        if _lib.lib().fmcmc_event_mark(self._h, slot):
            raise _lib.FmcmcError(A.ECUDA, "fmcmc_event_mark failed")

    def elapsed_ms(self, a: int, b: int) -> float:
        v = C.c_double()
        if _lib.lib().fmcmc_event_elapsed_ms(self._h, a, b, C.byref(v)):
            raise _lib.FmcmcError(A.ECUDA, "fmcmc_event_elapsed_ms failed")
        return v.value

    def trim(self, path: int):
        """fmcmc_model_trim: free the device copies of X that stepping path `path` does not read."""
        err = _lib.errbuf()
        _lib.check(_lib.lib().fmcmc_model_trim(self._h, path, err, len(err)), err)

    def logpost(self, theta):
        theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        out = np.empty(theta.shape[0])
        err = _lib.errbuf()
        _lib.check(_lib.lib().fmcmc_logpost(self._h, theta.shape[0], A.ptr(theta), A.ptr(out), err, len(err)), err)
        return out

    def run(self, kernel_spec: dict, nsteps, nchains, initial=None, burnin=0, thin=1, stream=None,
            istate=None, dstate=None, flags=0, chain_offset=0, want_draws=True, outputs=True, nchains_total=0, into=None):
        """One MCMC_without_conv_checker call (R/mcmc.R:485-838) for `nchains` chains.
        into = (ans, draws or None, logpost, row_offset): write this call's kept rows into rows row_offset.. of caller-owned
        arrays [nchains][rows_total][k] / [nchains][rows_total] (the bulk loop fills one set of arrays, bulk after bulk);
        the returned dict then holds views of those rows."""
        L = _lib.lib()
        k = self.k
        ks = A.marshal_kernel(kernel_spec)
        if istate is None:
            istate = np.zeros((nchains, A.ISTATE_LEN), dtype=np.int64)
        st = A.marshal_state(istate, dstate)
        if initial is not None:
            initial = np.ascontiguousarray(np.broadcast_to(np.asarray(initial, dtype=np.float64), (nchains, k)))
        keep = A.rows_kept(nsteps, burnin, thin)
        rows_total = row_off = 0
        ans = draws = lp = None
        if into is not None:
            ans, draws, lp, row_off = into
            cm = bool(flags & A.RUN_COLMAJOR)                    # [chain][param][row] instead of [chain][row][param]
            rows_total = ans.shape[2 if cm else 1]
            if not want_draws:
                draws = None
            for a in (ans, draws, lp):
                if a is not None and not (a.flags.c_contiguous and a.dtype == np.float64 and a.shape[0] == nchains):
                    raise ValueError("`into` arrays must be C-contiguous float64 with one block per chain")
        elif outputs:
            ans = np.empty((nchains, keep, k))
            lp = np.empty((nchains, keep))
            if want_draws:
                draws = np.empty((nchains, keep, k))
        rs = A.marshal_run(nsteps, nchains, initial, burnin, thin, flags | (0 if (outputs or into is not None) else A.RUN_NO_OUTPUT)
                           | (0 if want_draws else A.RUN_NO_DRAWS), chain_offset, nchains_total, rows_total, row_off)
        if stream is None:
            stream = A.marshal_stream()
        rep = A.RunReport()
        err = _lib.errbuf()
        rc = L.fmcmc_run(self._h, rs.byref(), ks.byref(), st.byref(), stream.byref(),
                         A.ptr(ans) if ans is not None else None,
                         A.ptr(draws) if draws is not None else None,
                         A.ptr(lp) if lp is not None else None, C.byref(rep), err, len(err))
        if rc:
            e = _lib.FmcmcError(rc, err.value.decode(errors="replace"))
            e.report = rep
            raise e
        if into is not None:
            sl = slice(row_off, row_off + keep)
            ans, lp = (ans[:, :, sl] if cm else ans[:, sl]), lp[:, sl]
            if draws is not None:
                draws = draws[:, :, sl] if cm else draws[:, sl]
        return dict(ans=ans, draws=draws, logpost=lp, report=rep, istate=istate, dstate=dstate)

    def fetch_state(self, istate, dstate=None):
        """fmcmc_kernel_state_fetch: the resident kernel state of the last run -> the given host arrays."""
        st = A.marshal_state(istate, dstate)
        err = _lib.errbuf()
        _lib.check(_lib.lib().fmcmc_kernel_state_fetch(self._h, st.byref(), err, len(err)), err)

    # ---- observation sharding across GPUs (include/fmcmc_b200.h: fmcmc_shard_*) -----------------
    def shard_alloc(self, world: int, max_cols: int, n_total: int) -> "A.ShardHandles":
        h = A.ShardHandles()
        err = _lib.errbuf()
        _lib.check(_lib.lib().fmcmc_shard_alloc(self._h, world, max_cols, n_total, C.byref(h), err, len(err)), err)
        return h

    def shard_attach(self, rank: int, world: int, handles) -> None:
        arr = (A.ShardHandles * world)(*handles)
        err = _lib.errbuf()
        _lib.check(_lib.lib().fmcmc_shard_attach(self._h, rank, world, arr, err, len(err)), err)

    # ---- sample store + Gelman -------------------------------------------------------------
    def store_reset(self, nchains, capacity_rows):
        err = _lib.errbuf()
        _lib.check(_lib.lib().fmcmc_store_reset(self._h, nchains, self.k, capacity_rows, err, len(err)), err)

    def store_rows(self) -> int:
        return _lib.lib().fmcmc_store_rows(self._h)

    def store_pooled(self, free_mask):
        """(count, mean, M2) of every element of this GPU's part of the store: rm_invariant's pooled variance (D9)."""
        mask = np.ascontiguousarray(free_mask, dtype=np.uint8)
        out = np.empty(3)
        err = _lib.errbuf()
        _lib.check(_lib.lib().fmcmc_store_pooled(self._h, A.ptr(mask, C.POINTER(C.c_uint8)), A.ptr(out), err, len(err)), err)
        return out

    def store_ess(self, row_begin, row_end, free_mask, nchains, max_lag=0):
        """fmcmc_store_ess: (ESS[nchains][kf], truncated) of the stored series, computed on the device."""
        mask = np.ascontiguousarray(free_mask, dtype=np.uint8)
        ess = np.empty((nchains, int(mask.sum())))
        tr = C.c_int32()
        err = _lib.errbuf()
        _lib.check(_lib.lib().fmcmc_store_ess(self._h, row_begin, row_end, A.ptr(mask, C.POINTER(C.c_uint8)), int(max_lag),
                                              A.ptr(ess), C.byref(tr), err, len(err)), err)
        return ess, bool(tr.value)

    def gelman_partials(self, row_begin, row_end, free_mask, nchains, out=None):
        """Host arrays by default; with `out=(xbar_ptr, s2_ptr, wsum_ptr)` raw device pointers."""
        mask = np.ascontiguousarray(free_mask, dtype=np.uint8)
        kf = int(mask.sum())
        err = _lib.errbuf()
        if out is None:
            xbar, s2, ws = np.empty((nchains, kf)), np.empty((nchains, kf)), np.empty((kf, kf), order="F")
            rc = _lib.lib().fmcmc_gelman_partials(self._h, row_begin, row_end, A.ptr(mask, C.POINTER(C.c_uint8)),
                                                  xbar.ctypes.data, s2.ctypes.data, ws.ctypes.data, 0, err, len(err))
            _lib.check(rc, err)
            return xbar, s2, ws
        rc = _lib.lib().fmcmc_gelman_partials(self._h, row_begin, row_end, A.ptr(mask, C.POINTER(C.c_uint8)),
                                              out[0], out[1], out[2], 1, err, len(err))
        _lib.check(rc, err)
        return None

    def gelman(self, free_mask, start_iter=1, thin=1):
        """fmcmc_gelman: coda's autoburnin window + statistics + finish in one call (single GPU).
        Returns (psrf, mpsrf, niter_used)."""
        mask = np.ascontiguousarray(free_mask, dtype=np.uint8)
        psrf = np.empty(int(mask.sum()))
        mpsrf, used = C.c_double(), C.c_int64()
        err = _lib.errbuf()
        rc = _lib.lib().fmcmc_gelman(self._h, A.ptr(mask, C.POINTER(C.c_uint8)), int(start_iter), int(thin),
                                     A.ptr(psrf), C.byref(mpsrf), C.byref(used), err, len(err))
        _lib.check(rc, err)
        return psrf, mpsrf.value, used.value

    def gelman_finish(self, niter, nchains_total, kf, xbar, s2, wsum, dev_in=False):
        psrf = np.empty(kf)
        mpsrf = C.c_double()
        err = _lib.errbuf()
        if dev_in:
            px, ps, pw = xbar, s2, wsum
        else:
            xbar = np.ascontiguousarray(xbar, dtype=np.float64)
            s2 = np.ascontiguousarray(s2, dtype=np.float64)
            wsum = np.asfortranarray(wsum, dtype=np.float64)
            px, ps, pw = xbar.ctypes.data, s2.ctypes.data, wsum.ctypes.data
        rc = _lib.lib().fmcmc_gelman_finish(self._h, niter, nchains_total, kf, px, ps, pw, 1 if dev_in else 0,
                                            A.ptr(psrf), C.byref(mpsrf), err, len(err))
        _lib.check(rc, err)
        return psrf, mpsrf.value


def cov_recursive(X_t, Cov_t, Mean_t_prev, t_, Mean_t=None, eps=0.0, Sd=1.0, Ik=None, device=0):
    """R/recursive.R:63-120 on the device.  Returns (Mean_t, Cov_t) after the last row of X_t."""
    X = np.ascontiguousarray(np.atleast_2d(X_t), dtype=np.float64)
    rows, k = X.shape
    mp = np.ascontiguousarray(Mean_t_prev, dtype=np.float64).reshape(k)
    cp = np.asfortranarray(Cov_t, dtype=np.float64)
    ik = np.asfortranarray(Ik, dtype=np.float64) if Ik is not None else None
    mo, co = np.empty(k), np.empty((k, k), order="F")
    err = _lib.errbuf()
    rc = _lib.lib().fmcmc_cov_recursive(device, k, rows, A.ptr(X), A.ptr(mp), A.ptr(cp), float(t_), float(eps),
                                        float(Sd), A.ptr(ik) if ik is not None else None, A.ptr(mo), A.ptr(co),
                                        err, len(err))
    _lib.check(rc, err)
    return mo, co


def mean_recursive(X_t, Mean_t_prev, t_, device=0):
    """R/recursive.R:124-139 on the device."""
    X = np.atleast_2d(X_t)
    k = X.shape[1]
    return cov_recursive(X, np.zeros((k, k)), Mean_t_prev, t_, device=device)[0]


def reflect_on_boundaries(x, lb, ub, which=None, device=0):
    """R/kernel.R:450-493 on the device; `which` is 1-based like in R (None = all)."""
    x = np.array(np.atleast_2d(x), dtype=np.float64, order="C")
    count, k = x.shape
    lb = np.ascontiguousarray(np.broadcast_to(np.asarray(lb, dtype=np.float64), (k,)))
    ub = np.ascontiguousarray(np.broadcast_to(np.asarray(ub, dtype=np.float64), (k,)))
    mask = None
    if which is not None:
        mask = np.zeros(k, dtype=np.uint8)
        mask[np.asarray(which, dtype=int) - 1] = 1
    err = _lib.errbuf()
    rc = _lib.lib().fmcmc_reflect(device, k, count, A.ptr(x), A.ptr(lb), A.ptr(ub),
                                  A.ptr(mask, C.POINTER(C.c_uint8)) if mask is not None else None, err, len(err))
    _lib.check(rc, err)
    return x if count > 1 else x[0]
