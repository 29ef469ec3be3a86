"""convergence_gelman (R/convergence.R:191-246) with its arithmetic on the device, plus the
LAST_CONV_CHECK side channel (R/convergence.R:28-165)."""
from __future__ import annotations

import warnings

import numpy as np

from . import _lib
from .coda import McmcList, window_first_row

LAST_CONV_CHECK = {"msg": None}


def convergence_data_flush():
    LAST_CONV_CHECK.clear()
    LAST_CONV_CHECK["msg"] = None


def convergence_data_set(x: dict):
    if not isinstance(x, dict):
        raise TypeError("-x- must be a named list.")
    if "msg" in x:
        raise ValueError("'msg' should be set using -convergence_msg_set-")
    LAST_CONV_CHECK.update(x)


def convergence_data_get(x):
    if isinstance(x, (list, tuple)):
        return {n: LAST_CONV_CHECK[n] for n in x}
    return LAST_CONV_CHECK[x]


def convergence_msg_set(msg=None):
    if msg is not None and not isinstance(msg, str):
        raise TypeError("-msg- must be a character.")
    LAST_CONV_CHECK["msg"] = msg


def convergence_msg_get():
    return LAST_CONV_CHECK.get("msg")


class GelmanChecker:
    """The object `convergence_gelman()` returns: callable(x) -> bool with attribute `freq`.
    MCMC() recognises it and feeds it the device-resident sample store instead of x."""

    def __init__(self, freq=1000, threshold=1.10, check_invariant=True):
        self.freq, self.threshold, self.check_invariant = int(freq), float(threshold), bool(check_invariant)
        self._device_ctx = None   # set by MCMC(): (model, nchains_local, free_mask, dist-info)
        self._last_kf = None
        self.timings = None       # a dict here receives the stage timings of every multi-GPU check (bench.py)

    # -- device path: statistics from the store on the GPU(s) ------------------------------------
    def _device_diag(self, ans):
        model, nlocal, free_mask, dist = self._device_ctx
        start, end, thin = ans.mcpar
        rows = model.store_rows()
        if self.check_invariant:
            # rm_invariant (R/convergence.R:169-186), literally (quirk D9): ONE pooled variance of every element of
            # rbind(chains) - all accumulated rows, not the window - against 1e-10; when it is below, column 1 goes, and with a
            # single column the reference hands FALSE to gelman.diag, whose error becomes the warning + FALSE below
            from .dist import combine_pooled
            var = dist.pooled_variance(model, free_mask) if dist is not None else \
                combine_pooled([model.store_pooled(free_mask)])
            if var < 1e-10:
                free_mask = np.array(free_mask, dtype=np.uint8)
                if int(free_mask.sum()) <= 1:
                    raise _lib.FmcmcError(1, "every column is invariant (rm_invariant returned FALSE)")
                free_mask[np.flatnonzero(free_mask)[0]] = 0
        first = 0
        if start < end / 2:                                   # coda autoburnin = TRUE
            first = window_first_row(start, end, thin, rows, end / 2 + 1)
        niter = rows - first
        kf = int(np.sum(free_mask))
        self._last_kf = kf
        if dist is None:
            # one GPU: window + statistics + finish in ONE library call, nothing but psrf / mpsrf crosses the bus
            # (fmcmc_gelman_window_begin is the same arithmetic as window_first_row; tests/test_host_logic.py pins both)
            psrf, mpsrf, used = model.gelman(free_mask, start, thin)
            assert used == niter, (used, niter)
            return (psrf, mpsrf), niter
        return dist.gelman(model, first, rows, free_mask, nlocal, kf, niter, timings=self.timings), niter

    def __call__(self, x):
        nchain = x.nchain() if hasattr(x, "nchain") else 1
        if nchain <= 1 and (self._device_ctx is None or self._device_ctx[3] is None):
            raise ValueError("Convergence test with the Gelman is only available when `nchains` > 1L.")
        if self._device_ctx is None:
            raise TypeError("convergence_gelman() runs on the device and must be called through MCMC().")
        try:
            (psrf, mpsrf), niter = self._device_diag(x)
        except _lib.FmcmcError as e:                           # R/convergence.R:207-217
            warnings.warn(f"At {x.niter()} `gelman.diag` failed to be computed. Will skip and try with the "
                          f"next batch. ({e})")
            return False
        d = dict(psrf=psrf, mpsrf=mpsrf, niter=niter)
        dat = dict(LAST_CONV_CHECK.get("dat", {}))
        dat[x.mcpar[1]] = d
        convergence_data_set({"dat": dat})
        val = mpsrf if (self._last_kf or x.nvar()) > 1 else psrf[0]   # R/convergence.R:229 (nvar after rm_invariant)
        convergence_msg_set("Gelman-Rubin's R: %.4f." % val)
        return bool(val < self.threshold)


def convergence_gelman(freq=1000, threshold=1.10, check_invariant=True):
    return GelmanChecker(freq, threshold, check_invariant)
