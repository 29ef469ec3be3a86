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


# ---- single-chain checkers (R/convergence.R:259-389): host-side, on the mcmc object MCMC() returns ----------------------

def _rm_invariant_host(x):
    """rm_invariant (R/convergence.R:169-186) on a host object: ONE pooled variance (quirk D9); column 1 goes when it is < 1e-10."""
    from .coda import Mcmc
    data = np.vstack([m.data for m in x]) if isinstance(x, McmcList) else np.asarray(x.data)
    if data.size > 1 and np.var(data.ravel(), ddof=1) < 1e-10:
        if data.shape[1] == 1:
            return None                                           # the reference returns FALSE; the diag then errors
        if isinstance(x, McmcList):
            return x.select(list(range(1, x.nvar())))
        return Mcmc(x.data[:, 1:], *x.mcpar, x.varnames[1:])
    return x


def _record(x, d):
    dat = dict(LAST_CONV_CHECK.get("dat", {}))
    dat[x.mcpar[1]] = d
    convergence_data_set({"dat": dat})


class _HostChecker:
    freq = 1000

    def _single_chain(self, x, what):
        if (x.nchain() if hasattr(x, "nchain") else 1) > 1:
            raise ValueError(what)


class GewekeChecker(_HostChecker):
    """convergence_geweke (R/convergence.R:259-303): TRUE iff H0 'equal means of the first 10 % and the last 50 %' is not
    rejected at `threshold` for every parameter.  coda::geweke.diag is restated in diagnostics.py (third-party, unpinned)."""

    def __init__(self, freq=1000, threshold=0.025, check_invariant=True):
        self.freq, self.threshold, self.check_invariant = int(freq), float(threshold), bool(check_invariant)

    def __call__(self, x):
        from scipy.special import ndtr
        from .diagnostics import geweke_z
        self._single_chain(x, "The `geweke` convergence check is only available with runs of a single chain.")
        niter = x.niter()
        if self.check_invariant:
            x = _rm_invariant_host(x)
        try:
            if x is None:
                raise ValueError("every column is invariant")
            z = geweke_z(x)
        except Exception:                                       # R/convergence.R:274-282
            warnings.warn(f"At {niter} `geweke.diag` failed to be computed. Will skip and try with the next batch.")
            return False
        _record(x, z)
        fin = z[np.isfinite(z)]
        convergence_msg_set("avg Geweke's Z: %.4f." % (fin.mean() if fin.size else float("nan")))
        p = 1.0 - ndtr(-np.abs(z)) * 2.0
        if np.any(~np.isfinite(p)):
            return False
        return bool(np.all(p > self.threshold))


class HeidelChecker(_HostChecker):
    """convergence_heildel (R/convergence.R:308-355): TRUE iff every parameter passes both the stationarity and the half-width
    test of coda::heidel.diag (restated in diagnostics.py; third-party, unpinned)."""

    def __init__(self, freq=1000, check_invariant=True, eps=0.1, pvalue=0.05):
        self.freq, self.check_invariant, self.eps, self.pvalue = int(freq), bool(check_invariant), float(eps), float(pvalue)

    def __call__(self, x):
        from .diagnostics import heidel_diag
        self._single_chain(x, "The -heidel- convergence check is only available with runs of a single chain.")
        niter = x.niter()
        if self.check_invariant:
            x = _rm_invariant_host(x)
        try:
            if x is None:
                raise ValueError("every column is invariant")
            d = heidel_diag(x, self.eps, self.pvalue)
        except Exception:                                       # R/convergence.R:321-333
            warnings.warn(f"At {niter} -coda::heidel.diag- failed to be computed. Will skip and try with the next batch.")
            return False
        _record(x, d)
        convergence_msg_set("Heidel's Avg. pval: %.2f" % np.mean(d[:, 2]))
        tests = d[:, [0, 3]]
        if np.any(~np.isfinite(tests)):
            return False
        return bool(np.all(tests == 1))


class AutoChecker(_HostChecker):
    """convergence_auto (R/convergence.R:365-389): Gelman-Rubin when nchains > 1 (on the device), Geweke otherwise."""

    def __init__(self, freq=1000):
        self.freq = int(freq)
        self.gelman, self.geweke = GelmanChecker(freq), GewekeChecker(freq)


def convergence_geweke(freq=1000, threshold=0.025, check_invariant=True):
    return GewekeChecker(freq, threshold, check_invariant)


def convergence_heildel(freq=1000, check_invariant=True, eps=0.1, pvalue=0.05):
    return HeidelChecker(freq, check_invariant, eps, pvalue)


def convergence_auto(freq=1000):
    return AutoChecker(freq)
