"""Single-chain convergence diagnostics behind convergence_geweke / convergence_heildel (R/convergence.R:259-355).

The reference calls coda::geweke.diag and coda::heidel.diag, which in turn call coda::spectrum0.ar -> stats::ar (Yule-Walker
with AIC order selection).  None of that code is under /root/reference (coda and stats are third-party): the
algorithms are restated here from their published sources, and **parity with R is unpinned** - the reference's own tests
only pin reproducibility and error behaviour for these checkers (inst/tinytest/test-convergence.R:19-121).  They are
single-chain, O(T k) host-side bookkeeping on the mcmc object MCMC() returns (SURVEY section 2 marks them host-side); the
data-parallel diagnostics - Gelman-Rubin, the pooled variance of rm_invariant, ESS - run on the device
(csrc/gelman.cuh).  tests/test_host_logic.py checks the pieces against independent formulations (Toeplitz solves for
every AR order, the 5 % point of the Cramer-von Mises law, AR(1) series with known spectral density)."""
from __future__ import annotations

import math

import numpy as np


def ar_yule_walker(x):
    """stats::ar(x, aic = TRUE, method = "yule-walker"): returns (coefficients, var.pred, order).
    order.max = min(n - 1, floor(10 log10 n)); autocovariances with divisor n about the sample mean; Levinson-Durbin;
    AIC_k = n log(v_k) + 2 k + 2; var.pred = v_order * n / (n - (order + 1))."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    omax = int(min(n - 1, math.floor(10 * math.log10(n))))
    xc = x - x.mean()
    r = np.array([np.dot(xc[:n - l], xc[l:]) / n for l in range(omax + 1)])
    if not r[0] > 0:
        raise ValueError("zero-variance series")
    coefs = [np.zeros(0)]
    v = [r[0]]
    phi = np.zeros(0)
    for kk in range(1, omax + 1):                                # Levinson-Durbin recursion
        acc = r[kk] - np.dot(phi, r[kk - 1:0:-1]) if kk > 1 else r[1]
        refl = acc / v[-1]
        phi = np.concatenate([phi - refl * phi[::-1], [refl]])
        v.append(v[-1] * (1.0 - refl * refl))
        coefs.append(phi.copy())
    v = np.array(v)
    with np.errstate(divide="ignore", invalid="ignore"):
        aic = n * np.log(v) + 2.0 * np.arange(omax + 1) + 2.0
    aic = np.where(np.isfinite(aic), aic, np.inf)
    order = int(np.argmin(aic))
    return coefs[order], float(v[order] * n / (n - (order + 1))), order


def spectrum0_ar(x):
    """coda::spectrum0.ar for one series: spectral density at frequency zero from the fitted AR model, 0 when the series has no
    variation about its linear trend."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    z = np.arange(1, n + 1, dtype=np.float64)
    zc = z - z.mean()
    slope = np.dot(zc, x - x.mean()) / np.dot(zc, zc)
    resid = (x - x.mean()) - slope * zc
    sd = math.sqrt(np.dot(resid, resid) / max(n - 1, 1))
    if sd < 1.5e-8:                                              # all.equal(sd(residuals), 0)
        return 0.0, 0
    ar, var_pred, order = ar_yule_walker(x)
    return var_pred / (1.0 - ar.sum()) ** 2, order


def _window_rows(start, end, thin, niter, wstart=None, wend=None):
    """Rows (0-based, half-open) coda's window.mcmc keeps for iterations [wstart, wend]: the start snaps UP and the end snaps
    DOWN to the iteration grid start + thin * row."""
    lo, hi = 0, niter
    if wstart is not None and wstart > start:
        lo = int(math.ceil((wstart - start) / thin - 1e-9))
    if wend is not None and wend < end:
        hi = int(math.floor((wend - start) / thin + 1e-9)) + 1
    return max(lo, 0), min(hi, niter)


def geweke_z(x, frac1=0.1, frac2=0.5):
    """coda::geweke.diag(x)$z for an Mcmc object: z-score per variable comparing the mean of the first `frac1` of the chain with
    the mean of the last `frac2`, variances from spectrum0.ar.  Non-finite where a window has no variation."""
    data = np.asarray(x.data, dtype=np.float64)
    start, end, thin = x.mcpar
    n = data.shape[0]
    if frac1 < 0 or frac1 > 1 or frac2 < 0 or frac2 > 1 or frac1 + frac2 > 1:
        raise ValueError("start and end sequences are overlapping")
    xstart = (start, math.floor(end - frac2 * (end - start)))
    xend = (math.ceil(start + frac1 * (end - start)), end)
    means, variances = [], []
    for i in range(2):
        lo, hi = _window_rows(start, end, thin, n, xstart[i], xend[i])
        y = data[lo:hi]
        means.append(y.mean(axis=0))
        variances.append(np.array([spectrum0_ar(y[:, j])[0] for j in range(y.shape[1])]) / y.shape[0])
    with np.errstate(divide="ignore", invalid="ignore"):
        return (means[0] - means[1]) / np.sqrt(variances[0] + variances[1])


def pcramer(q, eps=1e-5):
    """coda's pcramer: distribution function of the Cramer-von Mises statistic (four terms of the Anderson-Darling series)."""
    from scipy.special import gamma, kv
    q = float(q)
    if not q > 0:
        return 0.0
    log_eps = math.log(eps)
    total = 0.0
    for kk in range(4):
        z = gamma(kk + 0.5) * math.sqrt(4 * kk + 1) / (gamma(kk + 1) * math.pi ** 1.5 * math.sqrt(q))
        u = (4 * kk + 1) ** 2 / (16 * q)
        total += 0.0 if u > -log_eps else z * math.exp(-u) * kv(0.25, u)
    return total


def heidel_diag(x, eps=0.1, pvalue=0.05):
    """coda::heidel.diag(x): per variable (stest, start, pvalue, htest, mean, halfwidth); NaN plays R's NA."""
    data = np.asarray(x.data, dtype=np.float64)
    start, end, thin = x.mcpar
    n_all = data.shape[0]
    out = np.full((data.shape[1], 6), np.nan)
    for j in range(data.shape[1]):
        y_all = data[:, j]
        starts = np.arange(start, end / 2 + 1e-9, n_all / 10.0)      # seq(from = start, to = end / 2, by = niter / 10)
        lo2, _ = _window_rows(start, end, thin, n_all, end / 2)
        s0 = spectrum0_ar(y_all[lo2:])[0]
        converged, I, lo = False, float("nan"), 0
        y = y_all
        for st in starts:
            lo, _ = _window_rows(start, end, thin, n_all, st)
            y = y_all[lo:]
            n = y.size
            ybar = y.mean()
            B = np.cumsum(y) - ybar * np.arange(1, n + 1)
            with np.errstate(divide="ignore", invalid="ignore"):
                I = float(np.sum(B * B / (n * s0)) / n)
            converged = bool(np.isfinite(I) and pcramer(I) < 1 - pvalue)
            if converged:
                break
        n = y.size
        ybar = y.mean()
        s0ci = spectrum0_ar(y)[0]
        half = 1.96 * math.sqrt(s0ci / n) if s0ci >= 0 else float("nan")
        with np.errstate(divide="ignore", invalid="ignore"):
            passed = bool(np.isfinite(half) and abs(half / ybar) <= eps)
        pv = 1 - pcramer(I) if np.isfinite(I) else float("nan")
        if (not converged) or not np.isfinite(I) or not np.isfinite(half):
            out[j] = [1.0 if converged else 0.0, np.nan, pv, np.nan, np.nan, np.nan]
        else:
            out[j] = [1.0, start + lo * thin, pv, 1.0 if passed else 0.0, ybar, half]
    return out
