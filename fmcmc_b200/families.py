"""Built-in device log-posterior families: the replacement of the R closure `fun`
(north_star: arbitrary closures cannot run on the device).  Each constructor only records
the data; the upload to HBM happens once per MCMC() call (fmcmc_model_create)."""
from __future__ import annotations

import numpy as np

from . import _abi as A


class DeviceFamily:
    """Descriptor of a log-posterior the CUDA kernels know how to evaluate."""

    def __init__(self, family, n, p_x=0, n_groups=0, X=None, y=None, group=None, flags=0,
                 hyper=(0.0, 0.0, 0.0, 0.0), parnames=None):
        self.family, self.n, self.p_x, self.n_groups = family, int(n), int(p_x), int(n_groups)
        self.X, self.y, self.group, self.flags, self.hyper = X, y, group, flags, tuple(hyper)
        self.parnames = parnames

    @property
    def k(self) -> int:
        if self.family == A.FAMILY_GAUSSIAN_LM:
            return self.p_x + (1 if self.flags & A.MODEL_INTERCEPT else 0) + 1
        if self.family == A.FAMILY_LOGISTIC:
            return self.p_x
        return self.n_groups + 1 + (2 if self.flags & A.MODEL_SCALES else 0)

    # ---- the device-resident copy (X / y uploaded and packed once per family object and device) ----
    def device_model(self, device: int = 0):
        from .device import DeviceModel
        cache = self.__dict__.setdefault("_models", {})
        m = cache.get(device)
        if m is None or not m._h:
            m = cache[device] = DeviceModel(self, device=device)
        return m

    def release(self):
        """Free the HBM copy(ies) of this family's data (also happens when the object is collected)."""
        for m in self.__dict__.get("_models", {}).values():
            m.close()
        self.__dict__["_models"] = {}

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def marshal(self) -> A.Marshalled:
        return A.marshal_model(self.family, self.n, self.p_x, self.n_groups, self.X, self.y, self.group,
                               self.flags, self.hyper)


def ll_gaussian_lm(X, y, intercept=True, guard=True) -> DeviceFamily:
    """sum(dnorm(y - (p[1] + X %*% p[2:(k-1)]), sd = p[k], log = TRUE))  — README.md:128-139
    (guard=True returns -Inf instead of a non-finite sum) and README.md:356-360 (guard=False)."""
    X = np.asfortranarray(np.asarray(X, dtype=np.float64).reshape(len(y), -1))
    y = np.ascontiguousarray(y, dtype=np.float64)
    flags = (A.MODEL_INTERCEPT if intercept else 0) | (A.MODEL_GUARD if guard else 0)
    return DeviceFamily(A.FAMILY_GAUSSIAN_LM, len(y), p_x=X.shape[1], X=X, y=y, flags=flags)


def ll_logistic(X, y, prior_sd=2.0) -> DeviceFamily:
    """Bernoulli-logit log-likelihood + N(0, prior_sd^2) prior on every coefficient —
    vignettes/workflow-with-fmcmc.Rmd:35-41 (prior_sd = 2 gives `- sum(beta^2) / 8`)."""
    X = np.asfortranarray(np.asarray(X, dtype=np.float64).reshape(len(y), -1))
    y = np.ascontiguousarray(y, dtype=np.float64)
    return DeviceFamily(A.FAMILY_LOGISTIC, len(y), p_x=X.shape[1], X=X, y=y, hyper=(float(prior_sd), 0, 0, 0))


def ll_hier_normal(y, group, n_groups=None, gamma_bounds=(-1.0, 1.0), estimate_scales=False) -> DeviceFamily:
    """y_i ~ N(theta_g(i), sigma), theta_g ~ N(gamma, tau), gamma ~ U(lo, hi) —
    playground/hierarchical-bayes.Rmd:45-51 (sigma = tau = 1); with estimate_scales the
    parameter vector is (theta_1..theta_G, gamma, sigma, tau)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    group = np.ascontiguousarray(group, dtype=np.int32)
    if n_groups is None:
        n_groups = int(group.max()) + 1
    flags = A.MODEL_SCALES if estimate_scales else 0
    return DeviceFamily(A.FAMILY_HIER_NORMAL, len(y), n_groups=n_groups, y=y, group=group, flags=flags,
                        hyper=(float(gamma_bounds[0]), float(gamma_bounds[1]), 0, 0))
