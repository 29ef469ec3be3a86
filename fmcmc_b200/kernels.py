"""Transition-kernel constructors with the reference's names, arguments and defaults
(R/kernel_normal.R, R/kernel_unif.R, R/kernel_adapt.R, R/kernel_ram.R, R/kernel_mirror.R).

An `FmcmcKernel` plays the role of the R `fmcmc_kernel` environment (R/kernel.R:283-317):
hyper-parameters and mutable per-chain state (`abs_iter`, `Sigma`, `Mean_t_prev`, `mu`,
`scale`, `obs_arate`, `nerrors`) are attributes, readable after the run and reused when the
object is passed to another MCMC() call.  With nchains > 1 the object turns into a kernel
list (R/kernel.R:348-377 rep_kernel): `kernel[i]` is chain i's copy."""
from __future__ import annotations

import copy
import math

import numpy as np

from . import _abi as A

_SCHEMES = {"joint": A.SCHEME_JOINT, "ordered": A.SCHEME_ORDERED, "random": A.SCHEME_RANDOM}


def _check_dimensions(x, k, name):
    """R/kernel.R:1-15"""
    x = np.atleast_1d(np.asarray(x))
    if x.size > 1 and x.size != k:
        raise ValueError(f"Incorrect length of -{name}-.")
    if x.size == 1 and k > 1:
        return np.repeat(x, k)
    return x


def _process_bounds(b, lower):
    """R/kernel.R:25-41: NA -> -/+ .Machine$double.xmax"""
    b = np.array(b, dtype=np.float64, copy=True)
    b[np.isnan(b)] = -A.DBL_MAX if lower else A.DBL_MAX
    return b


_STATE_FIELDS = ("abs_iter", "nerrors", "Sigma", "Mean_t_prev", "mu", "scale", "obs_arate")
_MISSING = object()


class _ChainKernel:
    """One chain's kernel of a replicated kernel list: see FmcmcKernel._replicate."""
    __slots__ = ("_parent", "_c")

    def __init__(self, parent, c):
        object.__setattr__(self, "_parent", parent)
        object.__setattr__(self, "_c", c)

    def __getattr__(self, name):
        p, c = self._parent, self._c
        v = p._overrides.get((c, name), _MISSING)
        if v is not _MISSING:
            return v
        v = p._chain_state(c, name)
        if v is not _MISSING:
            return v
        return getattr(p._proto, name)

    def __setattr__(self, name, value):
        self._parent._overrides[(self._c, name)] = value

    def __repr__(self):
        return f"An environment of class fmcmc_kernel (chain {self._c + 1} of {len(self._parent)})"


class _ChainList:
    """The list a replicated kernel is: chain kernels made on demand."""

    def __init__(self, parent, n):
        self._parent, self._n, self._made = parent, n, {}

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError("kernel list index out of range")
        kc = self._made.get(i)
        if kc is None:
            kc = self._made[i] = _ChainKernel(self._parent, i)
        return kc

    def __iter__(self):
        return (self[i] for i in range(self._n))


class FmcmcKernel:
    def __init__(self, ktype, **hyper):
        self.type = ktype
        self.__dict__.update(hyper)
        self.k = None
        self.abs_iter = 0
        self.nerrors = 0
        self._chains = None      # list of per-chain kernels once replicated (_ChainList of _ChainKernel views)
        self._proto = None       # the environment every chain's kernel was copied from (hyper-parameters)
        self._overrides = {}     # (chain, attribute) -> value assigned to a single chain's kernel
        self._k_total = None     # length of the parameter vector (set by to_spec)
        self._istate = None      # [C][4] int64 / [C][dlen] float64 round-tripped through the ABI
        self._dstate = None
        self._kf = None

    # -- kernel-list behaviour (R/kernel.R:330-434) --------------------------------------
    @property
    def is_list(self):
        return self._chains is not None

    def __len__(self):
        return len(self._chains) if self._chains is not None else 1

    def __getitem__(self, i):
        if self._chains is None:
            raise TypeError("not a kernel list")
        return self._chains[i]

    def _chain_state(self, c, name):
        """The live value of a state field of chain c (what absorb_state writes into a single kernel), or _MISSING."""
        ist, dst = self._istate, self._dstate
        if name not in _STATE_FIELDS or ist is None or c >= ist.shape[0]:
            return _MISSING
        if name == "abs_iter":
            return int(ist[c, 0])
        if name == "nerrors":
            return int(ist[c, 2])
        kf, k, t = self._kf, self._k_total, self.type
        if t in (A.KERNEL_ADAPT, A.KERNEL_RAM) and name == "Sigma":
            return dst[c, :kf * kf].reshape(kf, kf, order="F") if ist[c, 1] & A.STATE_INIT else _MISSING
        if t == A.KERNEL_ADAPT and name == "Mean_t_prev":
            return dst[c, kf * kf:kf * kf + kf] if ist[c, 1] & A.STATE_HAS_MEAN else None
        if t in (A.KERNEL_NMIRROR, A.KERNEL_UMIRROR) and k is not None and ist[c, 1] & A.STATE_INIT:
            if name == "mu":
                return dst[c, :k]
            if name == "scale":
                return dst[c, k:2 * k]
            if name == "obs_arate":
                kind = (int(ist[c, 1]) >> A.STATE_OBS_SHIFT) & 3
                return None if kind == 0 else (float(dst[c, 2 * k]) if kind == 1 else dst[c, 2 * k:3 * k])
        return _MISSING

    def _replicate(self, nchains):
        """rep_kernel (R/kernel.R:407-434): one kernel per chain, each a copy of this one's environment.  The copies are
        VIEWS (_ChainKernel): hyper-parameters come from one deep copy of this object, the state fields (abs_iter, Sigma,
        Mean_t_prev, mu, scale, obs_arate, nerrors) are read live from the state arrays, and assigning an attribute of a chain's
        kernel overrides it for that chain only - 65 536 chains cost no 65 536 deep copies (1 024: 6 ms of a 33 ms call)."""
        proto = copy.copy(self)
        proto._chains = None
        proto._istate = proto._dstate = None
        self._proto = copy.deepcopy(proto)
        self._overrides = {}
        self._chains = _ChainList(self, int(nchains))

    # -- marshalling -------------------------------------------------------------------------
    def to_spec(self, k: int) -> dict:
        """Recycles / validates the hyper-parameters for a k-vector (the lazy init every R
        proposal closure does on its first call) and returns the dict _abi.marshal_kernel eats."""
        t = self.type
        s = dict(type=t, k=k)
        g = self.__dict__
        fixed = _check_dimensions(g.get("fixed", False), k, "fixed").astype(bool)
        s["fixed"] = fixed
        if "lb" in g:
            lb = _process_bounds(_check_dimensions(g["lb"], k, "lb"), True)
            ub = _process_bounds(_check_dimensions(g["ub"], k, "ub"), False)
            if np.any(ub <= lb):
                raise ValueError("-ub- cannot be <= than -lb-.")
            s["lb"], s["ub"] = lb, ub
        if "mu" in g:
            s["mu"] = _check_dimensions(g["mu"], k, "mu").astype(np.float64)
        if "scale" in g:
            s["scale"] = _check_dimensions(g["scale"], k, "scale").astype(np.float64)
        if "min_" in g:
            s["min_"] = _check_dimensions(g["min_"], k, "min.").astype(np.float64)
            s["max_"] = _check_dimensions(g["max_"], k, "max.").astype(np.float64)
            if np.any(s["max_"] <= s["min_"]):
                raise ValueError("-max.- cannot be <= than -min.-.")
        scheme = g.get("scheme", "joint")
        if isinstance(scheme, str):
            if scheme not in _SCHEMES:
                raise ValueError("-scheme- update must be either an integer sequence, 'joint', 'ordered', "
                                 "or 'random'.")
            s["scheme"] = _SCHEMES[scheme]
        else:                                          # explicit integer order, R/kernel.R:69-91
            order = np.asarray(scheme, dtype=np.int32)
            nfree = int((~fixed).sum())
            if order.size != nfree:
                raise ValueError("When setting the update scheme, it should have the same length as the "
                                 f"number of variables that will not be fixed. Right now length(scheme) = "
                                 f"{order.size} while sum(!fixed) = {nfree}.")
            missing = [j + 1 for j in np.where(~fixed)[0] if (j + 1) not in order]
            if missing:
                raise ValueError("One or more variables was not included in the ordering sequence. The full "
                                 f"list follows: {', '.join(map(str, missing))}. Only variables that are not "
                                 "fixed can be included in this list.")
            s["scheme"], s["order"] = A.SCHEME_EXPLICIT, order
        if (~fixed).sum() == 0:
            raise ValueError("The number of parameters to update, i.e. not fixed, cannot be zero. Check the "
                             "value -fixed- in the kernel initialization.")
        for name in ("warmup", "freq", "bw", "until", "eps", "Sd", "arate", "constr", "mvn_method"):
            if name in g and g[name] is not None:
                s[name] = g[name]
        if "nadapt_schedule" in g:
            s["nadapt"] = g["nadapt_schedule"]
        self.k = int((~fixed).sum()) if t in (A.KERNEL_ADAPT, A.KERNEL_RAM) else \
            (int((~fixed).sum()) if s["scheme"] == A.SCHEME_JOINT else 1)
        self._kf = int((~fixed).sum())
        self._k_total = int(k)
        self.which_ = np.where(~fixed)[0] + 1
        return s

    # -- state <-> flat ABI arrays -----------------------------------------------------------
    def state_arrays(self, nchains, k):
        kf = self._kf
        dlen = A.state_len(self.type, k, kf)
        if self._istate is not None and self._istate.shape[0] == nchains and \
                (self._dstate is None or self._dstate.shape[1] == max(dlen, 1)):
            return self._istate, self._dstate
        ist = np.zeros((nchains, A.ISTATE_LEN), dtype=np.int64)
        dst = np.zeros((nchains, max(dlen, 1)), dtype=np.float64)
        if self.type in (A.KERNEL_ADAPT, A.KERNEL_RAM):
            # a user-supplied Sigma seeds EVERY chain: rep_kernel copies the whole environment, Sigma included
            # (R/kernel.R:407-434), and each chain's copy may have been edited since (kernel[[i]]$Sigma <- ...)
            def seed(c, user_sigma):
                sig = np.asarray(user_sigma, dtype=np.float64)
                if sig.shape != (kf, kf):
                    raise ValueError(f"-Sigma- must be a {kf} x {kf} matrix (one row per non-fixed parameter), "
                                     f"got {sig.shape}.")
                dst[c, :kf * kf] = sig.reshape(-1, order="F")
                ist[c, 1] |= A.STATE_INIT
            common = getattr(self._proto if self._chains is not None else self, "Sigma", None)
            if common is not None:
                for c in range(nchains):
                    seed(c, common)
            if self._chains is not None:                       # kernel[[i]]$Sigma <- ... on single chains
                for (c, name), v in self._overrides.items():
                    if name == "Sigma" and v is not None and c < nchains:
                        seed(c, v)
        self._istate, self._dstate = ist, dst
        return ist, dst

    def load_state(self, istate, dstate, nchains, k):
        """Restore a saved per-chain state (what a previous MCMC() call left in this object - abs_iter, Sigma,
        Mean_t_prev, mu / scale / obs_arate): the next MCMC() continues the adaptation from it."""
        self.to_spec(k)
        if nchains > 1 and not self.is_list:
            self._replicate(nchains)
        self._istate = np.ascontiguousarray(istate, dtype=np.int64).reshape(nchains, A.ISTATE_LEN).copy()
        dlen = A.state_len(self.type, k, self._kf)
        self._dstate = np.ascontiguousarray(dstate, dtype=np.float64).reshape(nchains, max(dlen, 1)).copy()
        self.absorb_state(k)

    def absorb_state(self, k):
        """Write the device state back into user-visible attributes (R/mcmc.R:629-631)."""
        ist, dst = self._istate, self._dstate
        if ist is None:
            return
        kf = self._kf
        if self._chains is not None:
            # the chains' kernels read the state arrays live (_ChainKernel): only what a user assigned to a state field before
            # the run gives way to the state the run left
            for key in [q for q in self._overrides if q[1] in _STATE_FIELDS]:
                del self._overrides[key]
            return
        targets = [self]
        for c, kc in enumerate(targets):
            kc.abs_iter = int(ist[c, 0])
            kc.nerrors = int(ist[c, 2])
            # views of the state arrays, not copies: like the fields of the R environment they show the live state, and a
            # run over 65 536 chains does not duplicate 65 536 covariance matrices to expose them
            if self.type == A.KERNEL_ADAPT:
                kc.Sigma = dst[c, :kf * kf].reshape(kf, kf, order="F")
                kc.Mean_t_prev = dst[c, kf * kf:kf * kf + kf] if ist[c, 1] & A.STATE_HAS_MEAN else None
            elif self.type == A.KERNEL_RAM:
                kc.Sigma = dst[c, :kf * kf].reshape(kf, kf, order="F")
            elif self.type in (A.KERNEL_NMIRROR, A.KERNEL_UMIRROR):
                kc.mu = dst[c, :k]
                kc.scale = dst[c, k:2 * k]
                kind = (int(ist[c, 1]) >> A.STATE_OBS_SHIFT) & 3
                kc.obs_arate = None if kind == 0 else (float(dst[c, 2 * k]) if kind == 1 else dst[c, 2 * k:3 * k])

    def __repr__(self):
        if self.is_list:
            return f"A list of {len(self)} fmcmc_kernels."
        names = sorted(n for n in self.__dict__ if not n.startswith("_"))
        return "An environment of class fmcmc_kernel: " + ", ".join(names)


def kernel_normal(mu=0.0, scale=1.0, fixed=False, scheme="joint"):
    """R/kernel_normal.R:26-82"""
    return FmcmcKernel(A.KERNEL_NORMAL, mu=mu, scale=scale, fixed=fixed, scheme=scheme)


def kernel_normal_reflective(mu=0.0, scale=1.0, lb=-A.DBL_MAX, ub=A.DBL_MAX, fixed=False, scheme="joint"):
    """R/kernel_normal.R:96-177"""
    return FmcmcKernel(A.KERNEL_NORMAL_REFLECTIVE, mu=mu, scale=scale, lb=lb, ub=ub, fixed=fixed, scheme=scheme)


def kernel_unif(min_=-1.0, max_=1.0, fixed=False, scheme="joint"):
    """R/kernel_unif.R:15-66 (R's `min.`/`max.` are spelled min_/max_)"""
    return FmcmcKernel(A.KERNEL_UNIF, min_=min_, max_=max_, fixed=fixed, scheme=scheme)


def kernel_unif_reflective(min_=-1.0, max_=1.0, lb=None, ub=None, fixed=False, scheme="joint"):
    """R/kernel_unif.R:74-147 (lb = min., ub = max. by default)"""
    return FmcmcKernel(A.KERNEL_UNIF_REFLECTIVE, min_=min_, max_=max_, lb=min_ if lb is None else lb,
                       ub=max_ if ub is None else ub, fixed=fixed, scheme=scheme)


def kernel_adapt(mu=0.0, bw=0, lb=-A.DBL_MAX, ub=A.DBL_MAX, freq=1, warmup=500, Sigma=None, Sd=None,
                 eps=1e-4, fixed=False, until=math.inf, *, mvn="cholesky"):
    """R/kernel_adapt.R:54-208 (Haario et al. 2001).  `mvn` (not in the reference) picks the factor the normal draw
    goes through: "cholesky" (default) or "eigen", MASS::mvrnorm's own V sqrt(ev) z with R's eigenvalue order - same
    N(mu, Sigma) either way."""
    if bw > 0 and bw > warmup:
        raise ValueError("The `warmup` parameter must be greater than `bw`.")
    if mvn not in ("cholesky", "eigen"):
        raise ValueError("-mvn- must be 'cholesky' or 'eigen'.")
    return FmcmcKernel(A.KERNEL_ADAPT, mu=mu, bw=bw, lb=lb, ub=ub, freq=freq, warmup=warmup, Sigma=Sigma, Sd=Sd,
                       eps=eps, fixed=fixed, until=until, Mean_t_prev=None,
                       mvn_method=A.MVN_EIGEN if mvn == "eigen" else A.MVN_CHOLESKY)


kernel_am = kernel_adapt


def kernel_ram(mu=0.0, eta=None, qfun=None, arate=0.234, freq=1, warmup=0, Sigma=None, eps=1e-4,
               lb=-A.DBL_MAX, ub=A.DBL_MAX, fixed=False, until=math.inf, constr=None):
    """R/kernel_ram.R:65-181 (Vihola 2012).  `eta` and `qfun` are R closures in the reference;
    only their defaults (min(1, i^(-2/3) k) and rt(k, k)) exist on the device."""
    if eta is not None or qfun is not None:
        raise TypeError("kernel_ram: custom `eta` / `qfun` closures cannot run on the device; only the "
                        "reference defaults are built in.")
    return FmcmcKernel(A.KERNEL_RAM, mu=mu, arate=arate, freq=freq, warmup=warmup, Sigma=Sigma, eps=eps, lb=lb,
                       ub=ub, fixed=fixed, until=until, constr=constr)


def _nadapt_schedule(warmup, nadapt):
    """floor(seq(1, warmup, length.out = nadapt + 1)[-1])  (R/kernel_mirror.R:165)"""
    n = nadapt + 1
    if n < 2:
        return np.zeros(0, dtype=np.int64)
    by = (warmup - 1) / (n - 1)
    seq = [1.0] + [1.0 + i * by for i in range(1, n - 1)] + [float(warmup)]
    return np.floor(np.array(seq[1:])).astype(np.int64)


def kernel_nmirror(mu=0.0, scale=1.0, warmup=500, nadapt=4, arate=0.4, lb=-A.DBL_MAX, ub=A.DBL_MAX,
                   fixed=False, scheme="joint"):
    """R/kernel_mirror.R:54-173 (Thawornwattana et al. 2018)."""
    return FmcmcKernel(A.KERNEL_NMIRROR, mu=mu, scale=scale, warmup=warmup, nadapt=nadapt, arate=arate, lb=lb,
                       ub=ub, fixed=fixed, scheme=scheme, obs_arate=None,
                       nadapt_schedule=_nadapt_schedule(warmup, nadapt))


def kernel_umirror(mu=0.0, scale=1.0, warmup=500, nadapt=4, arate=0.4, lb=-A.DBL_MAX, ub=A.DBL_MAX,
                   fixed=False, scheme="joint"):
    """R/kernel_mirror.R:177-301"""
    return FmcmcKernel(A.KERNEL_UMIRROR, mu=mu, scale=scale, warmup=warmup, nadapt=nadapt, arate=arate, lb=lb,
                       ub=ub, fixed=fixed, scheme=scheme, obs_arate=None,
                       nadapt_schedule=_nadapt_schedule(warmup, nadapt))


def kernel_new(proposal=None, *args, **kwargs):
    """R/kernel.R:283-317: user closures cannot run on the device (north_star)."""
    raise TypeError("kernel_new(): a kernel made of R/Python closures cannot run on the device. Use one of "
                    "kernel_normal, kernel_normal_reflective, kernel_unif, kernel_unif_reflective, "
                    "kernel_adapt, kernel_ram, kernel_nmirror, kernel_umirror.")
