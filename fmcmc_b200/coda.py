"""Minimal stand-ins for coda's `mcmc` / `mcmc.list` containers (R/mcmc.R:829-836, 641, 671):
a numeric matrix with iteration row names, parameter column names and mcpar = (start, end, thin)."""
from __future__ import annotations

import numpy as np


class Mcmc:
    def __init__(self, data, start=1, end=None, thin=1, varnames=None):
        self.data = np.asarray(data, dtype=np.float64)
        if self.data.ndim == 1:
            self.data = self.data.reshape(-1, 1)
        n = self.data.shape[0]
        self.start = int(start)
        self.thin = int(thin)
        self.end = int(end) if end is not None else self.start + (n - 1) * self.thin
        self.varnames = list(varnames) if varnames is not None else [f"par{i + 1}" for i in range(self.data.shape[1])]

    @classmethod
    def _view(cls, data, start, end, thin, varnames):
        """A chain that is a VIEW of one slab of a [chain][row][param] array (no copy, shared names): what an mcmc.list of
        thousands of chains is made of."""
        self = object.__new__(cls)
        self.data, self.start, self.end, self.thin, self.varnames = data, start, end, thin, varnames
        return self

    # coda accessors
    @property
    def mcpar(self):
        return (self.start, self.end, self.thin)

    def niter(self):
        return self.data.shape[0]

    def nvar(self):
        return self.data.shape[1]

    def nchain(self):
        return 1

    def iterations(self):
        return self.start + self.thin * np.arange(self.niter())

    def __getitem__(self, idx):
        sub = self.data[idx]
        if isinstance(idx, tuple) and len(idx) == 2 and not np.isscalar(idx[1]) and sub.ndim == 2 \
                and sub.shape[0] == self.niter():
            cols = np.arange(self.nvar())[idx[1]]
            return Mcmc(sub, self.start, self.end, self.thin, [self.varnames[c] for c in np.atleast_1d(cols)])
        return sub

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)

    def mean(self, axis=0):
        return self.data.mean(axis=axis)

    def __repr__(self):
        return f"Mcmc(niter={self.niter()}, nvar={self.nvar()}, mcpar={self.mcpar})"


class McmcList(list):
    _arr = None          # [nchains][niter][nvar] when the chains are views of one array (from_array)
    _meta = None         # (start, end, thin, varnames) of an array-backed list
    _lazy = False        # array-backed and the per-chain Mcmc views not built yet (they are built on first element access)

    @classmethod
    def from_array(cls, arr, start=1, end=None, thin=1, varnames=None):
        """mcmc.list over one [nchains][niter][nvar] array: every chain is a view, as_array() / select() / append_chains
        work on the whole block at once (65 536 chains are one array, not 65 536 objects' worth of copies).  The per-chain
        views themselves are only created when an element is asked for: a bulk loop that never looks at them (the device
        checker reads the sample store) does not pay 1 024 object constructions per bulk."""
        arr = np.asarray(arr, dtype=np.float64)
        start, thin = int(start), int(thin)
        end = int(end) if end is not None else start + (arr.shape[1] - 1) * thin
        names = list(varnames) if varnames is not None else [f"par{i + 1}" for i in range(arr.shape[2])]
        self = cls()
        self._arr, self._meta, self._lazy = arr, (start, end, thin, names), True
        return self

    def _fill(self):
        if self._lazy:
            self._lazy = False
            start, end, thin, names = self._meta
            list.extend(self, (Mcmc._view(self._arr[c], start, end, thin, names) for c in range(self._arr.shape[0])))

    def __len__(self):
        return self._arr.shape[0] if self._lazy else list.__len__(self)

    def __getitem__(self, idx):
        self._fill()
        return list.__getitem__(self, idx)

    def __iter__(self):
        self._fill()
        return list.__iter__(self)

    def __bool__(self):
        return len(self) > 0

    def __eq__(self, other):
        self._fill()
        if isinstance(other, McmcList):
            other._fill()
        return list.__eq__(self, other)

    __hash__ = None

    def nchain(self):
        return len(self)

    def niter(self):
        return self._arr.shape[1] if self._meta is not None else self[0].niter()

    def nvar(self):
        return self._arr.shape[2] if self._meta is not None else self[0].nvar()

    @property
    def mcpar(self):
        return self._meta[:3] if self._meta is not None else self[0].mcpar

    @property
    def varnames(self):
        return self._meta[3] if self._meta is not None else self[0].varnames

    def as_array(self):
        """[nchains][niter][nvar]"""
        return self._arr if self._arr is not None else np.stack([m.data for m in self])

    def select(self, cols):
        if self._arr is not None:
            cols = np.atleast_1d(np.arange(self.nvar())[cols])
            return McmcList.from_array(self._arr[:, :, cols], *self.mcpar, [self.varnames[c] for c in cols])
        return McmcList([m[:, cols] for m in self])

    def __repr__(self):
        return f"McmcList(nchain={len(self)}, niter={self.niter()}, nvar={self.nvar()}, mcpar={self.mcpar})"


def append_chains(*objs):
    """R/append_chains.R:44-145: rbind consecutive runs, renumber iterations, rebuild mcpar."""
    objs = [o for o in objs if o is not None and len(o) > 0]
    if not objs:
        raise ValueError("No method available to append these chains.")
    if len(objs) == 1:
        return objs[0]
    if isinstance(objs[0], McmcList):
        nch = {o.nchain() for o in objs}
        if len(nch) != 1:
            raise ValueError("All mcmc.list objects must have the same number of chains. The passed objects have "
                             + ", ".join(str(o.nchain()) for o in objs) + " respectively.")
        if all(o._arr is not None for o in objs) and len({o._arr.shape[2] for o in objs}) == 1 \
                and len({o[0].thin for o in objs}) == 1:
            start, end, thin = append_mcpar([o.mcpar for o in objs])
            return McmcList.from_array(np.concatenate([o._arr for o in objs], axis=1), start, end, thin, objs[0].varnames)
        return McmcList([append_chains(*[o[i] for o in objs]) for i in range(objs[0].nchain())])
    thin = [o.thin for o in objs]
    if len(set(thin)) != 1:
        raise ValueError("All `mcmc` objects have to have the same `thin` parameter.Observed: "
                         + ", ".join(map(str, thin)) + " respectively.")
    nvar = [o.nvar() for o in objs]
    if len(set(nvar)) != 1:
        raise ValueError("All `mcmc` objects have to have the same number of parameters.Observed: "
                         + ", ".join(map(str, nvar)) + " respectively.")
    start = [o.start for o in objs]
    end = [o.end for o in objs]
    for i in range(1, len(objs)):                     # R/append_chains.R:128
        end[i] = end[i] + thin[i] - start[i]
    data = np.concatenate([o.data for o in objs], axis=0)
    return Mcmc(data, start=start[0], end=sum(end), thin=thin[0], varnames=objs[0].varnames)


def append_mcpar(mcpars):
    """mcpar of rbind-ed consecutive runs (R/append_chains.R:113-142): start of the first, the ends renumbered."""
    start = [m[0] for m in mcpars]
    end = [m[1] for m in mcpars]
    thin = mcpars[0][2]
    for i in range(1, len(mcpars)):                   # R/append_chains.R:128
        end[i] = end[i] + thin - start[i]
    return start[0], sum(end), thin


def __len_mcmc(self):
    return self.niter()


Mcmc.__len__ = __len_mcmc


def window_first_row(start: int, end: int, thin: int, niter: int, new_start: float) -> int:
    """0-based first row kept by coda's window.mcmc(x, start = new_start) (start snaps UP to
    the next kept iteration; third-party coda, restated from its published source)."""
    if new_start < start:
        new_start = start
    xtime = start + thin * np.arange(niter)
    ts_eps = 1e-5
    if np.all(np.abs(xtime - new_start) > abs(new_start) * ts_eps):
        cand = xtime[(xtime > new_start) & ((new_start + thin) > xtime)]
        new_start = cand[0]
    first = int(np.trunc((new_start - start) / thin + 1.5))   # 1-based
    return first - 1
