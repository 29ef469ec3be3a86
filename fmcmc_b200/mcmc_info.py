"""Run-state side channel of the reference (R/mcmc_info.R:62-372): MCMC_OUTPUT, get_logpost(),
get_draws(), get_elapsed(), ...  Filled from the device buffers after every MCMC() call."""
from __future__ import annotations

import time


class _Output:
    def __init__(self):
        self.clear(0)

    def clear(self, nchains):
        self.nchains = nchains
        self.logpost = [None] * nchains
        self.draws = [None] * nchains
        self.args = {}
        self.t0 = self.t1 = None
        self.kernel = None
        self.report = None
        self.reports = []         # one fmcmc_run_report per bulk (device time, launches, bytes copied)


MCMC_OUTPUT = _Output()


def MCMC_init(**args):
    MCMC_OUTPUT.args = args
    MCMC_OUTPUT.t0 = time.perf_counter()


def MCMC_finalize():
    MCMC_OUTPUT.t1 = time.perf_counter()


def _one_or_list(v):
    return v[0] if len(v) == 1 else list(v)


def get_logpost():
    """Trace of f(proposal) per chain (quirk D1, R/mcmc.R:754,822)."""
    return _one_or_list(MCMC_OUTPUT.logpost)


def get_draws():
    """The proposed states per chain (R/mcmc.R:752,823)."""
    return _one_or_list(MCMC_OUTPUT.draws)


def get_elapsed():
    return (MCMC_OUTPUT.t1 or time.perf_counter()) - MCMC_OUTPUT.t0


def get_nchains():
    return MCMC_OUTPUT.nchains


def get_kernel():
    return MCMC_OUTPUT.kernel


def get_initial():
    return MCMC_OUTPUT.args.get("initial")


def get_nsteps():
    return MCMC_OUTPUT.args.get("nsteps")


def get_seed():
    return MCMC_OUTPUT.args.get("seed")


def get_burnin():
    return MCMC_OUTPUT.args.get("burnin")


def get_thin():
    return MCMC_OUTPUT.args.get("thin")


def ith_step(*a, **k):
    raise RuntimeError("ith_step(): the MCMC loop runs inside a CUDA kernel; there is no R/Python frame "
                       "to inspect (R/mcmc_info.R:573-583 is meaningless without a host closure).")


set_userdata = get_userdata = add_userdata = ith_step
