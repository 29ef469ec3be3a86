"""fmcmc_b200 — B200-native drop-in for the multi-chain Metropolis-Hastings path of
USCbiostats/fmcmc: MCMC(), the kernel_*() constructors, convergence_gelman(), cov_recursive /
mean_recursive / reflect_on_boundaries, all backed by hand-written sm_100a CUDA kernels behind
the C ABI in include/fmcmc_b200.h.  There is no CPU path in this package."""
from ._lib import FmcmcError, build, lib  # noqa: F401
from .api import MCMC, FedStream, check_initial  # noqa: F401
from .coda import Mcmc, McmcList, append_chains  # noqa: F401
from .convergence import (LAST_CONV_CHECK, convergence_auto, convergence_data_get, convergence_data_set,  # noqa: F401
                          convergence_gelman, convergence_geweke, convergence_heildel, convergence_msg_get,
                          convergence_msg_set)
from .device import DeviceModel, cov_recursive, mean_recursive, reflect_on_boundaries  # noqa: F401
from .families import DeviceFamily, ll_gaussian_lm, ll_hier_normal, ll_logistic  # noqa: F401
from .kernels import (FmcmcKernel, kernel_adapt, kernel_am, kernel_new, kernel_nmirror,  # noqa: F401
                      kernel_normal, kernel_normal_reflective, kernel_ram, kernel_umirror,
                      kernel_unif, kernel_unif_reflective)
from .mcmc_info import (MCMC_OUTPUT, get_draws, get_elapsed, get_kernel, get_logpost,  # noqa: F401
                        get_nchains, ith_step)

__version__ = "0.1.0"
