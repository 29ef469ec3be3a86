"""Shared helpers for the parity tests: fed streams drawn in R's order."""
import numpy as np

from fmcmc_b200 import _abi as A


def r_fed_stream(R, nchains, T, kdraw, kind="normal"):
    """Serial-path stream order of R/mcmc.R:647-668 + 726: for each chain, runif(T) first,
    then the proposals' draws row by row."""
    logu = np.zeros((nchains, T))
    z = np.zeros((nchains, T, kdraw))
    for c in range(nchains):
        logu[c] = R.log_runif(T)
        if kind == "normal":
            z[c, 1:, :] = R.norm_rand((T - 1) * kdraw).reshape(T - 1, kdraw)
        else:
            z[c, 1:, :] = R.runif((T - 1) * kdraw).reshape(T - 1, kdraw)
    return logu, z


def np_fed_stream(rng, nchains, T, kdraw, kind="normal", df=None):
    logu = np.log(rng.random((nchains, T)))
    if kind == "normal":
        z = rng.standard_normal((nchains, T, kdraw))
    elif kind == "unif":
        z = rng.random((nchains, T, kdraw))
    elif kind == "t":
        z = rng.standard_t(df, size=(nchains, T, kdraw))
    else:
        raise ValueError(kind)
    return logu, z


def readme_model(d, guard=True):
    flags = A.MODEL_INTERCEPT | (A.MODEL_GUARD if guard else 0)
    return A.marshal_model(A.FAMILY_GAUSSIAN_LM, d["n"], p_x=1, X=d["X"], y=d["y"], flags=flags)
