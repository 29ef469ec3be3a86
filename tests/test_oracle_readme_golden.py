"""Pins the oracle against the outputs the reference itself publishes.

README.md:183-201 (posterior summary, seed 1215) and README.md:315-339 / 388-412 (the
two Gelman-Rubin traces) are regenerated WITHOUT R: oracle/r_rng.c replays R's
Mersenne-Twister/Inversion streams in the serial path's order and
oracle/fmcmc_oracle.c restates the loop, kernel_normal, kernel_normal_reflective,
reflect_on_boundaries, the bulk loop and coda::gelman.diag.
"""
import numpy as np

from fmcmc_b200 import _abi as A
from helpers import r_fed_stream, readme_model

README_MEAN = [3.113, 1.975, 4.093]                      # README.md:192-194
README_SD = [0.17593, 0.10647, 0.07843]
README_Q = [[2.975, 3.029, 3.068, 3.255, 3.354],         # README.md:198-201
            [1.749, 1.907, 1.980, 2.020, 2.145],
            [3.978, 4.070, 4.101, 4.102, 4.226]]
README_GELMAN_NORMAL = [4.5843, 1.1877, 1.4297, 1.1582, 1.3414, 1.2727, 1.4456, 1.3792,
                        1.2069, 1.1789, 1.1208, 1.1196, 1.0792]      # README.md:315-339
README_GELMAN_REFLECTIVE = [3.7891, 1.1257, 1.4696, 1.1313, 1.4384, 1.3696, 1.5243, 1.3720,
                            1.1722, 1.1492, 1.1004, 1.1161, 1.0815]  # README.md:388-412


def _sig(x, n):
    return float(f"{x:.{n}g}")


def test_readme_first_run_summary(oracle, readme_data):
    R = oracle.RRng
    model = readme_model(readme_data)
    R.set_seed(1215)
    T = 5000
    logu, z = r_fed_stream(R, 1, T, 3)
    stream = A.marshal_stream(A.STREAM_FED, logu=logu, z=z)
    ks = dict(type=A.KERNEL_NORMAL, k=3, mu=0.0, scale=1.0)   # kernel_normal() defaults
    out = oracle.run(model, ks, [0, 0, readme_data["sd_y"]], T, stream=stream)
    a = out["ans"][0]
    assert [_sig(v, 4) for v in a.mean(0)] == README_MEAN
    assert [_sig(v, 5 if v >= .1 else 4) for v in a.std(0, ddof=1)] == README_SD
    q = np.quantile(a, [.025, .25, .5, .75, .975], axis=0).T
    assert [[_sig(v, 4) for v in row] for row in q] == README_Q


def _gelman_trace(oracle, d, ks, guard):
    R = oracle.RRng
    model = readme_model(d, guard=guard)
    R.set_seed(1215)
    nchains, freq = 2, 200
    init = np.tile([0, 0, d["sd_y"]], (nchains, 1))
    acc, vals = None, []
    for _ in range(5000 // freq):
        logu, z = r_fed_stream(R, nchains, freq, 3)
        out = oracle.run(model, ks, init, freq, nchains=nchains,
                         stream=A.marshal_stream(A.STREAM_FED, logu=logu, z=z))
        a = out["ans"]
        acc = a if acc is None else np.concatenate([acc, a], axis=1)   # append_chains
        init = a[:, -1, :]                                             # R/mcmc.R:909-911
        end = acc.shape[1]
        w = acc[:, end // 2:, :]                  # window(start = end/2 + 1), autoburnin
        _, mpsrf, rc = oracle.gelman(w)
        assert rc == 0
        vals.append(round(mpsrf, 4))
        if mpsrf < 1.10:
            break
    return vals, acc.shape[1]


def test_readme_gelman_trace_kernel_normal(oracle, readme_data):
    ks = dict(type=A.KERNEL_NORMAL, k=3, mu=0.0, scale=0.05)
    vals, steps = _gelman_trace(oracle, readme_data, ks, guard=True)
    assert vals == README_GELMAN_NORMAL
    assert steps == 2600                                         # README.md:339


def test_readme_gelman_trace_kernel_normal_reflective(oracle, readme_data):
    ks = dict(type=A.KERNEL_NORMAL_REFLECTIVE, k=3, mu=0.0, scale=0.05,
              lb=[-5.0, 0.0, 0.0], ub=5.0)
    vals, steps = _gelman_trace(oracle, readme_data, ks, guard=False)
    assert vals == README_GELMAN_REFLECTIVE
    assert steps == 2600                                         # README.md:412
