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


# ---- README.md:203-269: the same seed-1215 stream continues through kernel_normal(scale = .05),
# kernel_ram() and kernel_adapt(), all restarted from the last row of the second run ----------------
README_RAM_ACCEPT = 0.3522705        # README.md:245-246   1 - rejectionRate(ans_RAM)   (= 1761 / 4999)
README_AM_ACCEPT = 0.5365073         # README.md:268-269   1 - rejectionRate(ans_AM)    (= 2682 / 4999)


def _accept_rate(a):
    return float(np.mean(np.any(a[1:] != a[:-1], axis=1)))


def readme_flow(oracle, d):
    """Replays README.md:160-262 up to the point where the kernel_adapt() run starts; returns the RAM
    run and the fed stream / initial state of the AM run."""
    R = oracle.RRng
    model = readme_model(d)
    R.set_seed(1215)
    T = 5000
    logu, z = r_fed_stream(R, 1, T, 3)
    o1 = oracle.run(model, dict(type=A.KERNEL_NORMAL, k=3, mu=0.0, scale=1.0), [0, 0, d["sd_y"]], T,
                    stream=A.marshal_stream(A.STREAM_FED, logu=logu, z=z))
    logu, z = r_fed_stream(R, 1, T, 3)
    o2 = oracle.run(model, dict(type=A.KERNEL_NORMAL, k=3, mu=0.0, scale=0.05), o1["ans"][0, -1], T,
                    stream=A.marshal_stream(A.STREAM_FED, logu=logu, z=z))
    init = o2["ans"][0, -1].copy()                       # MCMC.mcmc: initial = last row (R/mcmc.R:361)
    # kernel_ram(): runif(T) first (R/mcmc.R:726), then qfun(k) = rt(k, k) per step (R/kernel_ram.R:68,124)
    logu_ram = R.log_runif(T)[None]
    U = np.zeros((1, T, 3))
    U[0, 1:] = R.rt((T - 1) * 3, 3).reshape(T - 1, 3)
    ram_spec = dict(type=A.KERNEL_RAM, k=3, mu=0.0, arate=0.234, freq=1, warmup=0, eps=1e-4)
    o3 = oracle.run(model, ram_spec, init, T, stream=A.marshal_stream(A.STREAM_FED, logu=logu_ram, z=U))
    logu_am, z_am = r_fed_stream(R, 1, T, 3)             # kernel_adapt(): mvrnorm -> rnorm(3) per step
    return dict(model=model, init=init, ram=o3, ram_spec=ram_spec, ram_stream=(logu_ram, U),
                am_stream=(logu_am, z_am), T=T)


def test_readme_kernel_ram_acceptance_rate(oracle, readme_data):
    """Pins R's rt / rchisq / rgamma / exp_rand (oracle/r_rng.c) and the RAM restatement: 1 761 accepted
    transitions out of 4 999, after 10 000 earlier rows drawn from the same stream."""
    f = readme_flow(oracle, readme_data)
    acc = _accept_rate(f["ram"]["ans"][0])
    assert round(acc, 7) == README_RAM_ACCEPT
    assert int(round(acc * 4999)) == 1761
    assert int(f["ram"]["istate"][0, 2]) == 0            # nerrors: chol() never failed (nearPD unpinned, unused)


def _envelope_px(lp, fx):
    """Lowest point of the plotted polyline in every pixel column of man/figures/get_-1.png."""
    cols = fx["cols"]
    lo = np.full(cols.size, np.nan)
    px = np.round(fx["x0_col"] + np.arange(1, lp.size + 1) * fx["px_per_iter"]).astype(int)
    ypix = fx["y_ref_row"] + (fx["y_ref"] - lp) * fx["px_per_unit"]
    for i in range(lp.size - 1):
        v = max(ypix[i], ypix[i + 1])
        for j in {px[i], px[i + 1]}:
            jj = j - cols[0]
            if 0 <= jj < cols.size:
                lo[jj] = v if np.isnan(lo[jj]) else max(lo[jj], v)
    return lo


def test_readme_kernel_adapt_what_is_and_is_not_pinned(oracle, readme_data):
    """kernel_adapt draws through MASS::mvrnorm -> eigen() -> LAPACK dsyevr (R/kernel_adapt.R:173-178).
    (1) During the warm-up Sigma = eps*I and R's eigen() returns the EXCHANGE matrix (ascending LAPACK
        order reversed), so the draw is sqrt(eps)*(z3, z2, z1).  The reference publishes the log-posterior
        trace of exactly this run (man/figures/get_-1.png); its first 500 iterations follow the oracle's
        EIGEN mode to ~1 pixel and do NOT follow the natural order (z1, z2, z3).
    (2) After the warm-up the eigenvector signs (and, at the first adapted step, the basis of an exactly
        repeated eigenvalue) are LAPACK-build artefacts: neither the oracle's convention nor OpenBLAS'
        own dsyevr (scipy driver='evr': 2 698 / 4 999) reproduces README's 2 682 / 4 999 - the acceptance
        rate is matched to 0.01 only, and that is the documented contract."""
    import os
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                              "readme_am_logpost_envelope.npz"))
    f = readme_flow(oracle, readme_data)
    logu, z = f["am_stream"]
    res = {}
    for name, mvn in (("eigen", A.MVN_EIGEN), ("cholesky", A.MVN_CHOLESKY)):
        spec = dict(type=A.KERNEL_ADAPT, k=3, mu=0.0, warmup=500, freq=1, eps=1e-4, mvn_method=mvn)
        o = oracle.run(f["model"], spec, f["init"], f["T"], stream=A.marshal_stream(A.STREAM_FED, logu=logu, z=z))
        lo = _envelope_px(o["logpost"][0], fx)
        warm = slice(21, 71)                              # pixel columns 100..149 = iterations 10..490
        err = np.abs(lo[warm] - fx["lo"][warm])
        res[name] = (float(np.median(err)), _accept_rate(o["ans"][0]))
    assert res["eigen"][0] <= 1.5, res                    # the published trace, to ~1 pixel (15.4 px per unit)
    assert res["cholesky"][0] >= 3.0, res                 # natural-order draws are a different path
    for name in res:                                      # the unpinned part: rate only
        assert abs(res[name][1] - README_AM_ACCEPT) < 0.01, res


def test_mvn_eigen_and_cholesky_agree_distributionally(oracle, readme_data):
    """A A' = Sigma for both factors, so both draws are N(mu, Sigma): posterior means within 4 MCSE and
    equal acceptance rates within Monte-Carlo error, on the README model with Philox-free numpy streams."""
    d = readme_data
    model = readme_model(d)
    rng = np.random.default_rng(7)
    C, T = 16, 3000
    out = {}
    for name, mvn in (("eigen", A.MVN_EIGEN), ("cholesky", A.MVN_CHOLESKY)):
        logu = np.log(rng.random((C, T)))
        z = rng.standard_normal((C, T, 3))
        spec = dict(type=A.KERNEL_ADAPT, k=3, mu=0.0, warmup=200, freq=1, eps=1e-4, mvn_method=mvn)
        o = oracle.run(model, spec, [3.0, 2.0, 4.0], T, nchains=C, threads=8,
                       stream=A.marshal_stream(A.STREAM_FED, logu=logu, z=z))
        out[name] = o["ans"][:, 1000:, :]
    for j in range(3):
        m = {n: a[:, :, j].mean(axis=1) for n, a in out.items()}           # per-chain means
        mcse = np.sqrt(m["eigen"].var(ddof=1) / C + m["cholesky"].var(ddof=1) / C)
        assert abs(m["eigen"].mean() - m["cholesky"].mean()) < 4 * mcse, (j, m, mcse)
    acc = {n: np.mean(np.any(a[:, 1:] != a[:, :-1], axis=2)) for n, a in out.items()}
    assert abs(acc["eigen"] - acc["cholesky"]) < 0.02, acc
