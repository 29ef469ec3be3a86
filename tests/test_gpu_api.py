"""The reference's own test strategy (inst/tinytest/*.R) re-expressed against the CUDA path through the
Python mirror of the fmcmc API: invariants, acceptance bands, error messages, and the README's
convergence traces through MCMC(conv_checker = convergence_gelman())."""
import re

import numpy as np
import pytest

import fmcmc_b200 as fm
from fmcmc_b200 import _abi as A
from helpers import r_fed_stream

pytestmark = pytest.mark.gpu


def _readme_ll(d, guard=True):
    return fm.ll_gaussian_lm(d["X"], d["y"], intercept=True, guard=guard)


@pytest.mark.parametrize("which", ["normal", "reflective"])
def test_readme_gelman_autostop(oracle, readme_data, capfd, which):
    """Config 2 / README.md:298-339 and 372-412: 2 chains, convergence_gelman(200), R's own streams."""
    R = oracle.RRng
    R.set_seed(1215)
    feds = []
    for _ in range(25):
        logu, z = r_fed_stream(R, 2, 200, 3)
        feds.append(fm.FedStream(logu, z))
    if which == "normal":
        kern = fm.kernel_normal(scale=.05)
        want = [4.5843, 1.1877, 1.4297, 1.1582, 1.3414, 1.2727, 1.4456, 1.3792, 1.2069, 1.1789, 1.1208,
                1.1196, 1.0792]
    else:
        kern = fm.kernel_normal_reflective(ub=5.0, lb=[-5.0, 0.0, 0.0], scale=0.05)
        want = [3.7891, 1.1257, 1.4696, 1.1313, 1.4384, 1.3696, 1.5243, 1.3720, 1.1722, 1.1492, 1.1004,
                1.1161, 1.0815]
    with pytest.warns(UserWarning, match="single initial point"):
        ans = fm.MCMC([0, 0, readme_data["sd_y"]], _readme_ll(readme_data, guard=(which == "normal")), 5000,
                      kernel=kern, nchains=2, conv_checker=fm.convergence_gelman(200), fed=feds)
    err = capfd.readouterr().err
    vals = [float(v) for v in re.findall(r"Gelman-Rubin's R: ([0-9.]+)\.", err)]
    assert vals == want
    assert "Convergence has been reached with 2600 steps." in err and "(2600 final count of samples)" in err
    assert isinstance(ans, fm.McmcList) and ans.niter() == 2600 and ans.mcpar == (1, 2600, 1)
    assert kern.is_list and len(kern) == 2
    assert len(fm.get_logpost()) == 2 and fm.get_logpost()[0].shape == (2600,)


def test_same_seed_same_output(readme_data):
    """inst/tinytest/test-mcmc.R:123-137"""
    ll = _readme_ll(readme_data)
    with pytest.warns(UserWarning):
        a = fm.MCMC([1, 1, 4], ll, 500, nchains=3, seed=1231, kernel=fm.kernel_normal(scale=.1))
        b = fm.MCMC([1, 1, 4], ll, 500, nchains=3, seed=1231, kernel=fm.kernel_normal(scale=.1))
        c = fm.MCMC([1, 1, 4], ll, 500, nchains=3, seed=1232, kernel=fm.kernel_normal(scale=.1))
    assert np.array_equal(a.as_array(), b.as_array())
    assert not np.array_equal(a.as_array(), c.as_array())
    assert not np.array_equal(a[0].data, a[1].data)


def test_posterior_mean_band(readme_data):
    """inst/tinytest/test-mcmc.R:43-48 style acceptance band + BASELINE 'within 4 MCSE'."""
    ll = _readme_ll(readme_data)
    ans = fm.MCMC(np.tile([3, 2, 4.0], (64, 1)), ll, 4000, nchains=64, burnin=1000, seed=7,
                  kernel=fm.kernel_normal_reflective(scale=.08, lb=[np.nan, np.nan, 0.0]))
    x = ans.as_array()
    X1 = np.c_[np.ones(readme_data["n"]), readme_data["X"]]
    bhat = np.linalg.lstsq(X1, readme_data["y"], rcond=None)[0]
    chain_means = x.mean(axis=1)
    mcse = chain_means.std(axis=0, ddof=1) / np.sqrt(64)
    assert np.all(np.abs(chain_means.mean(0)[:2] - bhat) < np.maximum(4 * mcse[:2], 0.02))
    assert np.all(x[:, :, 2] > 0)


@pytest.mark.parametrize("kname", ["normal_reflective", "adapt", "ram", "nmirror"])
def test_production_streams_agree_with_reference_streams(oracle, readme_data, kname):
    """north_star correctness (2): production Philox streams on the GPU agree DISTRIBUTIONALLY with the reference's
    own streams (R's Mersenne-Twister + inversion, replayed by the oracle in the serial path's order): posterior
    means within 4 MCSE, posterior sds within 10 %, and the two sets of chains are indistinguishable to
    Gelman-Rubin (R-hat of the pooled set < 1.05)."""
    from fmcmc_b200 import _abi as A
    from helpers import r_fed_stream, readme_model
    ll = _readme_ll(readme_data)
    C, T, burn = 32, 3000, 1000
    lb = [np.nan, np.nan, 0.0]
    kern = {"normal_reflective": lambda: fm.kernel_normal_reflective(scale=.1, lb=lb),
            "adapt": lambda: fm.kernel_adapt(warmup=300, lb=lb),
            "ram": lambda: fm.kernel_ram(lb=lb),
            "nmirror": lambda: fm.kernel_nmirror(mu=[3.0, 2.0, 4.0], scale=.2, warmup=400, lb=lb)}[kname]
    init = np.tile([3.0, 2.0, 4.0], (C, 1))
    g = fm.MCMC(init, ll, T, nchains=C, burnin=burn, seed=11, kernel=kern()).as_array()        # [C][T-burn][3]
    # the reference's streams: R RNG -> fed into the oracle (kernel_ram's U = rt(k, k) is third-party: numpy t)
    R = oracle.RRng
    R.set_seed(4242)
    spec = kern().to_spec(3)                                                                   # a fresh kernel object
    kd = 3
    if kname == "ram":
        rng = np.random.default_rng(4242)
        logu, z = np.log(rng.random((C, T))), rng.standard_t(3, size=(C, T, kd))
    else:
        logu, z = r_fed_stream(R, C, T, kd)
    o = oracle.run(readme_model(readme_data), spec, init, T, nchains=C, burnin=burn,
                   stream=A.marshal_stream(A.STREAM_FED, logu=logu, z=z), threads=8)["ans"]
    mg, mo = g.mean(axis=1), o.mean(axis=1)                                                     # per-chain means
    mcse = np.sqrt(mg.var(axis=0, ddof=1) / C + mo.var(axis=0, ddof=1) / C)
    assert np.all(np.abs(mg.mean(0) - mo.mean(0)) < 4 * mcse), (mg.mean(0), mo.mean(0), mcse)
    sg, so = g.reshape(-1, 3).std(axis=0, ddof=1), o.reshape(-1, 3).std(axis=0, ddof=1)
    assert np.all(np.abs(sg / so - 1) < 0.10), (sg, so)
    pooled = np.concatenate([g, o], axis=0)
    psrf, mpsrf, rc = oracle.gelman(pooled)
    assert rc == 0 and mpsrf < 1.05 and np.all(psrf < 1.05), (psrf, mpsrf)


def test_fixed_and_bounds(readme_data):
    """test-mcmc.R:155-162 (fixed column constant), test-na-bounds.R:63-93 (NA == +-xmax, samples in range)."""
    ll = _readme_ll(readme_data)
    a = fm.MCMC([1, 2, 4.0], ll, 600, seed=3, kernel=fm.kernel_normal_reflective(
        scale=.3, lb=[np.nan, np.nan, 3.5], ub=[np.nan, np.nan, 4.5], fixed=[False, True, False]))
    assert np.all(a.data[:, 1] == 2.0)
    assert a.data[:, 2].min() >= 3.5 and a.data[:, 2].max() <= 4.5
    b = fm.MCMC([1, 2, 4.0], ll, 600, seed=3, kernel=fm.kernel_normal_reflective(
        scale=.3, lb=[-A.DBL_MAX, -A.DBL_MAX, 3.5], ub=[A.DBL_MAX, A.DBL_MAX, 4.5], fixed=[False, True, False]))
    assert np.array_equal(a.data, b.data)
    d = fm.get_draws()
    assert d[:, 2].min() >= 3.5 and d[:, 2].max() <= 4.5


def test_restart_from_mcmc_list_and_kernel_reuse(readme_data):
    """inst/tinytest/test-mcmc.R:252-267 (initial = a previous mcmc.list) and the workflow vignette's kernel reuse
    (vignettes/workflow-with-fmcmc.Rmd:162-252): the second run starts from the last row of the first one, and the
    adaptive kernel's state (abs_iter, Sigma, Mean_t_prev) carries over through the kernel object."""
    ll = _readme_ll(readme_data)
    kern = fm.kernel_adapt(warmup=100, lb=[np.nan, np.nan, 0.0])
    a = fm.MCMC(np.tile([3.0, 2.0, 4.0], (3, 1)), ll, 400, nchains=3, seed=5, kernel=kern)
    assert kern.is_list and [kc.abs_iter for kc in kern] == [399] * 3
    sig1 = [kc.Sigma.copy() for kc in kern]
    b = fm.MCMC(a, ll, 300, nchains=3, seed=6, kernel=kern)            # restart: initial = the mcmc.list
    assert [kc.abs_iter for kc in kern] == [399 + 299] * 3
    for c in range(3):
        assert np.array_equal(b[c].data[0], a[c].data[-1])              # first row = where the last run stopped
        assert not np.array_equal(kern[c].Sigma, sig1[c])               # still adapting, from the carried state
        assert np.all(np.linalg.eigvalsh(kern[c].Sigma) > 0)
    x = b.as_array()
    X1 = np.c_[np.ones(readme_data["n"]), readme_data["X"]]
    bhat = np.linalg.lstsq(X1, readme_data["y"], rcond=None)[0]
    assert np.all(np.abs(x[:, :, :2].mean(axis=(0, 1)) - bhat) < 0.25)
    with pytest.raises(ValueError, match="nchains"):
        fm.MCMC(a, ll, 100, nchains=2, kernel=fm.kernel_normal())       # R/mcmc.R:382-392


def test_ordered_scheme_alternates(readme_data):
    """inst/tinytest/test-kernel_normal.R:92-98"""
    ll = _readme_ll(readme_data)
    fm.MCMC([1, 2, 4.0], ll, 50, seed=3, kernel=fm.kernel_normal(scale=.1, scheme="ordered"))
    d = fm.get_draws()
    a = fm.MCMC([1, 2, 4.0], ll, 50, seed=3, kernel=fm.kernel_normal(scale=.1, scheme="ordered"))
    prev = np.vstack([a.data[:1], a.data[:-1]])
    moved = (d != prev)[1:]
    for r, row in enumerate(moved, start=2):
        assert list(np.where(row)[0]) == [(r - 1) % 3]


def test_errors(readme_data):
    """inst/tinytest/test-mcmc.R:4-24, test-kernels.R:14-86, test-convergence.R:57-67"""
    ll = _readme_ll(readme_data, guard=False)
    with pytest.raises(ValueError, match="burnin"):
        fm.MCMC([1, 1, 1], ll, 100, burnin=100)
    with pytest.raises(ValueError, match="thin"):
        fm.MCMC([1, 1, 1], ll, 100, thin=100)
    with pytest.raises(fm.FmcmcError, match="undefined") as ei:      # sd < 0 -> dnorm NaN -> abort
        fm.MCMC([1, 1, 0.05], ll, 200, seed=1, kernel=fm.kernel_normal(mu=[0, 0, -1.0], scale=.01))
    assert ei.value.code == A.ENAN and "step i =" in str(ei.value)
    with pytest.raises(TypeError, match="closure"):
        fm.MCMC([1, 1, 1], lambda p: 0.0, 100)
    with pytest.raises(ValueError, match="-ub- cannot be <= than -lb-."):
        fm.MCMC([1, 1, 1], ll, 100, kernel=fm.kernel_normal_reflective(lb=1.0, ub=0.0))
    with pytest.raises(ValueError, match="cannot be zero"):
        fm.MCMC([1, 1, 1], ll, 100, kernel=fm.kernel_normal(fixed=True))
    with pytest.raises(ValueError, match="only available when `nchains` > 1L"):
        fm.MCMC([1, 1, 4.0], _readme_ll(readme_data), 2000, conv_checker=fm.convergence_gelman(500), seed=1)
    with pytest.raises(TypeError, match="closures cannot run on the device"):
        fm.kernel_new(lambda env: env)


def test_adaptive_kernels_reach_target(readme_data):
    """test-kernel_adapt.R:16-25, test-kernel_ram.R:17-27, test-kernel_mirror.R:18-41: acceptance bands,
    and kernel state written back (vignettes/workflow-with-fmcmc.Rmd:203-252)."""
    ll = _readme_ll(readme_data)
    X1 = np.c_[np.ones(readme_data["n"]), readme_data["X"]]
    bhat = np.linalg.lstsq(X1, readme_data["y"], rcond=None)[0]
    lb = [np.nan, np.nan, 1e-3]
    for kern in (fm.kernel_adapt(lb=lb, warmup=300), fm.kernel_ram(lb=lb),
                 fm.kernel_nmirror(lb=lb, warmup=400, scale=.2), fm.kernel_umirror(lb=lb, warmup=400, scale=.2)):
        with pytest.warns(UserWarning):
            ans = fm.MCMC([3, 2, 4.0], ll, 4000, nchains=8, burnin=2000, seed=11, kernel=kern)
        m = ans.as_array().mean(axis=(0, 1))
        assert np.linalg.norm(m[:2] - bhat) < 0.25, (kern.type, m)
        assert kern[0].abs_iter == 3999
        if kern.type in (A.KERNEL_ADAPT, A.KERNEL_RAM):
            assert kern[0].Sigma.shape == (3, 3) and np.all(np.isfinite(kern[0].Sigma))
    # resume: pass the previous run + the same kernel object (abs_iter keeps counting)
    k2 = fm.kernel_adapt(lb=lb, warmup=100)
    a1 = fm.MCMC([2, 1, 4.0], ll, 500, seed=1, kernel=k2)
    a2 = fm.MCMC(a1, ll, 500, seed=2, kernel=k2)
    assert k2.abs_iter == 998 and np.array_equal(a2.data[0], a1.data[-1])


def test_exported_helpers(oracle):
    """cov_recursive == cov, mean_recursive == colMeans (inst/tinytest/test-kernel_adapt.R:33-55, the
    reference's only KAT) and reflect_on_boundaries' worked examples."""
    rng = np.random.default_rng(1231)
    X = rng.standard_normal((3, 4))
    m = fm.mean_recursive(X[0], X[1:].mean(0), 2)
    np.testing.assert_allclose(m, X.mean(0), rtol=1e-14)
    _, c = fm.cov_recursive(X[0], np.cov(X[1:].T), X[1:].mean(0), 2)
    np.testing.assert_allclose(c, np.cov(X.T), atol=1e-10)
    Y = rng.standard_normal((40, 5))
    mo, co = fm.cov_recursive(Y[2:], np.cov(Y[:2].T), Y[:2].mean(0), 2)
    np.testing.assert_allclose(mo, Y.mean(0), rtol=1e-13)
    np.testing.assert_allclose(co, np.cov(Y.T), atol=1e-10)
    om, oc = oracle.cov_recursive(Y[2:], Y[:2].mean(0), np.cov(Y[:2].T), 2)
    assert np.array_equal(mo, om) and np.array_equal(co, oc)      # same unfused arithmetic
    x = np.array([2.3, 3.7, 5.2, -0.4, -1.6, -3.1, 0.7])
    got = fm.reflect_on_boundaries(x, 0.0, 1.5)
    np.testing.assert_allclose(got, [0.7, 0.7, 0.8, 0.4, 1.4, 0.1, 0.7], atol=1e-15)
    assert np.array_equal(got, oracle.reflect(x, 0.0, 1.5))
    assert fm.reflect_on_boundaries([-0.2], 0.0, A.DBL_MAX)[0] == 0.2
    big = rng.normal(0, 50, (1000, 3))
    lo, hi = np.array([-1.0, 0.0, -A.DBL_MAX]), np.array([2.0, 0.25, 3.0])
    g = fm.reflect_on_boundaries(big, lo, hi)
    o = np.stack([oracle.reflect(r, lo, hi) for r in big])
    np.testing.assert_allclose(g, o, rtol=1e-13, atol=1e-13)
    assert np.all(g >= lo) and np.all(g <= hi)


def test_single_chain_checkers_through_mcmc(readme_data, capfd):
    """inst/tinytest/test-convergence.R:19-121: auto / geweke / heidel auto-stop runs are reproducible for equal seeds; the
    single-chain checkers refuse more than one chain, Gelman refuses one; convergence_auto picks by the number of chains."""
    ll = _readme_ll(readme_data)
    init = [3.0, 2.0, 4.0]

    def go(chk, nchains=1, seed=31):
        return fm.MCMC(init if nchains == 1 else np.tile(init, (nchains, 1)), ll, 3000, nchains=nchains, seed=seed,
                       kernel=fm.kernel_normal_reflective(scale=.1, lb=[np.nan, np.nan, 0.0]), conv_checker=chk)

    for make in (lambda: fm.convergence_geweke(500), lambda: fm.convergence_heildel(500), lambda: fm.convergence_auto(500)):
        a, b = go(make()), go(make())
        assert isinstance(a, fm.Mcmc) and a.niter() % 500 == 0 and np.array_equal(a.data, b.data)
    err = capfd.readouterr().err
    assert "avg Geweke's Z:" in err and "Heidel's Avg. pval:" in err
    two = go(fm.convergence_auto(500), nchains=2)                 # auto with 2 chains = Gelman-Rubin on the device
    assert isinstance(two, fm.McmcList) and "Gelman-Rubin's R:" in capfd.readouterr().err
    with pytest.raises(ValueError, match="single chain"):
        go(fm.convergence_heildel(500), nchains=2)
    with pytest.raises(ValueError, match="single chain"):
        go(fm.convergence_geweke(500), nchains=2)
    with pytest.raises(ValueError, match="nchains` > 1L"):
        go(fm.convergence_gelman(500), nchains=1)
