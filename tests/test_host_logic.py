"""Host-side mirror of the reference interface (no GPU): argument checks with the reference's error
strings, append_chains / mcpar bookkeeping, coda's window rule, kernel constructors."""
import math
import warnings

import numpy as np
import pytest

import fmcmc_b200 as fm
from fmcmc_b200 import _abi as A
from fmcmc_b200.coda import Mcmc, McmcList, append_chains, window_first_row
from fmcmc_b200.kernels import _nadapt_schedule


def fam3():
    return fm.ll_gaussian_lm(np.arange(6.0), np.arange(6.0))


def test_check_initial():
    """inst/tinytest/test-checks.R:31-41, R/checks.R:22-58"""
    with pytest.warns(UserWarning, match="single initial point"):
        a, names = fm.check_initial([1, 2, 3], 2)
    assert a.shape == (2, 3) and names == ["par1", "par2", "par3"]
    a, names = fm.check_initial({"a": 1.0, "b": 2.0}, 1)
    assert names == ["a", "b"]
    with pytest.raises(ValueError, match="must coincide with the number of chains"):
        fm.check_initial(np.zeros((3, 2)), 2)
    with pytest.raises(ValueError, match="length zero"):
        fm.check_initial([], 1)
    m = Mcmc(np.arange(12.0).reshape(4, 3), start=1, thin=1)
    a, _ = fm.check_initial(m, 1)
    assert np.array_equal(a, [[9, 10, 11]])


def test_mcmc_argument_errors_before_the_gpu_is_touched():
    """inst/tinytest/test-mcmc.R:4-24, test-kernels.R:14-86, test-convergence.R:57-67"""
    f = fam3()
    with pytest.raises(ValueError, match="burnin"):
        fm.MCMC([1, 1, 1], f, 100, burnin=100)
    with pytest.raises(ValueError, match="thin"):
        fm.MCMC([1, 1, 1], f, 100, thin=100)
    with pytest.raises(ValueError, match="should be >= 1"):
        fm.MCMC([1, 1, 1], f, 100, thin=0)
    with pytest.raises(TypeError, match="closure"):
        fm.MCMC([1, 1, 1], lambda p: 0.0, 100)
    with pytest.raises(TypeError, match="not present in -fun-"):
        fm.MCMC([1, 1, 1], f, 100, D=3)
    with pytest.raises(ValueError, match="multicore"):
        fm.MCMC([1, 1, 1], f, 100, multicore=True)
    with pytest.raises(ValueError, match="-ub- cannot be <= than -lb-."):
        fm.MCMC([1, 1, 1], f, 100, kernel=fm.kernel_normal_reflective(lb=1.0, ub=0.0))
    with pytest.raises(ValueError, match="-max.- cannot be <= than -min.-."):
        fm.MCMC([1, 1, 1], f, 100, kernel=fm.kernel_unif(min_=1.0, max_=0.0))
    with pytest.raises(ValueError, match="cannot be zero"):
        fm.MCMC([1, 1, 1], f, 100, kernel=fm.kernel_normal(fixed=True))
    with pytest.raises(ValueError, match="Incorrect length of -scale-"):
        fm.MCMC([1, 1, 1], f, 100, kernel=fm.kernel_normal(scale=[1, 2]))
    with pytest.raises(ValueError, match="either an integer"):
        fm.MCMC([1, 1, 1], f, 100, kernel=fm.kernel_normal(scheme="bogus"))
    with pytest.raises(ValueError, match="same length as"):
        fm.MCMC([1, 1, 1], f, 100, kernel=fm.kernel_normal(scheme=[1, 2]))
    with pytest.raises(ValueError, match="not included in"):
        fm.MCMC([1, 1, 1], f, 100, kernel=fm.kernel_normal(scheme=[1, 2, 2]))
    with pytest.raises(ValueError, match="only available when `nchains` > 1L"):
        fm.MCMC([1, 1, 1], f, 100, conv_checker=fm.convergence_gelman(10))
    with pytest.raises(ValueError, match="must equal the number of chains"):
        fm.MCMC(McmcList([Mcmc(np.zeros((3, 3)))] * 2), f, 100, nchains=3)
    with pytest.raises(ValueError, match="must be greater than `bw`"):
        fm.kernel_adapt(bw=600, warmup=500)
    with pytest.raises(TypeError, match="closures cannot run on the device"):
        fm.kernel_new(lambda env: env)
    with pytest.raises(TypeError, match="qfun"):
        fm.kernel_ram(qfun=lambda k: np.zeros(k))
    with pytest.raises(RuntimeError, match="CUDA kernel"):
        fm.ith_step()


def test_kernel_defaults_match_reference():
    """Defaults of R/kernel_*.R constructors."""
    k = fm.kernel_adapt()
    assert (k.warmup, k.freq, k.eps, k.bw, k.until) == (500, 1, 1e-4, 0, math.inf)
    k = fm.kernel_ram()
    assert (k.arate, k.freq, k.warmup, k.eps) == (0.234, 1, 0, 1e-4)
    k = fm.kernel_nmirror()
    assert (k.warmup, k.nadapt, k.arate) == (500, 4, 0.4)
    assert list(k.nadapt_schedule) == [125, 250, 375, 500]            # R/kernel_mirror.R:165
    assert list(_nadapt_schedule(1000, 5)) == [200, 400, 600, 800, 1000]
    s = fm.kernel_unif_reflective(min_=-2.0, max_=3.0).to_spec(2)
    assert np.array_equal(s["lb"], [-2, -2]) and np.array_equal(s["ub"], [3, 3])   # lb = min., ub = max.
    s = fm.kernel_normal_reflective(lb=[np.nan, 0.0], ub=np.nan).to_spec(2)        # process_bounds, R/kernel.R:25-41
    assert s["lb"][0] == -A.DBL_MAX and s["ub"][1] == A.DBL_MAX
    s = fm.kernel_normal(scheme=[2, 1]).to_spec(2)
    assert s["scheme"] == A.SCHEME_EXPLICIT and list(s["order"]) == [2, 1]


def test_append_chains_mcpar():
    """inst/tinytest/test-append_chains.R:19-60 and R/append_chains.R:113-142"""
    a = Mcmc(np.zeros((100, 2)), start=1, end=100, thin=1)
    b = Mcmc(np.ones((50, 2)), start=1, end=50, thin=1)
    ab = append_chains(a, b)
    assert ab.mcpar == (1, 150, 1) and ab.niter() == 150
    t1 = Mcmc(np.zeros((10, 2)), start=110, end=200, thin=10)      # burnin 100, thin 10, nsteps 200
    t2 = Mcmc(np.zeros((20, 2)), start=10, end=200, thin=10)
    t12 = append_chains(t1, t2)
    assert t12.mcpar == (110, 400, 10) and t12.niter() == 30
    assert list(t12.iterations()[:2]) == [110, 120] and t12.iterations()[-1] == 400
    l = append_chains(McmcList([a, a]), McmcList([b, b]))
    assert isinstance(l, McmcList) and l.nchain() == 2 and l.mcpar == (1, 150, 1)
    with pytest.raises(ValueError, match="same number of chains"):
        append_chains(McmcList([a, a]), McmcList([b]))
    with pytest.raises(ValueError, match="same `thin`"):
        append_chains(a, t1)
    with pytest.raises(ValueError, match="same number of parameters"):
        append_chains(a, Mcmc(np.zeros((5, 3))))


def test_array_backed_mcmc_list_builds_its_views_on_demand():
    """McmcList.from_array (what every MCMC() call returns): shape, mcpar and names are known without creating one Mcmc per
    chain; the views appear on first element access, share the array's memory and behave like an ordinary list afterwards."""
    arr = np.arange(3 * 5 * 2, dtype=np.float64).reshape(3, 5, 2)
    l = McmcList.from_array(arr, 11, 19, 2, ["a", "b"])
    assert l._lazy and len(l) == 3 and bool(l) and l.nchain() == 3 and l.niter() == 5 and l.nvar() == 2
    assert l.mcpar == (11, 19, 2) and l.varnames == ["a", "b"] and l.as_array() is arr and l._lazy
    sel = l.select([1])
    assert sel.as_array().shape == (3, 5, 1) and sel.varnames == ["b"] and l._lazy
    m = l[2]
    assert not l._lazy and isinstance(m, Mcmc) and m.mcpar == (11, 19, 2) and np.shares_memory(m.data, arr)
    assert [x.data[0, 0] for x in l] == [0.0, 10.0, 20.0] and list.__len__(l) == 3
    assert l == McmcList.from_array(arr, 11, 19, 2, ["a", "b"]) or True     # comparison materialises both sides without error
    ap = append_chains(McmcList.from_array(arr, 1, 5, 1), McmcList.from_array(arr, 1, 5, 1))
    assert ap.nchain() == 3 and ap.niter() == 10 and ap.mcpar == (1, 10, 1)
    assert McmcList.from_array(np.empty((0, 4, 2))).nchain() == 0 and not McmcList.from_array(np.empty((0, 4, 2)))


def test_coda_window_rule():
    """gelman.diag's autoburnin: window(x, start = end/2 + 1); off-grid starts snap UP."""
    assert window_first_row(1, 200, 1, 200, 200 / 2 + 1) == 100
    assert window_first_row(1, 400, 1, 400, 201) == 200
    assert window_first_row(1, 201, 1, 201, 201 / 2 + 1) == 101          # 101.5 -> iteration 102
    assert window_first_row(10, 2000, 10, 200, 1001) == 100              # 1001 -> iteration 1010
    assert window_first_row(110, 400, 10, 30, 201) == 10                 # -> iteration 210


def test_rows_kept_and_labels():
    """R/mcmc.R:786-813: post-burnin POSITIONS with pos %% thin == 0 are kept (quirk D3)."""
    for nsteps, burnin, thin in [(10, 0, 1), (500, 100, 7), (1000, 0, 10), (11, 10, 1), (20, 3, 4)]:
        rows = np.arange(1, nsteps + 1)[burnin:]
        kept = rows[(np.arange(1, len(rows) + 1) % thin) == 0]
        assert A.rows_kept(nsteps, burnin, thin) == len(kept)
        if len(kept):
            assert kept[0] == burnin + thin and kept[-1] == burnin + len(kept) * thin


def test_kernel_list_semantics():
    """rep_kernel turns the user's object into a list in place (R/kernel.R:348-377, R/mcmc.R:526-527)."""
    k = fm.kernel_adapt()
    assert not k.is_list and len(k) == 1
    k.to_spec(3)
    k._replicate(4)
    assert k.is_list and len(k) == 4 and k[2].type == A.KERNEL_ADAPT and k[2] is not k[1]
    ist, dst = k.state_arrays(4, 3)
    assert ist.shape == (4, A.ISTATE_LEN) and dst.shape == (4, 12)
    ist[:, 0] = 7
    ist[:, 1] = A.STATE_INIT
    dst[1, :9] = np.eye(3).reshape(-1)
    k.absorb_state(3)
    assert k[1].abs_iter == 7 and np.array_equal(k[1].Sigma, np.eye(3)) and k[1].Mean_t_prev is None


def test_observation_sharding_row_slices_partition_the_rows():
    """dist.ObservationSharding.row_slice: contiguous, even-sized blocks (16-byte aligned columns) that cover [0, n)."""
    from fmcmc_b200.dist import ObservationSharding
    for n in (40_006, 1_000_001, 17):
        for world in (2, 3, 8):
            covered = []
            for rank in range(world):
                sh = ObservationSharding.__new__(ObservationSharding)
                sh.rank, sh.world = rank, world
                sl = sh.row_slice(n)
                covered.append((sl.start, sl.stop))
                if rank < world - 1:
                    assert (sl.stop - sl.start) % 2 == 0
            assert covered[0][0] == 0 and covered[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))


def test_bench_ess_estimator_on_ar1():
    """bench.py's pooled ESS (Geyer initial positive sequence) recovers (1 - phi) / (1 + phi) on an AR(1) chain."""
    import bench
    rng = np.random.default_rng(0)
    C, T, phi = 40, 3000, 0.8
    x = np.zeros((C, T, 2))
    e = rng.standard_normal((C, T, 2))
    for t in range(1, T):
        x[:, t, 0] = phi * x[:, t - 1, 0] + e[:, t, 0]
    x[:, :, 1] = e[:, :, 1]
    ess = bench.ess_pooled(x) / (C * T)
    assert abs(ess[0] - (1 - phi) / (1 + phi)) < 0.02 and abs(ess[1] - 1.0) < 0.1


def test_c_gelman_window_matches_the_python_glue_and_coda():
    """fmcmc_gelman's autoburnin window: coda only windows when start(x) < end(x)/2, to iteration
    end/2 + 1 snapped UP to the next kept iteration; cases from the advisor's report + a sweep against the Python
    restatement the MCMC() glue uses."""
    import fmcmc_b200
    L = fmcmc_b200.lib()
    wb = L.fmcmc_gelman_window_begin
    assert wb(1, 1, 200) == 100 and wb(1, 1, 400) == 200
    assert wb(1, 1, 201) == 101                  # end/2 + 1 = 101.5 -> iteration 102
    assert wb(1, 1, 5) == 3                      # rows = 5: coda keeps iterations 4, 5 (2 rows), not 3
    assert wb(501, 1, 1000) == 250               # burnin = 500, 1000 kept rows: start 501 < 750 -> from iteration 751: 750 rows
    assert wb(1501, 1, 1000) == 0                # start >= end/2: no window at all
    assert wb(10, 10, 200) == 100 and wb(110, 10, 30) == 10
    assert wb(1, 1, 2) == 0                      # start = 1 is not < end/2 = 1: both rows stay (N = 2)
    for start in (1, 2, 7, 50, 101, 1000):
        for thin in (1, 2, 3, 10):
            for rows in (2, 3, 5, 10, 11, 64, 101, 1000):
                end = start + (rows - 1) * thin
                want = window_first_row(start, end, thin, rows, end / 2 + 1) if start < end / 2 else 0
                assert wb(start, thin, rows) == want, (start, thin, rows)


def test_host_sym_eigmax_against_numpy():
    """The Gelman finish's scalar tail: Householder tridiagonalisation + Sturm bisection == numpy's eigvalsh."""
    import ctypes as C
    import fmcmc_b200
    L = fmcmc_b200.lib()
    rng = np.random.default_rng(0)
    for p in (1, 2, 3, 5, 32, 128, 129):
        for kind in ("spd", "indef", "diag", "rank1", "blocks"):
            G = rng.standard_normal((p, p))
            if kind == "spd":
                M = G @ G.T / p
            elif kind == "indef":
                M = G + G.T
            elif kind == "diag":
                M = np.diag(rng.standard_normal(p))
            elif kind == "rank1":
                v = rng.standard_normal(p)
                M = np.outer(v, v) + 1e-4 * np.eye(p)
            else:
                M = np.zeros((p, p))
                h = p // 2
                M[:h, :h] = (G @ G.T)[:h, :h]
                M[h:, h:] = 3 * np.eye(p - h)
            A_ = np.asfortranarray(M)
            out = C.c_double()
            assert L.fmcmc_host_sym_eigmax(p, A_.ctypes.data_as(C.POINTER(C.c_double)), C.byref(out)) == 0
            want = np.linalg.eigvalsh(M)[-1]
            scale = max(np.abs(np.linalg.eigvalsh(M)).max(), 1e-300)
            assert abs(out.value - want) <= 1e-13 * scale, (p, kind, out.value, want)


# ---- single-chain diagnostics (fmcmc_b200/diagnostics.py: coda::geweke.diag / heidel.diag / spectrum0.ar, stats::ar restated) ----

def _ar1(n, phi, seed, drift=0.0):
    rng = np.random.default_rng(seed)
    e = rng.standard_normal(n)
    x = np.zeros(n)
    for t in range(1, n):
        x[t] = phi * x[t - 1] + e[t]
    return x + drift * np.arange(n)


def test_ar_yule_walker_against_toeplitz_solves():
    """Levinson-Durbin + AIC selection == solving the Yule-Walker system of every order directly (scipy Toeplitz solver) and
    picking the order with the smallest n log(v_k) + 2 k."""
    from scipy.linalg import solve_toeplitz
    from fmcmc_b200.diagnostics import ar_yule_walker
    for seed, phi in ((1, 0.5), (2, 0.9), (3, -0.4)):
        x = _ar1(1500, phi, seed)
        n = x.size
        xc = x - x.mean()
        omax = int(min(n - 1, np.floor(10 * np.log10(n))))
        r = np.array([xc[:n - l] @ xc[l:] / n for l in range(omax + 1)])
        best = (n * np.log(r[0]) + 2.0, np.zeros(0), r[0])
        for k in range(1, omax + 1):
            a = solve_toeplitz(r[:k], r[1:k + 1])
            v = r[0] - a @ r[1:k + 1]
            aic = n * np.log(v) + 2 * k + 2.0
            if aic < best[0]:
                best = (aic, a, v)
        ar, var_pred, order = ar_yule_walker(x)
        assert order == best[1].size and order >= 1
        np.testing.assert_allclose(ar, best[1], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(var_pred, best[2] * n / (n - (order + 1)), rtol=1e-10)


def test_spectrum0_and_cramer_von_mises_known_values():
    from fmcmc_b200.diagnostics import pcramer, spectrum0_ar
    # AR(1) with unit innovations: spectral density at zero = 1 / (1 - phi)^2
    s0 = np.mean([spectrum0_ar(_ar1(20000, 0.6, s))[0] for s in range(5)])
    assert abs(s0 / (1 / 0.4 ** 2) - 1) < 0.08
    assert spectrum0_ar(np.full(50, 3.0))[0] == 0.0 and spectrum0_ar(np.arange(50.0))[0] == 0.0   # no variation about the trend
    # the asymptotic Cramer-von Mises law: 5 % and 1 % points 0.46136 and 0.74346 (Anderson & Darling 1952, table 1)
    assert abs(pcramer(0.46136) - 0.95) < 1e-4 and abs(pcramer(0.74346) - 0.99) < 1e-4


def test_geweke_and_heidel_on_stationary_and_drifting_chains():
    import fmcmc_b200 as fm
    from fmcmc_b200.diagnostics import geweke_z, heidel_diag
    n = 4000
    good = np.c_[_ar1(n, 0.5, 11) + 10.0, _ar1(n, 0.2, 12) - 3.0]
    bad = np.c_[_ar1(n, 0.5, 13) + 10.0, _ar1(n, 0.5, 14, drift=2e-3)]
    zg, zb = geweke_z(fm.Mcmc(good, 1, n, 1)), geweke_z(fm.Mcmc(bad, 1, n, 1))
    assert np.all(np.abs(zg) < 3.5) and abs(zb[1]) > 6
    # windows live on the iteration grid start + thin * row: starts snap up, ends snap down (coda's window.mcmc)
    from fmcmc_b200.diagnostics import _window_rows
    assert _window_rows(5, 32, 3, 10, wstart=10) == (2, 10)           # iterations 5, 8, 11, ...: the first >= 10 is row 2
    assert _window_rows(5, 32, 3, 10, wend=20) == (0, 6)              # the last <= 20 is iteration 20 = row 5
    assert _window_rows(5, 32, 3, 10, wstart=11, wend=11) == (2, 3)
    zt = geweke_z(fm.Mcmc(good, 5, 5 + 3 * (n - 1), 3))
    assert np.all(np.isfinite(zt)) and np.all(np.abs(zt - zg) < 0.1)  # a row more or less at the window edges
    hg, hb = heidel_diag(fm.Mcmc(good, 1, n, 1)), heidel_diag(fm.Mcmc(bad, 1, n, 1))
    assert np.all(hg[:, 0] == 1) and np.all(hg[:, 3] == 1) and np.all(hg[:, 1] >= 1)
    np.testing.assert_allclose(hg[0, 4], good[int(hg[0, 1]) - 1:, 0].mean(), rtol=1e-12)   # mean of the part that passed
    assert hb[1, 0] == 0 and np.isnan(hb[1, 3])                       # the drifting column fails the stationarity test


def test_single_chain_checkers_follow_the_reference():
    """R/convergence.R:259-389: messages, side channel, errors with more than one chain, the literal decision rule of
    convergence_geweke (it compares 1 - p-value with the threshold), convergence_auto's choice."""
    import warnings as W
    import fmcmc_b200 as fm
    n = 2000
    one = fm.Mcmc(np.c_[_ar1(n, 0.5, 21) + 5.0, _ar1(n, 0.3, 22) + 1.0], 1, n, 1)
    g = fm.convergence_geweke(500)
    assert g.freq == 500 and isinstance(g(one), bool)
    assert fm.convergence_msg_get().startswith("avg Geweke's Z: ")
    assert n in fm.convergence_data_get("dat")
    h = fm.convergence_heildel(500)
    assert h(one) is True and fm.convergence_msg_get().startswith("Heidel's Avg. pval: ")
    two = fm.McmcList([one, one])
    with pytest.raises(ValueError, match="only available with runs of a single chain"):
        g(two)
    with pytest.raises(ValueError, match="only available with runs of a single chain"):
        h(two)
    const = fm.Mcmc(np.full((100, 1), 2.0), 1, 100, 1)                 # rm_invariant removes the only column -> warning + FALSE
    with W.catch_warnings(record=True) as rec:
        W.simplefilter("always")
        assert g(const) is False and h(const) is False
    assert any("failed to be computed" in str(r.message) for r in rec)
    a = fm.convergence_auto(300)
    assert a.freq == 300 and a.gelman.freq == 300 and a.geweke.freq == 300
