"""Configurations round 1 left untested (VERDICT r01 "What's missing" 2 and 4), all against the CPU oracle on fed streams:
MASS::mvrnorm's eigen draw on the device (FMCMC_MVN_EIGEN), scheme = "random" (planned sequence fed, and Philox),
kernel_ram's `constr`, kernel_adapt past its warm-up at BASELINE configs[2]'s full n, configs[4]'s geometry with
kernel_nmirror, user-supplied Sigma with several chains."""
import numpy as np
import pytest

from fmcmc_b200 import _abi as A
from gpu_util import assert_parity, run_both
from test_gpu_parity import _kernels, _logistic_family, _readme_family

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _gaussian_family(rng, n, p):
    from fmcmc_b200 import ll_gaussian_lm
    X = rng.standard_normal((n, p))
    y = 1.0 + X @ rng.standard_normal(p) + rng.normal(0, 2.0, n)
    return ll_gaussian_lm(X, y, intercept=True, guard=True), p + 2


# ---- A10: the reference's own draw, mu + V sqrt(max(ev, 0)) z (R/kernel_adapt.R:173-178) ----------------------------
@pytest.mark.parametrize("nchains", [1, 4, 320])
@pytest.mark.parametrize("variant", ["freq1", "freq3", "bw30"])
def test_adapt_eigen_draw_readme_model(oracle, readme_data, nchains, variant):
    """Device Jacobi == oracle Jacobi bit for bit (same rotations, unfused arithmetic), so decisions are identical and
    samples agree to 1e-12 even through the exactly repeated eigenvalue of the first adapted Sigma."""
    name = {"freq1": "adapt", "freq3": "adapt_freq3", "bw30": "adapt_bw30"}[variant]
    if variant == "bw30" and nchains > 4:
        # Windowed branch: Sigma = Sd (cov(window) + eps I).  A chain that sat still for the whole window has
        # cov = rounding noise, i.e. an eigenvalue repeated to ~1e-16: the eigenBASIS is then decided by that noise
        # (in R as well: by cov()'s long-double accumulation), and the device's compensated sums are not the oracle's
        # long doubles bit for bit.  The recursive branch is bit-identical, so it is tested at every chain count.
        pytest.skip("eigen draw of a numerically repeated eigenvalue is rounding-noise dependent (see comment)")
    spec = dict(_kernels(3)[name], mvn_method=A.MVN_EIGEN)
    rng = np.random.default_rng(101)
    init = np.tile([1.0, 1.0, readme_data["sd_y"]], (nchains, 1)) + rng.normal(0, 0.05, (nchains, 3))
    T = 400 if nchains < 100 else 160
    g, o, st = run_both(oracle, _readme_family(readme_data), spec, init, T, nchains, rng=rng)
    assert_parity(g[0], o[0], RTOL, f"adapt eigen {variant}")
    assert np.array_equal(st[0], st[2])
    np.testing.assert_allclose(st[1], st[3], rtol=1e-10, atol=1e-12 * np.abs(st[3]).max())


@pytest.mark.parametrize("path", [1, 2, 3, 4])
@pytest.mark.parametrize("p", [7, 32, 60])
def test_adapt_eigen_draw_tiled_paths(oracle, path, p):
    """Eigen draw behind every stepping path; p = 32 is the bench kernel's k (factor in shared memory), p = 60 runs the
    Jacobi in global memory (2 kf^2 doubles of scratch exceed the head kernel's shared-memory budget)."""
    if path == 2 and p > 32:
        pytest.skip("path 2 handles p_x <= 32")
    rng = np.random.default_rng(102 + p)
    n, C, T = 2 * 128 * 3 + 77, 40, 60
    fam = _logistic_family(rng, n, p)
    spec = dict(type=A.KERNEL_ADAPT, k=p, mu=0.0, warmup=12, freq=1, eps=1e-4, mvn_method=A.MVN_EIGEN)
    g, o, _ = run_both(oracle, fam, spec, rng.normal(0, 0.05, (C, p)), T, C, rng=rng, path=path)
    assert g[0]["report"].path == path
    assert_parity(g[0], o[0], RTOL, f"adapt eigen path {path} p {p}")


def test_adapt_eigen_warmup_is_the_exchange_matrix(oracle, readme_data):
    """R's eigen(eps * I) returns the eigenvectors in reversed order: during the warm-up the device draw is
    sqrt(eps) * (z_k, ..., z_1) exactly (pinned on the reference's published trace, test_oracle_readme_golden.py)."""
    from fmcmc_b200.device import DeviceModel
    rng = np.random.default_rng(103)
    T = 30
    logu = np.log(rng.random((1, T)))
    z = rng.standard_normal((1, T, 3))
    spec = dict(type=A.KERNEL_ADAPT, k=3, mu=0.0, warmup=500, freq=1, eps=1e-4, mvn_method=A.MVN_EIGEN)
    m = DeviceModel(_readme_family(readme_data))
    g = m.run(spec, T, 1, initial=[3.0, 2.0, 4.0], stream=A.marshal_stream(A.STREAM_FED, logu=logu, z=z))
    m.close()
    prev = g["ans"][0, :-1]
    want = prev + np.sqrt(1e-4) * z[0, 1:, ::-1]
    assert np.array_equal(g["draws"][0, 1:], want)


# ---- A6: scheme = "random" (R/kernel.R:106-113) ------------------------------------------------------------------------
@pytest.mark.parametrize("path", [1, 3, 4])
@pytest.mark.parametrize("kname", ["normal", "unif_reflective", "nmirror"])
def test_random_scheme_fed_sequence(oracle, kname, path):
    """The planned sequence sample(which(!fixed), nsteps, TRUE) is uploaded with the stream (fed mode): one coordinate
    moves per row, chosen per chain and row; fixed coordinates never appear."""
    rng = np.random.default_rng(104)
    n, p, C, T = 900, 6, 40, 120
    fam = _logistic_family(rng, n, p)
    fixed = [False, False, True, False, False, False]
    free = np.flatnonzero(~np.asarray(fixed)) + 1
    seq = rng.choice(free, size=(C, T)).astype(np.int32)
    spec = dict(_kernels(p)[kname], scheme=A.SCHEME_RANDOM, seq=seq, fixed=fixed)
    for key in ("lb", "ub"):
        if key in spec:
            spec[key] = -5.0 if key == "lb" else 5.0
    g, o, _ = run_both(oracle, fam, spec, rng.normal(0, 0.1, (C, p)), T, C, rng=rng, path=path)
    assert g[0]["report"].path == path
    assert_parity(g[0], o[0], RTOL, f"random scheme fed {kname} path {path}")
    moved = g[0]["draws"][:, 1:] != g[0]["ans"][:, :-1]                    # proposal vs previous state
    assert moved.sum(axis=2).max() == 1 and not moved[:, :, 2].any()
    rows, chains = np.nonzero(moved.any(axis=2).T)
    assert np.array_equal(np.argmax(moved, axis=2)[chains, rows] + 1, seq[chains, rows + 1])


@pytest.mark.parametrize("path", [1, 3])
def test_random_scheme_philox(oracle, path):
    """Production mode: the coordinate comes from the chain's Philox plan stream (run = PLAN_RUN), same as the oracle."""
    rng = np.random.default_rng(105)
    fam = _logistic_family(rng, 700, 5)
    spec = dict(type=A.KERNEL_NORMAL, k=5, mu=0.0, scale=0.2, scheme=A.SCHEME_RANDOM)
    g, o, _ = run_both(oracle, fam, spec, np.zeros(5), 150, 24, path=path, philox_seed=99, chain_offset=7)
    assert_parity(g[0], o[0], 1e-9, "random scheme philox")
    moved = g[0]["draws"][:, 1:] != g[0]["ans"][:, :-1]
    assert moved.sum(axis=2).max() == 1
    counts = np.bincount(np.argmax(moved, axis=2)[moved.any(axis=2)], minlength=5)
    assert counts.min() > 0.1 * counts.sum()                               # every coordinate is visited


# ---- A12: kernel_ram(constr = ) (R/kernel_ram.R:149-150) -------------------------------------------------------------
@pytest.mark.parametrize("path", [1, 3, 4])
def test_ram_constr(oracle, path):
    """Sigma <- constr[which., which.] * Sigma after every adaptation: a block-diagonal 0/1 mask keeps two parameter
    groups uncorrelated; one fixed coordinate exercises the which. sub-setting."""
    rng = np.random.default_rng(106)
    n, p, C, T = 2 * 128 * 3 + 5, 6, 40, 120
    fam = _logistic_family(rng, n, p)
    constr = np.zeros((p, p))
    constr[:3, :3] = 1.0
    constr[3:, 3:] = 1.0
    fixed = [False, False, False, False, True, False]
    spec = dict(type=A.KERNEL_RAM, k=p, warmup=0, freq=1, eps=1e-2, arate=0.234, constr=constr, fixed=fixed)
    g, o, st = run_both(oracle, fam, spec, rng.normal(0, 0.1, (C, p)), T, C, rng=rng, path=path)
    assert g[0]["report"].path == path
    assert_parity(g[0], o[0], RTOL, f"ram constr path {path}")
    S = st[1][:, :25].reshape(C, 5, 5)                                     # col-major kf x kf factor of every chain
    free_block = np.array([0, 0, 0, 1, 1])                                 # free params 0,1,2 | 3,5
    off = free_block[:, None] != free_block[None, :]
    assert np.all(S[:, off] == 0.0) and np.any(S[:, ~off] != 0.0)
    np.testing.assert_allclose(st[1], st[3], rtol=1e-9, atol=1e-12)


# ---- user-supplied Sigma with several chains --------------------------------------------------------------
@pytest.mark.parametrize("maker", ["kernel_adapt", "kernel_ram"])
def test_user_sigma_reaches_every_chain(readme_data, maker):
    """rep_kernel copies Sigma into every chain's environment (R/kernel.R:407-434): the first proposals already have
    the user's scale, not eps * I."""
    import fmcmc_b200 as fm
    fam = _readme_family(readme_data)
    big = np.diag([0.04, 0.09, 0.01])
    kern = getattr(fm, maker)(Sigma=big, warmup=10_000)
    out = fm.MCMC(np.tile([3.0, 2.0, 4.0], (4, 1)), fam, nsteps=400, nchains=4, kernel=kern, seed=5)
    step = np.concatenate([np.asarray(fm.get_draws()[c])[1:] - np.asarray(out[c])[:-1] for c in range(4)])
    sd = step.std(axis=0)
    if maker == "kernel_adapt":                                            # N(0, Sigma)
        np.testing.assert_allclose(sd, np.sqrt(np.diag(big)), rtol=0.15)
    else:                                                                  # Sigma is RAM's FACTOR; U ~ t_3 (heavy tails)
        assert np.all(sd > 5 * 1e-4) and sd[1] > sd[0] > sd[2]


# ---- full-size checks against the oracle -------------------------------------------------------------------------------
def test_cfg3_kernel_adapt_past_warmup_full_n(oracle):
    """BASELINE configs[2] (logistic n = 1e6, p = 32, kernel_adapt) against the CPU oracle with the adaptation ACTIVE:
    warmup = 2, so rows 5.. update Mean / Sigma by cov_recursive and factorise a dense 32 x 32 Sigma on the device
    (default path: split-integer tcgen05 kernel).  160 chains x 8 rows."""
    import fmcmc_b200 as fm
    import bench
    X, y = bench.make_data()                                               # n = 1e6, p = 32
    p = X.shape[1]
    C, T = 160, 8
    rng = np.random.default_rng(18)
    init = rng.normal(0, 0.1, (C, p))
    spec = dict(type=A.KERNEL_ADAPT, k=p, mu=0.0, warmup=2, freq=1, eps=1e-6)
    g, o, st = run_both(oracle, fm.ll_logistic(X, y), spec, init, T, C, rng=rng)
    assert g[0]["report"].path == 4
    assert_parity(g[0], o[0], RTOL, "cfg3 adapt full n")
    assert np.all(st[0][:, 0] == T - 1) and np.all(st[0][:, 1] & A.STATE_HAS_MEAN)
    np.testing.assert_allclose(st[1], st[3], rtol=1e-10, atol=1e-12 * np.abs(st[3]).max())
    acc = np.any(g[0]["ans"][:, 1:] != g[0]["ans"][:, :-1], axis=2)
    assert 0.02 < acc.mean() < 0.98


def test_cfg5_shape_kernel_nmirror_against_the_oracle(oracle):
    """BASELINE configs[4]'s geometry AND kernel: Gaussian, 127 columns + sd, kernel_nmirror(lb = c(rep(NA, 127), 0)) with
    its mean / acceptance-rate / scale adaptation inside the run (warmup 12, nadapt 4, 8, 12), n = 200 000."""
    import fmcmc_b200 as fm
    rng = np.random.default_rng(24)
    n, p, C, T = 200_000, 127, 160, 18
    X = np.empty((n, p), order="F")
    X[:, 0] = 1.0
    for j in range(1, p):
        X[:, j] = rng.standard_normal(n)
    beta = rng.standard_normal(p)
    y = X @ beta + 2.0 * rng.standard_normal(n)
    fam = fm.ll_gaussian_lm(X, y, intercept=False, guard=True)
    k = p + 1
    lb = np.full(k, -A.DBL_MAX); lb[-1] = 0.0
    spec = dict(type=A.KERNEL_NMIRROR, k=k, mu=np.r_[beta, 2.0], scale=3e-4, warmup=12, arate=0.4, lb=lb, ub=A.DBL_MAX,
                nadapt=np.array([4, 8, 12]))
    init = np.c_[beta + rng.normal(0, 1e-3, (C, p)), rng.uniform(1.9, 2.1, C)]
    g, o, st = run_both(oracle, fam, spec, init, T, C, rng=rng)
    assert g[0]["report"].path == 4
    assert_parity(g[0], o[0], RTOL, "cfg5 shape nmirror")
    np.testing.assert_allclose(st[1], st[3], rtol=1e-10, atol=1e-300)


# ---- A14: coda::gelman.diag on the device at scale (tiled moments + SYRK, gelman.cuh) ------------------------------------
@pytest.mark.parametrize("p,C,T,fixed_some", [(1, 4, 203, False), (30, 37, 150, True), (62, 20, 97, False),
                                               (127, 300, 45, False), (127, 12, 333, True), (150, 9, 260, False)])
def test_gelman_tiled_matches_oracle(oracle, p, C, T, fixed_some):
    """kf = 3 .. 152 free parameters (16 / 32 / 64 / 128 / 144-wide register tiles and, beyond 144, the per-pair kernel),
    ragged windows (N not a multiple of the 16-row tile), fixed parameters (gathered columns), more chains than CTAs."""
    from fmcmc_b200.device import DeviceModel
    rng = np.random.default_rng(200 + p)
    fam, k = _gaussian_family(rng, 64, p)
    fixed = np.zeros(k, dtype=bool)
    if fixed_some:
        fixed[[1, k // 2]] = True
    spec = dict(type=A.KERNEL_NORMAL, k=k, mu=0.0, scale=0.02, fixed=fixed)
    init = np.c_[rng.normal(0, 0.3, (C, k - 1)), rng.uniform(2.0, 4.0, C)]
    m = DeviceModel(fam)
    m.store_reset(C, T)
    g = m.run(spec, T, C, initial=init, flags=A.RUN_APPEND, stream=A.marshal_stream(A.STREAM_PHILOX, seed=3, run_index=0))
    free = (~fixed).astype(np.uint8)
    psrf, mpsrf, used = m.gelman(free, start_iter=1, thin=1)
    first = T // 2 if T % 2 == 0 else T // 2 + 1
    assert used == T - first
    xb, s2, ws = m.gelman_partials(first, T, free, C)
    m.close()
    w = g["ans"][:, first:, :][:, :, ~fixed]
    np.testing.assert_allclose(xb, w.mean(axis=1), rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(s2, w.var(axis=1, ddof=1), rtol=1e-10)
    W = sum(np.cov(w[c].T, ddof=1).reshape(k - fixed.sum(), -1) for c in range(C))
    np.testing.assert_allclose(ws, W, rtol=1e-9, atol=1e-12 * np.abs(W).max())
    o_psrf, o_mpsrf, rc = oracle.gelman(w)
    if rc == 0:
        np.testing.assert_allclose(psrf, o_psrf, rtol=1e-9)
        np.testing.assert_allclose(mpsrf, o_mpsrf, rtol=1e-8)


@pytest.mark.parametrize("burnin,thin,T", [(0, 1, 201), (500, 1, 1500), (100, 3, 700), (0, 2, 11), (900, 1, 1000)])
def test_fmcmc_gelman_window_is_codas(oracle, readme_data, burnin, thin, T):
    """fmcmc_gelman (the entry the R glue calls) on stores with burnin / thin: same window as coda."""
    from fmcmc_b200.coda import window_first_row
    from fmcmc_b200.device import DeviceModel
    C = 5
    m = DeviceModel(_readme_family(readme_data))
    keep = (T - burnin) // thin
    m.store_reset(C, keep)
    spec = dict(type=A.KERNEL_NORMAL, k=3, mu=0.0, scale=0.1)
    g = m.run(spec, T, C, initial=[3.0, 2.0, 4.0], burnin=burnin, thin=thin, flags=A.RUN_APPEND,
              stream=A.marshal_stream(A.STREAM_PHILOX, seed=8, run_index=0))
    start, end = g["report"].first_iter, g["report"].last_iter
    psrf, mpsrf, used = m.gelman(np.ones(3, dtype=np.uint8), start_iter=start, thin=thin)
    m.close()
    first = window_first_row(start, end, thin, keep, end / 2 + 1) if start < end / 2 else 0
    assert used == keep - first
    o_psrf, o_mpsrf, rc = oracle.gelman(g["ans"][:, first:, :])
    assert rc == 0
    np.testing.assert_allclose(psrf, o_psrf, rtol=1e-9)
    np.testing.assert_allclose(mpsrf, o_mpsrf, rtol=1e-9)
