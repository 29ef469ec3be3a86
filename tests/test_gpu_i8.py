"""Path 4 — the split-integer tcgen05 likelihood kernel (fmcmc_b200/csrc/tiled_i8.cuh) — against the CPU oracle.

Same fed-stream contract as the FP64 paths (every accept/reject decision identical, samples within 1e-12
relative): the int8 slicing of X and Theta is exact until the FP64 reassembly, and the truncated slice pairs
leave an error far below the band."""
import numpy as np
import pytest

from fmcmc_b200 import _abi as A
from gpu_util import assert_parity, run_both
from test_gpu_parity import _kernels, _logistic_family

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _gaussian_family(rng, n, p):
    from fmcmc_b200 import ll_gaussian_lm
    X = rng.standard_normal((n, p))
    y = 1.0 + X @ rng.standard_normal(p) + rng.normal(0, 2.0, n)
    return ll_gaussian_lm(X, y, intercept=True, guard=True), p + 2


@pytest.mark.parametrize("family", ["logistic", "gaussian"])
@pytest.mark.parametrize("kname", ["normal", "normal_reflective", "adapt", "ram", "nmirror", "unif"])
def test_i8_path_parity(oracle, family, kname):
    """Ragged n (tail tile), chains not a multiple of the 128-chain block, every kernel class (RAM adds the
    second likelihood column block and, because its adaptation consumes f itself, runs on 7 slices)."""
    rng = np.random.default_rng(33)
    n, p = 2 * 128 * 3 + 77, 7
    if family == "logistic":
        fam, k = _logistic_family(rng, n, p), p
        init = rng.normal(0, 0.1, (70, k))
    else:
        fam, k = _gaussian_family(rng, n, p)
        init = np.c_[rng.normal(0, 0.1, (70, k - 1)), np.full(70, 3.0)]
    spec = dict(_kernels(k)[kname])
    if family == "logistic":
        for key in ("lb", "ub"):
            if key in spec:
                spec[key] = -A.DBL_MAX if key == "lb" else A.DBL_MAX
    if "scale" in spec:
        spec["scale"] = 0.05
    g, o, _ = run_both(oracle, fam, spec, init, 140, 70, rng=rng, path=4)
    assert g[0]["report"].path == 4
    assert_parity(g[0], o[0], RTOL, f"{family}/{kname}")


def test_i8_many_chain_blocks(oracle):
    """600 chains => 5 chain blocks of 128; p_x = 32 (the bench's shape: one K block, Theta slices in TMEM)."""
    rng = np.random.default_rng(44)
    fam = _logistic_family(rng, 1500, 32)
    spec = dict(type=A.KERNEL_NORMAL, k=32, mu=0.0, scale=0.03)
    g, o, _ = run_both(oracle, fam, spec, rng.normal(0, 0.1, (600, 32)), 40, 600, rng=rng, path=4)
    assert g[0]["report"].path == 4
    assert_parity(g[0], o[0], RTOL)


@pytest.mark.parametrize("family,p", [("gaussian", 127), ("gaussian", 50), ("logistic", 100), ("logistic", 33)])
def test_i8_wide_design_matrix(oracle, family, p):
    """p_x > 32: 2 / 4 K blocks per slice pair; at 4 the Theta slices live in shared memory (SS form of the MMA)."""
    rng = np.random.default_rng(55)
    n, C = 1000 + 13, 150
    if family == "logistic":
        fam, k = _logistic_family(rng, n, p), p
        init = rng.normal(0, 0.05, (C, k))
        spec = dict(type=A.KERNEL_NORMAL, k=k, mu=0.0, scale=0.02)
    else:
        fam, k = _gaussian_family(rng, n, p)
        init = np.c_[rng.normal(0, 0.1, (C, k - 1)), np.full(C, 3.0)]
        lb = np.full(k, -A.DBL_MAX); lb[-1] = 0.0
        spec = dict(type=A.KERNEL_NMIRROR, k=k, mu=0.0, scale=0.05, warmup=30, arate=0.4, lb=lb, ub=A.DBL_MAX,
                    nadapt=np.array([10, 20, 30]))
    g, o, _ = run_both(oracle, fam, spec, init, 60, C, rng=rng, path=4)
    assert g[0]["report"].path == 4
    assert_parity(g[0], o[0], RTOL, f"{family}/p={p}")


def test_i8_badly_scaled_columns(oracle):
    """Rows / chains whose entries span many orders of magnitude: the per-row and per-chain exponents keep the
    absolute error of eta at ~2^-47 of |x|max |theta|max, which is what the 1e-12 band on the log-posterior needs."""
    rng = np.random.default_rng(77)
    n, p, C = 900, 12, 40
    from fmcmc_b200 import ll_logistic
    X = rng.standard_normal((n, p)) * np.logspace(-6, 3, p)
    X[:, 0] = 1.0
    beta = rng.standard_normal(p) / np.logspace(-6, 3, p)
    y = (rng.random(n) < 1 / (1 + np.exp(-X @ beta))).astype(np.float64)
    fam = ll_logistic(X, y, prior_sd=2.0)
    spec = dict(type=A.KERNEL_NORMAL, k=p, mu=0.0, scale=0.02 / np.logspace(-6, 3, p))
    init = beta + rng.normal(0, 0.01, (C, p)) / np.logspace(-6, 3, p)
    g, o, _ = run_both(oracle, fam, spec, init, 80, C, rng=rng, path=4)
    assert_parity(g[0], o[0], RTOL)


def test_i8_nonbinary_response_and_nan(oracle):
    """y outside {0, 1} contributes nothing (sum(logp[y == 1]) + sum(logq[y == 0])); a NaN parameter aborts like
    R/mcmc.R:758-765."""
    from fmcmc_b200 import ll_logistic, _lib
    from fmcmc_b200.device import DeviceModel
    rng = np.random.default_rng(88)
    n, p, C = 700, 5, 9
    X = rng.standard_normal((n, p)); X[:, 0] = 1.0
    y = (rng.random(n) < 0.5).astype(np.float64)
    y[::17] = 0.5
    fam = ll_logistic(X, y, prior_sd=2.0)
    spec = dict(type=A.KERNEL_NORMAL, k=p, mu=0.0, scale=0.05)
    g, o, _ = run_both(oracle, fam, spec, rng.normal(0, 0.1, (C, p)), 60, C, rng=rng, path=4)
    assert_parity(g[0], o[0], RTOL)
    m = DeviceModel(fam)
    m.set_path(4)
    init = rng.normal(0, 0.1, (C, p))
    init[3, 2] = np.nan
    with pytest.raises(_lib.FmcmcError) as ei:
        m.run(spec, 10, C, initial=init, stream=A.marshal_stream(A.STREAM_PHILOX, seed=1))
    m.close()
    assert "undefined" in str(ei.value)


def test_i8_matches_fp64_kernel_at_scale(oracle):
    """n = 200 000, p = 32, 256 chains (too slow for the oracle): the log-posteriors of the tcgen05 path and the
    FP64 DMMA path agree to 1e-13 relative and the chains take identical decisions."""
    from fmcmc_b200.device import DeviceModel
    rng = np.random.default_rng(99)
    n, p, C, T = 200_000, 32, 256, 30
    fam = _logistic_family(rng, n, p)
    spec = dict(type=A.KERNEL_NORMAL, k=p, mu=0.0, scale=0.004)
    init = rng.normal(0, 0.05, (C, p))
    outs = {}
    for path in (3, 4):
        m = DeviceModel(fam)
        m.set_path(path)
        outs[path] = m.run(spec, T, C, initial=init, stream=A.marshal_stream(A.STREAM_PHILOX, seed=5))
        assert outs[path]["report"].path == path
        m.close()
    a, b = outs[4], outs[3]
    assert np.array_equal(a["ans"], b["ans"])
    err = np.max(np.abs(a["logpost"] - b["logpost"]) / np.abs(b["logpost"]))
    assert err < 1e-13, err


def test_i8_falls_back_on_unsliceable_design_matrix(oracle):
    """An entry beyond the exponent window (|x| >= 2^480; likewise Inf / NaN) cannot be sliced: the auto-selected path
    silently falls back to the FP64 DMMA kernel (same results as the oracle), forcing path 4 is an error."""
    from fmcmc_b200 import ll_gaussian_lm, _lib
    from fmcmc_b200.device import DeviceModel
    rng = np.random.default_rng(12)
    n, p, C = 700, 20, 140
    X = rng.standard_normal((n, p))
    y = 1.0 + X @ rng.standard_normal(p) + rng.normal(0, 2.0, n)
    X[17, 3] = 1e200                                    # its coefficient is pinned at 0, so every mean stays finite
    fam = ll_gaussian_lm(X, y, intercept=True, guard=True)
    k = p + 2
    fixed = np.zeros(k, dtype=bool); fixed[4] = True    # parameter 0 = intercept, 1 + j = coefficient of column j
    spec = dict(type=A.KERNEL_NORMAL, k=k, mu=0.0, scale=0.05, fixed=fixed)
    init = np.c_[rng.normal(0, 0.1, (C, k - 1)), np.full(C, 3.0)]
    init[:, 4] = 0.0
    g, o, _ = run_both(oracle, fam, spec, init, 20, C, rng=rng)
    assert g[0]["report"].path == 3
    assert_parity(g[0], o[0], RTOL)
    m = DeviceModel(fam)
    m.set_path(4)
    with pytest.raises(_lib.FmcmcError) as ei:
        m.run(spec, 5, C, initial=init, stream=A.marshal_stream(A.STREAM_PHILOX, seed=1))
    m.close()
    assert "finite" in str(ei.value)


def test_i8_large_linear_predictors(oracle):
    """|eta| far beyond the table of the logistic epilogue (40): no chain passes the |eta| <= 39.9 bound, so every warp takes
    the argument-clamped variant (sum |t| apart + g = h(a_c) - a_c / 2), whatever the magnitude; it must reproduce the oracle
    (terms ~ -|eta| or ~ 0)."""
    from fmcmc_b200 import ll_logistic
    rng = np.random.default_rng(21)
    n, p, C = 900, 10, 150
    X = rng.standard_normal((n, p)); X[:, 0] = 1.0
    y = (rng.random(n) < 0.5).astype(np.float64)
    fam = ll_logistic(X, y, prior_sd=2.0)
    spec = dict(type=A.KERNEL_NORMAL, k=p, mu=0.0, scale=0.5)
    init = rng.normal(0, 30.0, (C, p))                  # |eta| ~ 100
    init[7] *= 1e5                                      # |theta| ~ 3e6 > 2^20: clamped variant for that warp
    init[140] *= 1e9
    g, o, _ = run_both(oracle, fam, spec, init, 25, C, rng=rng, path=4)
    assert g[0]["report"].path == 4
    assert_parity(g[0], o[0], RTOL)


def test_i8_mixed_bounded_and_unbounded_warps(oracle):
    """The un-clamped log(2 cosh) epilogue is chosen per warp from the chains' bound on |eta| (min of sum_j |theta_j| max_i |x_ij|
    and |theta|_2 max_i |x_i|_2 <= 39.9): here warps 0 and 2 of the first chain block pass it, warp 1 holds one chain with
    |eta| ~ 300 and warp 3 chains whose bound sits just above 39.9 although their actual |eta| stays small - all four must
    reproduce the oracle, and the last partial tile (n is not a multiple of 128) goes through the clamped variant anyway."""
    from fmcmc_b200 import ll_logistic
    rng = np.random.default_rng(33)
    n, p, C = 1100, 12, 140
    X = rng.standard_normal((n, p)) / np.sqrt(p); X[:, 0] = 1.0
    y = (rng.random(n) < 1 / (1 + np.exp(-X @ rng.standard_normal(p)))).astype(np.float64)
    fam = ll_logistic(X, y, prior_sd=2.0)
    spec = dict(type=A.KERNEL_NORMAL, k=p, mu=0.0, scale=0.02)
    init = rng.normal(0, 0.5, (C, p))
    init[40] = rng.normal(0, 120.0, p)                  # warp 1: one chain far outside the table
    rmax = np.sqrt((X * X).sum(axis=1).max())
    cmax = np.abs(X).max(axis=0)
    for c in range(96, 128):                            # warp 3: both bounds a little above 39.9
        v = rng.normal(0, 1.0, p)
        v *= 1.02 * 39.9 / min(np.abs(v) @ cmax, np.linalg.norm(v) * rmax)
        init[c] = v
    g, o, _ = run_both(oracle, fam, spec, init, 30, C, rng=rng, path=4)
    assert g[0]["report"].path == 4
    assert_parity(g[0], o[0], RTOL)


def test_i8_table_selection_per_chain_block(oracle):
    """The hot loop's table is chosen per CTA (= block of 128 chains) from the chains' bound on |eta|: below the end of the
    bank-group-replicated cubic table (level 3; |eta| <= 10.74 with six slices, 11.15 with five) the two conflict-free loads, up to
    39.9 the 256-per-unit mean-corrected table, beyond that the exact table with the clamped epilogue.  Block 0: every chain
    parallel to the longest row of X with |eta| reaching 0.995 of the level-3 bound - the top of the replicated table is really
    indexed; block 1: bounds 2 % above 11.15; block 2: bounds ~30; block 3: a mix of all of them and one chain at |eta| ~ 300."""
    from fmcmc_b200 import ll_logistic
    rng = np.random.default_rng(77)
    n, p, C = 1300, 12, 4 * 128
    X = rng.standard_normal((n, p)) / np.sqrt(p); X[:, 0] = 1.0
    y = (rng.random(n) < 1 / (1 + np.exp(-X @ rng.standard_normal(p)))).astype(np.float64)
    fam = ll_logistic(X, y, prior_sd=2.0)
    norms = np.sqrt((X * X).sum(axis=1))
    imax, rmax, cmax = int(norms.argmax()), norms.max(), np.abs(X).max(axis=0)

    def with_bound(v, b):
        return v * b / min(np.abs(v) @ cmax, np.linalg.norm(v) * rmax)

    init = np.empty((C, p))
    for c in range(128):
        init[c] = with_bound(X[imax] * (1 if c % 2 else -1) + rng.normal(0, 1e-3, p), 0.995 * (694 / 64.0 - 0.1))
    for c in range(128, 256):
        init[c] = with_bound(rng.normal(0, 1.0, p), 1.02 * 11.15)
    for c in range(256, 384):
        init[c] = with_bound(rng.normal(0, 1.0, p), rng.uniform(25.0, 35.0))
    for c in range(384, 512):
        init[c] = with_bound(rng.normal(0, 1.0, p), (3.0, 10.7, 11.3, 39.0)[c % 4])
    init[500] = rng.normal(0, 120.0, p)
    spec = dict(type=A.KERNEL_NORMAL, k=p, mu=0.0, scale=0.002)
    g, o, _ = run_both(oracle, fam, spec, init, 20, C, rng=rng, path=4)
    assert g[0]["report"].path == 4
    eta = X @ init[:128].T
    assert 10.5 < np.abs(eta).max() < 694 / 64.0 - 0.1
    assert_parity(g[0], o[0], RTOL)


@pytest.mark.parametrize("family", ["logistic", "gaussian"])
def test_i8_tiny_problem(oracle, family):
    """Forced path 4 far below its intended regime: fewer observations than one 128-row tile, one covariate (31 of the 32
    K columns are padding), 3 chains in a 128-lane block."""
    rng = np.random.default_rng(5)
    n, p, C = 50, 1, 3
    if family == "logistic":
        from fmcmc_b200 import ll_logistic
        X = rng.standard_normal((n, p))
        y = (rng.random(n) < 1 / (1 + np.exp(-1.5 * X[:, 0]))).astype(np.float64)
        fam, k = ll_logistic(X, y, prior_sd=2.0), p
        init = rng.normal(0, 0.3, (C, k))
    else:
        fam, k = _gaussian_family(rng, n, p)
        init = np.c_[rng.normal(0, 0.3, (C, k - 1)), np.full(C, 2.0)]
    spec = dict(type=A.KERNEL_NORMAL, k=k, mu=0.0, scale=0.2)
    g, o, _ = run_both(oracle, fam, spec, init, 200, C, rng=rng, path=4)
    assert g[0]["report"].path == 4
    assert_parity(g[0], o[0], RTOL, family)
