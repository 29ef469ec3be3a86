"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same fed streams.

Contract (BASELINE.json north_star): every accept/reject decision identical, every sample
within 1e-12 relative in FP64."""
import numpy as np
import pytest

from fmcmc_b200 import _abi as A
from gpu_util import assert_parity, run_both
from helpers import r_fed_stream

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _readme_family(d, guard=True):
    from fmcmc_b200 import ll_gaussian_lm
    return ll_gaussian_lm(d["X"], d["y"], intercept=True, guard=guard)


def _kernels(k):
    lb = np.full(k, -A.DBL_MAX); lb[-1] = 0.0
    return {
        "normal": dict(type=A.KERNEL_NORMAL, k=k, mu=0.0, scale=0.1),
        "normal_reflective": dict(type=A.KERNEL_NORMAL_REFLECTIVE, k=k, mu=0.0, scale=0.3, lb=lb, ub=6.0),
        "unif": dict(type=A.KERNEL_UNIF, k=k, min_=-0.2, max_=0.2),
        "unif_reflective": dict(type=A.KERNEL_UNIF_REFLECTIVE, k=k, min_=-0.5, max_=0.5, lb=lb, ub=6.0),
        "adapt": dict(type=A.KERNEL_ADAPT, k=k, mu=0.0, warmup=50, freq=1, eps=1e-4, lb=lb, ub=A.DBL_MAX),
        "adapt_freq3": dict(type=A.KERNEL_ADAPT, k=k, mu=0.0, warmup=40, freq=3, eps=1e-4, lb=lb, ub=A.DBL_MAX),
        "adapt_bw30": dict(type=A.KERNEL_ADAPT, k=k, mu=0.0, warmup=45, freq=2, bw=30, eps=1e-4, lb=lb, ub=A.DBL_MAX),
        "ram": dict(type=A.KERNEL_RAM, k=k, warmup=0, freq=1, eps=1e-2, arate=0.234, lb=lb, ub=A.DBL_MAX),
        "nmirror": dict(type=A.KERNEL_NMIRROR, k=k, mu=0.0, scale=0.5, warmup=100, arate=0.4, lb=lb, ub=A.DBL_MAX,
                        nadapt=np.array([25, 50, 75, 100])),
        "umirror": dict(type=A.KERNEL_UMIRROR, k=k, mu=0.0, scale=0.5, warmup=100, arate=0.4, lb=lb, ub=A.DBL_MAX,
                        nadapt=np.array([25, 50, 75, 100])),
        "normal_ordered": dict(type=A.KERNEL_NORMAL, k=k, mu=0.0, scale=0.2, scheme=A.SCHEME_ORDERED),
        "normal_fixed": dict(type=A.KERNEL_NORMAL_REFLECTIVE, k=k, mu=0.0, scale=0.2, lb=lb, ub=9.0,
                             fixed=[False, True] + [False] * (k - 2)),
        "unif_explicit": dict(type=A.KERNEL_UNIF, k=k, min_=-0.3, max_=0.3, scheme=A.SCHEME_EXPLICIT,
                              order=np.arange(k, 0, -1)),
    }


@pytest.mark.parametrize("name", list(_kernels(3)))
@pytest.mark.parametrize("nchains", [1, 4, 320])          # CTA-per-chain, CTA-per-chain, warp-per-chain
def test_fed_parity_readme_model(oracle, readme_data, name, nchains):
    spec = _kernels(3)[name]
    rng = np.random.default_rng(11)
    init = np.tile([1.0, 1.0, readme_data["sd_y"]], (nchains, 1)) + rng.normal(0, 0.05, (nchains, 3))
    T = 400 if nchains < 100 else 160
    g, o, st = run_both(oracle, _readme_family(readme_data), spec, init, T, nchains, rng=rng)
    assert_parity(g[0], o[0], RTOL, name)
    assert np.array_equal(st[0], st[2]), "integer kernel state differs"
    np.testing.assert_allclose(st[1], st[3], rtol=1e-10, atol=1e-12 * max(np.abs(st[3]).max(), 1e-300))


def test_readme_golden_through_cuda(oracle, readme_data):
    """Config 1: the README's seed-1215 run (README.md:161-201) replayed on the GPU from R's own streams."""
    R = oracle.RRng
    R.set_seed(1215)
    T = 5000
    logu, z = r_fed_stream(R, 1, T, 3)
    from fmcmc_b200.device import DeviceModel
    model = DeviceModel(_readme_family(readme_data))
    out = model.run(dict(type=A.KERNEL_NORMAL, k=3, mu=0.0, scale=1.0), T, 1,
                    initial=[0, 0, readme_data["sd_y"]], stream=A.marshal_stream(A.STREAM_FED, logu=logu, z=z))
    model.close()
    a = out["ans"][0]
    assert [float(f"{v:.4g}") for v in a.mean(0)] == [3.113, 1.975, 4.093]
    assert [float(f"{v:.4g}") for v in a.std(0, ddof=1)] == [0.1759, 0.1065, 0.07843]
    assert out["report"].path == 1


def test_bulks_carry_state(oracle, readme_data):
    """Restart from the device-resident last state + kernel state across bulks (R/mcmc.R:901-947)."""
    spec = _kernels(3)["adapt"]
    rng = np.random.default_rng(5)
    g, o, st = run_both(oracle, _readme_family(readme_data), spec, [1.0, 1.0, 4.0], 90, 3, rng=rng, bulks=3)
    for b in range(3):
        assert_parity(g[b], o[b], RTOL, f"bulk {b}")
    assert np.array_equal(st[0][:, 0], st[2][:, 0]) and st[0][0, 0] == 3 * 89


def test_burnin_thin(oracle, readme_data):
    spec = _kernels(3)["normal"]
    rng = np.random.default_rng(6)
    g, o, _ = run_both(oracle, _readme_family(readme_data), spec, [1.0, 1.0, 4.0], 500, 2, rng=rng, burnin=100, thin=7)
    assert g[0]["ans"].shape == (2, 57, 3)
    assert g[0]["report"].first_iter == 107 and g[0]["report"].last_iter == 100 + 57 * 7
    assert_parity(g[0], o[0], RTOL)


def _logistic_family(rng, n, p):
    from fmcmc_b200 import ll_logistic
    X = rng.standard_normal((n, p)) / np.sqrt(p)
    X[:, 0] = 1.0
    beta = rng.standard_normal(p)
    y = (rng.random(n) < 1 / (1 + np.exp(-X @ beta))).astype(np.float64)
    return ll_logistic(X, y, prior_sd=2.0)


@pytest.mark.parametrize("kname", ["normal", "adapt", "ram", "nmirror"])
def test_logistic_resident(oracle, kname):
    rng = np.random.default_rng(21)
    fam = _logistic_family(rng, 777, 5)
    spec = dict(_kernels(5)[kname])
    for key in ("lb", "ub"):
        if key in spec:
            spec[key] = -A.DBL_MAX if key == "lb" else A.DBL_MAX
    spec["scale"] = 0.15
    g, o, _ = run_both(oracle, fam, spec, rng.normal(0, 0.1, (6, 5)), 300, 6, rng=rng, path=1)
    assert_parity(g[0], o[0], RTOL, kname)


def test_hier_normal_resident(oracle):
    """playground/hierarchical-bayes.Rmd:28-51: N=1000, Nc=20, unit variances, gamma ~ U(-1, 1)."""
    from fmcmc_b200 import ll_hier_normal
    rng = np.random.default_rng(12315)
    N, Nc = 1000, 20
    group = np.arange(N) % Nc
    theta = rng.normal(0.3, 1, Nc)
    y = rng.normal(theta[group], 1.0)
    fam = ll_hier_normal(y, group, n_groups=Nc, gamma_bounds=(-1, 1))
    k = Nc + 1
    spec = dict(type=A.KERNEL_NORMAL_REFLECTIVE, k=k, mu=0.0, scale=0.05,
                lb=np.r_[np.full(Nc, -A.DBL_MAX), -1.0], ub=np.r_[np.full(Nc, A.DBL_MAX), 1.0])
    g, o, _ = run_both(oracle, fam, spec, np.zeros(k), 300, 5, rng=rng)
    assert_parity(g[0], o[0], RTOL)
    fam2 = ll_hier_normal(y * 3 + 70, group % 4, n_groups=4, gamma_bounds=(0, 150), estimate_scales=True)
    spec2 = dict(type=A.KERNEL_RAM, k=7, warmup=0, freq=1, eps=1e-2, arate=0.234,
                 lb=np.r_[np.full(5, -A.DBL_MAX), 1e-3, 1e-3], ub=A.DBL_MAX)
    g, o, _ = run_both(oracle, fam2, spec2, [70, 70, 70, 70, 70, 3, 3], 300, 300, rng=rng)
    assert_parity(g[0], o[0], 1e-11, "hier + ram")


@pytest.mark.parametrize("path", [2, 3])
@pytest.mark.parametrize("family", ["logistic", "gaussian"])
@pytest.mark.parametrize("kname", ["normal", "normal_reflective", "adapt", "ram", "nmirror", "unif"])
def test_tiled_path_parity(oracle, family, kname, path):
    """Paths 2 / 3 (observation-tiled TMA pipeline; DFMA lane<->chain kernel / DMMA kernel) forced on a
    problem the oracle finishes in seconds: ragged n (tail tile, odd n -> padded ld), chains not a
    multiple of the chain block."""
    rng = np.random.default_rng(33)
    n, p = 2 * 128 * 3 + 77, 7
    if family == "logistic":
        fam, k = _logistic_family(rng, n, p), p
        init = rng.normal(0, 0.1, (70, k))
    else:
        from fmcmc_b200 import ll_gaussian_lm
        X = rng.standard_normal((n, p))
        y = 1.0 + X @ rng.standard_normal(p) + rng.normal(0, 2.0, n)
        fam, k = ll_gaussian_lm(X, y, intercept=True, guard=True), p + 2
        init = np.c_[rng.normal(0, 0.1, (70, k - 1)), np.full(70, 3.0)]
    spec = dict(_kernels(k)[kname])
    if family == "logistic":
        for key in ("lb", "ub"):
            if key in spec:
                spec[key] = -A.DBL_MAX if key == "lb" else A.DBL_MAX
    if "scale" in spec:
        spec["scale"] = 0.05
    g, o, _ = run_both(oracle, fam, spec, init, 140, 70, rng=rng, path=path)
    assert g[0]["report"].path == path
    assert_parity(g[0], o[0], RTOL, f"{family}/{kname}")


@pytest.mark.parametrize("path", [2, 3])
def test_tiled_many_chain_blocks(oracle, path):
    """> 512 chains => several chain blocks in the tiled grid; p_x = 32 (the bench's register tier)."""
    rng = np.random.default_rng(44)
    fam = _logistic_family(rng, 1500, 32)
    spec = dict(type=A.KERNEL_NORMAL, k=32, mu=0.0, scale=0.03)
    g, o, _ = run_both(oracle, fam, spec, rng.normal(0, 0.1, (600, 32)), 40, 600, rng=rng, path=path)
    assert_parity(g[0], o[0], RTOL)


@pytest.mark.parametrize("family,p", [("gaussian", 127), ("gaussian", 50), ("logistic", 100), ("logistic", 33)])
def test_tiled_wide_design_matrix(oracle, family, p):
    """p_x > 32 (config 5 has k = 128): only the DMMA kernel handles it; Theta stays in registers as B
    fragments, the stage holds 64 / 32 observations.  Ragged n, chains not a multiple of the chain block."""
    rng = np.random.default_rng(55)
    n, C = 1000 + 13, 150
    if family == "logistic":
        fam, k = _logistic_family(rng, n, p), p
        init = rng.normal(0, 0.05, (C, k))
        spec = dict(type=A.KERNEL_NORMAL, k=k, mu=0.0, scale=0.02)
    else:
        from fmcmc_b200 import ll_gaussian_lm
        X = rng.standard_normal((n, p))
        y = 1.0 + X @ rng.standard_normal(p) + rng.normal(0, 2.0, n)
        fam, k = ll_gaussian_lm(X, y, intercept=True, guard=True), p + 2
        init = np.c_[rng.normal(0, 0.1, (C, k - 1)), np.full(C, 3.0)]
        lb = np.full(k, -A.DBL_MAX); lb[-1] = 0.0
        spec = dict(type=A.KERNEL_NMIRROR, k=k, mu=0.0, scale=0.05, warmup=30, arate=0.4, lb=lb, ub=A.DBL_MAX,
                    nadapt=np.array([10, 20, 30]))
    g, o, _ = run_both(oracle, fam, spec, init, 60, C, rng=rng, path=3)
    assert g[0]["report"].path == 3
    assert_parity(g[0], o[0], RTOL, f"{family}/p={p}")


@pytest.mark.parametrize("family,p,C", [("logistic", 32, 1), ("logistic", 9, 5), ("logistic", 20, 12), ("gaussian", 6, 3),
                                        ("gaussian", 100, 2), ("logistic", 127, 16), ("gaussian", 60, 30)])
def test_tiled_few_chains(oracle, family, p, C):
    """Few chains on a large n (the reference's typical usage): the DMMA kernel's observation-split mapping —
    all warps share the chains and split the observations of each stage; cross-warp reduction in fixed order."""
    rng = np.random.default_rng(66)
    n = 3000 + 41
    if family == "logistic":
        fam, k = _logistic_family(rng, n, p), p
        init = rng.normal(0, 0.05, (C, k))
        spec = dict(type=A.KERNEL_ADAPT, k=k, mu=0.0, warmup=15, freq=1, eps=1e-4)
    else:
        from fmcmc_b200 import ll_gaussian_lm
        X = rng.standard_normal((n, p))
        y = 1.0 + X @ rng.standard_normal(p) + rng.normal(0, 2.0, n)
        fam, k = ll_gaussian_lm(X, y, intercept=True, guard=True), p + 2
        init = np.c_[rng.normal(0, 0.1, (C, k - 1)), np.full(C, 3.0)]
        lb = np.full(k, -A.DBL_MAX); lb[-1] = 0.0
        spec = dict(type=A.KERNEL_RAM, k=k, warmup=0, freq=1, eps=1e-3, arate=0.234, lb=lb, ub=A.DBL_MAX)
    g, o, _ = run_both(oracle, fam, spec, init, 50, C, rng=rng, path=3)
    assert g[0]["report"].path == 3
    assert_parity(g[0], o[0], RTOL, f"{family}/p={p}/C={C}")


@pytest.mark.parametrize("kname", ["normal", "adapt", "ram", "nmirror", "normal_reflective"])
@pytest.mark.parametrize("path", [1, 2, 3])
def test_philox_stream_matches_oracle(oracle, readme_data, kname, path):
    """Production streams: the device Philox4x32-10 + AS241 inversion is the oracle's, so whole
    runs agree (up to libm ulps in log/qnorm tails; a flipped decision would show up as O(1) error)."""
    rng = np.random.default_rng(3)
    fam = _logistic_family(rng, 900, 6)
    spec = dict(_kernels(6)[kname])
    for key in ("lb", "ub"):
        if key in spec:
            spec[key] = -A.DBL_MAX if key == "lb" else A.DBL_MAX
    spec["scale"] = 0.1
    g, o, _ = run_both(oracle, fam, spec, np.zeros(6), 150, 40, path=path, philox_seed=20260317, chain_offset=1000)
    assert_parity(g[0], o[0], 1e-9, kname)


def test_philox_independent_of_sharding(oracle):
    """Chains keyed by GLOBAL chain id: running chains [0,8) at once == [0,4) and [4,8) separately."""
    from fmcmc_b200.device import DeviceModel
    rng = np.random.default_rng(9)
    fam = _logistic_family(rng, 600, 4)
    spec = dict(type=A.KERNEL_NORMAL, k=4, mu=0.0, scale=0.2)
    m = DeviceModel(fam)
    st = lambda: A.marshal_stream(A.STREAM_PHILOX, seed=77, run_index=0)
    full = m.run(spec, 100, 8, initial=np.zeros(4), stream=st())["ans"]
    lo = m.run(spec, 100, 4, initial=np.zeros(4), stream=st(), chain_offset=0)["ans"]
    hi = m.run(spec, 100, 4, initial=np.zeros(4), stream=st(), chain_offset=4)["ans"]
    m.close()
    assert np.array_equal(full[:4], lo) and np.array_equal(full[4:], hi)


def test_logpost_matches_oracle(oracle, readme_data):
    from fmcmc_b200.device import DeviceModel
    rng = np.random.default_rng(2)
    fam = _readme_family(readme_data)
    th = np.c_[rng.normal(2, 1, (50, 2)), rng.uniform(0.5, 6, 50)]
    th[3, 2] = -1.0       # sd < 0 -> NaN -> guarded -> -Inf
    th[4, 2] = 0.0
    m = DeviceModel(fam)
    got = m.logpost(th)
    m.close()
    desc = fam.marshal()
    ref = np.array([oracle.logpost(desc, t) for t in th])
    assert np.array_equal(np.isfinite(got), np.isfinite(ref))
    f = np.isfinite(ref)
    np.testing.assert_allclose(got[f], ref[f], rtol=1e-13)
    assert np.array_equal(got[~f], ref[~f])


@pytest.mark.parametrize("burnin,thin", [(0, 1), (37, 3)])
def test_streamed_outputs_match_oracle(oracle, burnin, thin):
    """Outputs above 8 MB leave the device chunk by chunk while later rows are still being computed (fmcmc_run, tiled
    paths): same rows, same order, burnin / thin included (R/mcmc.R:786-813), checked against the oracle."""
    rng = np.random.default_rng(71)
    n, p, C, T = 2500, 8, 600, 400
    fam = _logistic_family(rng, n, p)
    spec = dict(type=A.KERNEL_NORMAL, k=p, mu=0.0, scale=0.05)
    g, o, _ = run_both(oracle, fam, spec, rng.normal(0, 0.1, (C, p)), T, C, rng=rng, burnin=burnin, thin=thin)
    keep = (T - burnin) // thin
    assert g[0]["report"].path in (2, 3, 4) and g[0]["ans"].shape == (C, keep, p)
    assert g[0]["ans"].nbytes * 2 >= 8 << 20                     # large enough to take the streamed path
    assert_parity(g[0], o[0], RTOL, f"streamed outputs burnin={burnin} thin={thin}")
