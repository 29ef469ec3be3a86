"""Round 2, second session: the host-facing pieces around the hot path.

* streamed outputs in R's column-major layout (the layout INTEGRATION.md's .Call shim requests)
* the HBM copy of X / y cached on the family object (MCMC() uploads and packs once, like the data an R closure
  captures) and fmcmc_model_trim
* the stepping path chosen from the whole job's chain count (sharding must not change a chain's bits)
* bulks that restart from the last KEPT row when thin does not divide the bulk (R/mcmc.R:909-911, quirk D2)
* kernel-state save / restore across MCMC() calls (FmcmcKernel.load_state)
"""
import numpy as np
import pytest

import fmcmc_b200 as fm
from fmcmc_b200 import _abi as A
from fmcmc_b200.device import DeviceModel

pytestmark = pytest.mark.gpu


def _logistic(rng, n, p):
    X = rng.standard_normal((n, p)) / np.sqrt(p)
    X[:, 0] = 1.0
    y = (rng.random(n) < 1 / (1 + np.exp(-X @ rng.standard_normal(p)))).astype(float)
    return fm.ll_logistic(X, y)


@pytest.mark.parametrize("burnin,thin", [(0, 1), (37, 3)])
@pytest.mark.parametrize("path", [2, 3, 4])
def test_colmajor_streamed_outputs(path, burnin, thin):
    """FMCMC_RUN_COLMAJOR ([chain][param][row], an R matrix per chain) takes the streamed-output route as well: same values
    as the row-major call, which tests/test_gpu_parity.py::test_streamed_outputs_match_oracle pins on the oracle."""
    rng = np.random.default_rng(5)
    n, p, C, T = 2500, 8 if path == 2 else 20, 600, 400
    fam = _logistic(rng, n, p)
    spec = dict(type=A.KERNEL_NORMAL, k=p, mu=0.0, scale=0.05)
    init = rng.normal(0, 0.1, (C, p))
    m = DeviceModel(fam)
    m.set_path(path)
    try:
        r = m.run(spec, T, C, initial=init, burnin=burnin, thin=thin, stream=A.marshal_stream(A.STREAM_PHILOX, seed=3))
        c = m.run(spec, T, C, initial=init, burnin=burnin, thin=thin, stream=A.marshal_stream(A.STREAM_PHILOX, seed=3),
                  flags=A.RUN_COLMAJOR)
    finally:
        m.close()
    keep = (T - burnin) // thin
    assert r["report"].path == path and r["ans"].nbytes * 2 >= 8 << 20
    for name in ("ans", "draws"):
        cm = c[name].reshape(C, p, keep).transpose(0, 2, 1)        # the buffer holds [chain][param][row]
        assert np.array_equal(cm, r[name]), name
    assert np.array_equal(c["logpost"], r["logpost"])


def test_family_keeps_its_device_copy_between_calls():
    """MCMC() on the same family object reuses the resident X / y (and the int8 slices of path 4): one upload, same results."""
    rng = np.random.default_rng(9)
    fam = _logistic(rng, 70000, 24)
    init = rng.normal(0, 0.1, (160, 24))
    a = fm.MCMC(init, fam, 30, nchains=160, seed=5, kernel=fm.kernel_normal(scale=0.02))
    m1 = fam.device_model(0)
    b = fm.MCMC(init, fam, 30, nchains=160, seed=5, kernel=fm.kernel_normal(scale=0.02))
    assert fam.device_model(0) is m1 and m1._h
    assert np.array_equal(a.as_array(), b.as_array())
    assert fm.MCMC_OUTPUT.report.path == 4
    # a forced path on one call does not stick to the next one
    c = fm.MCMC(init, fam, 30, nchains=160, seed=5, kernel=fm.kernel_normal(scale=0.02), path=3)
    assert fm.MCMC_OUTPUT.report.path == 3
    d = fm.MCMC(init, fam, 30, nchains=160, seed=5, kernel=fm.kernel_normal(scale=0.02))
    assert fm.MCMC_OUTPUT.report.path == 4 and np.array_equal(d.as_array(), a.as_array())
    assert np.allclose(c.as_array(), a.as_array(), rtol=0, atol=1e-9)
    fam.release()
    assert not m1._h
    e = fm.MCMC(init, fam, 30, nchains=160, seed=5, kernel=fm.kernel_normal(scale=0.02))   # re-uploads on demand
    assert np.array_equal(e.as_array(), a.as_array())
    fam.release()


def test_model_trim_keeps_the_path_and_frees_the_rest():
    rng = np.random.default_rng(10)
    fam = _logistic(rng, 70000, 24)
    C, T = 160, 12
    init = rng.normal(0, 0.1, (C, 24))
    spec = dict(type=A.KERNEL_NORMAL, k=24, mu=0.0, scale=0.02)
    m = DeviceModel(fam)
    try:
        with pytest.raises(fm.FmcmcError, match="has not run"):
            m.trim(4)
        full = m.run(spec, T, C, initial=init, stream=A.marshal_stream(A.STREAM_PHILOX, seed=1))
        assert full["report"].path == 4
        m.trim(4)
        again = m.run(spec, T, C, initial=init, stream=A.marshal_stream(A.STREAM_PHILOX, seed=1))
        assert again["report"].path == 4 and np.array_equal(again["ans"], full["ans"])
        m.set_path(3)
        with pytest.raises(fm.FmcmcError, match="trimmed"):
            m.run(spec, T, C, initial=init, stream=A.marshal_stream(A.STREAM_PHILOX, seed=1))
        with pytest.raises(fm.FmcmcError, match="released"):
            m.logpost(init[:2])
        ram = dict(type=A.KERNEL_RAM, k=24)          # kernel_ram needs one more slice: X is gone, so this must be refused, not mis-run
        m.set_path(4)
        with pytest.raises(fm.FmcmcError, match="needs 6"):
            m.run(ram, T, C, initial=init, stream=A.marshal_stream(A.STREAM_PHILOX, seed=1))
    finally:
        m.close()


def test_path_follows_the_whole_jobs_chain_count():
    """64 chains of a 256-chain job (what one of 4 GPUs holds) run the path 256 chains run - and give the bits the un-sharded
    run gives for those chains (per-call path selection made results depend on the number of GPUs)."""
    rng = np.random.default_rng(12)
    fam = _logistic(rng, 70000, 24)
    C, T = 256, 10
    init = rng.normal(0, 0.1, (C, 24))
    spec = dict(type=A.KERNEL_NORMAL, k=24, mu=0.0, scale=0.02)
    m = DeviceModel(fam)
    try:
        whole = m.run(spec, T, C, initial=init, stream=A.marshal_stream(A.STREAM_PHILOX, seed=8))
        alone = m.run(spec, T, 64, initial=init[128:192], stream=A.marshal_stream(A.STREAM_PHILOX, seed=8), chain_offset=128)
        shard = m.run(spec, T, 64, initial=init[128:192], stream=A.marshal_stream(A.STREAM_PHILOX, seed=8), chain_offset=128,
                      nchains_total=C)
    finally:
        m.close()
    assert whole["report"].path == 4 and alone["report"].path == 3 and shard["report"].path == 4
    assert np.array_equal(shard["ans"], whole["ans"][128:192]) and np.array_equal(shard["logpost"], whole["logpost"][128:192])


@pytest.mark.parametrize("thin,burnin", [(3, 0), (7, 50), (4, 0)])
def test_bulks_restart_from_the_last_kept_row(oracle, readme_data, thin, burnin):
    """R/mcmc.R:909-911: bulk b + 1 starts from ans[niter(ans), ], the last KEPT row - an earlier row than the last one
    computed when thin does not divide the bulk.  Fed streams, two chains, three bulks of 200 against the oracle driven
    the way the reference's loop drives it."""
    from helpers import r_fed_stream, readme_model
    R = oracle.RRng
    R.set_seed(99)
    freq, nsteps, C = 200, 600 + burnin, 2
    bulks = [freq + burnin, freq, freq]
    feds, raw = [], []
    for b in bulks:
        logu, z = r_fed_stream(R, C, b, 3)
        raw.append((logu, z))
        feds.append(fm.FedStream(logu, z))
    never = fm.convergence_gelman(freq, threshold=1.0)            # R-hat < 1 never holds: all bulks run
    kern = fm.kernel_normal_reflective(scale=0.05, lb=[-5.0, 0.0, 0.0], ub=5.0)
    init = np.array([[0.0, 0.0, readme_data["sd_y"]], [1.0, 1.0, 3.0]])
    g = fm.MCMC(init, fm.ll_gaussian_lm(readme_data["X"], readme_data["y"], intercept=True, guard=False), nsteps,
                nchains=C, thin=thin, burnin=burnin, kernel=kern, conv_checker=never, fed=feds)
    spec = fm.kernel_normal_reflective(scale=0.05, lb=[-5.0, 0.0, 0.0], ub=5.0).to_spec(3)
    cur, rows = init, []
    for i, b in enumerate(bulks):
        o = oracle.run(readme_model(readme_data, guard=False), spec, cur, b, nchains=C, burnin=burnin if i == 0 else 0,
                       thin=thin, stream=A.marshal_stream(A.STREAM_FED, logu=raw[i][0], z=raw[i][1]), threads=2)
        rows.append(o["ans"])
        cur = o["ans"][:, -1, :]
    want = np.concatenate(rows, axis=1)
    got = g.as_array()
    assert got.shape == want.shape
    assert np.array_equal(np.any(got[:, 1:] != got[:, :-1], axis=2), np.any(want[:, 1:] != want[:, :-1], axis=2))
    assert np.max(np.abs(got - want) / np.abs(want).max(axis=(0, 1))) < 1e-12


def test_kernel_state_save_and_restore(readme_data):
    """A kernel's per-chain state survives outside the object: load_state() into a fresh kernel continues the adaptation
    exactly where the saved one stopped (the reference keeps it in the kernel environment, R/mcmc.R:629-631)."""
    ll = fm.ll_gaussian_lm(readme_data["X"], readme_data["y"], intercept=True, guard=True)
    C = 8
    init = np.tile([3.0, 2.0, 4.0], (C, 1))
    lb = [np.nan, np.nan, 0.0]
    k1 = fm.kernel_adapt(warmup=50, lb=lb)
    a = fm.MCMC(init, ll, 300, nchains=C, seed=3, kernel=k1)
    ist, dst = k1._istate.copy(), k1._dstate.copy()
    b = fm.MCMC(a, ll, 200, nchains=C, seed=4, kernel=k1)
    k2 = fm.kernel_adapt(warmup=50, lb=lb)
    k2.load_state(ist, dst, C, 3)
    assert k2.is_list and k2[0].abs_iter == 299
    c = fm.MCMC(a, ll, 200, nchains=C, seed=4, kernel=k2)
    assert np.array_equal(b.as_array(), c.as_array())
    ll.release()


@pytest.mark.parametrize("colmajor", [False, True])
@pytest.mark.parametrize("path,C,T", [(1, 40, 61), (3, 600, 300)])
def test_bulks_write_into_one_set_of_host_arrays(path, C, T, colmajor):
    """fmcmc_run_spec.out_rows_total / out_row_offset: three bulks fill ONE caller-owned array each for ans / draws / logpost
    (direct and streamed copy routes, both layouts) - the same numbers three separate calls return."""
    rng = np.random.default_rng(15)
    n, p = (300, 6) if path == 1 else (2500, 20)
    fam = _logistic(rng, n, p)
    spec = dict(type=A.KERNEL_ADAPT, k=p, mu=0.0, warmup=15, freq=1, eps=1e-4)
    init = rng.normal(0, 0.1, (C, p))
    fl = A.RUN_COLMAJOR if colmajor else 0
    dl = A.state_len(A.KERNEL_ADAPT, p, p)
    m = DeviceModel(fam)
    m.set_path(path)
    try:
        sep, ist, dst = [], np.zeros((C, A.ISTATE_LEN), dtype=np.int64), np.zeros((C, dl))
        for b in range(3):
            sep.append(m.run(spec, T, C, initial=init if b == 0 else None, istate=ist, dstate=dst, flags=fl,
                             burnin=7 if b == 0 else 0, thin=2, stream=A.marshal_stream(A.STREAM_PHILOX, seed=4, run_index=b)))
        keeps = [s_["logpost"].shape[1] for s_ in sep]
        R = sum(keeps) + 5                                       # a few spare rows: never written
        shape = (C, p, R) if colmajor else (C, R, p)
        ans, drw, lp = np.full(shape, -7.0), np.full(shape, -7.0), np.full((C, R), -7.0)
        ist, dst, off = np.zeros((C, A.ISTATE_LEN), dtype=np.int64), np.zeros((C, dl)), 0
        for b in range(3):
            o = m.run(spec, T, C, initial=init if b == 0 else None, istate=ist, dstate=dst, flags=fl, into=(ans, drw, lp, off),
                      burnin=7 if b == 0 else 0, thin=2, stream=A.marshal_stream(A.STREAM_PHILOX, seed=4, run_index=b))
            assert o["report"].path == path
            want = (lambda a: a.reshape(C, p, keeps[b])) if colmajor else (lambda a: a)   # a separate call returns the raw buffer
            assert np.array_equal(o["ans"], want(sep[b]["ans"])) and np.array_equal(o["draws"], want(sep[b]["draws"]))
            assert np.array_equal(o["logpost"], sep[b]["logpost"])
            off += keeps[b]
        tail = ans[:, :, off:] if colmajor else ans[:, off:]
        assert np.all(tail == -7.0) and np.all(lp[:, off:] == -7.0)
        with pytest.raises(fm.FmcmcError, match="out_row_offset"):
            m.run(spec, T, C, initial=init, flags=fl, into=(ans, drw, lp, R - 3), stream=A.marshal_stream(A.STREAM_PHILOX, seed=4))
    finally:
        m.close()


def test_mcmc_with_checker_returns_views_of_one_array(readme_data):
    """The bulk loop accumulates into one array: the mcmc.list, get_logpost() and get_draws() are views of it, the kernel
    state is fetched once at the end, and nothing differs from running the bulks by hand."""
    ll = fm.ll_gaussian_lm(readme_data["X"], readme_data["y"], intercept=True, guard=True)
    C = 6
    init = np.tile([3.0, 2.0, 4.0], (C, 1)) + np.linspace(0, 0.5, C)[:, None]
    kern = fm.kernel_adapt(warmup=40, lb=[np.nan, np.nan, 0.0])
    ans = fm.MCMC(init, ll, 900, nchains=C, seed=21, kernel=kern, conv_checker=fm.convergence_gelman(300, threshold=0.0))
    assert ans.niter() == 900 and ans.as_array().shape == (C, 900, 3)
    assert fm.get_logpost()[0].shape == (900,) and fm.get_draws()[2].shape == (900, 3)
    assert kern[0].abs_iter == 3 * 299 and kern[C - 1].Sigma.shape == (3, 3)
    # by hand: three runs of 300 rows, each restarting from the previous last row, one kernel object
    k2 = fm.kernel_adapt(warmup=40, lb=[np.nan, np.nan, 0.0])
    parts, cur = [], init
    m = ll.device_model(0)
    spec = k2.to_spec(3)
    ist, dst = np.zeros((C, A.ISTATE_LEN), dtype=np.int64), np.zeros((C, A.state_len(A.KERNEL_ADAPT, 3, 3)))
    for b in range(3):
        o = m.run(spec, 300, C, initial=cur, istate=ist, dstate=dst, stream=A.marshal_stream(A.STREAM_PHILOX, seed=21, run_index=b))
        parts.append(o["ans"])
        cur = o["ans"][:, -1, :]
    assert np.array_equal(ans.as_array(), np.concatenate(parts, axis=1))
    assert np.array_equal(kern[3].Sigma, dst[3, :9].reshape(3, 3, order="F")) and kern[3].abs_iter == ist[3, 0]
    ll.release()


def _ess_numpy(x):
    """Per-series Geyer initial-positive-sequence ESS from direct-lag autocovariances (divisor N).  x: [C][N][k]."""
    C, N, k = x.shape
    out = np.empty((C, k))
    for c in range(C):
        for a in range(k):
            v = x[c, :, a] - x[c, :, a].mean()
            g = np.array([np.dot(v[:N - l], v[l:]) / N for l in range(N)])
            if not g[0] > 0:
                out[c, a] = 0.0
                continue
            tau, m = -1.0, 0
            while 2 * m + 1 <= N - 1:
                pair = (g[2 * m] + g[2 * m + 1]) / g[0]
                if not pair > 0:
                    break
                tau += 2 * pair
                m += 1
            out[c, a] = N / max(tau, 1.0 / N)
    return out


def test_device_ess_matches_numpy(readme_data):
    """fmcmc_store_ess (autocovariances + Geyer's initial positive sequence per (chain, parameter) on the device) against the
    same estimator in numpy on the same samples; a fixed parameter (constant series) has ESS 0 and is masked out."""
    ll = fm.ll_gaussian_lm(readme_data["X"], readme_data["y"], intercept=True, guard=True)
    C, T = 12, 700
    spec = fm.kernel_normal_reflective(scale=[0.1, 0.1, 0.1], lb=[np.nan, np.nan, 0.0], fixed=[False, True, False]).to_spec(3)
    m = DeviceModel(ll)
    try:
        m.store_reset(C, T)
        o = m.run(spec, T, C, initial=np.tile([3.0, 2.0, 4.0], (C, 1)), flags=A.RUN_APPEND,
                  stream=A.marshal_stream(A.STREAM_PHILOX, seed=9))
        ess, trunc = m.store_ess(100, T, [1, 0, 1], C)
        want = _ess_numpy(o["ans"][:, 100:, :][:, :, [0, 2]])
        assert not trunc and ess.shape == (C, 2)
        np.testing.assert_allclose(ess, want, rtol=1e-9)
        assert np.all(ess > 5) and np.all(ess < T)               # a random-walk chain: well below one sample per row
        all3, _ = m.store_ess(100, T, [1, 1, 1], C)
        assert np.all(all3[:, 1] == 0.0)                         # the fixed column never moves
        short, tr2 = m.store_ess(100, T, [1, 0, 1], C, max_lag=3)
        assert tr2 and np.all(short >= ess - 1e-9)               # truncated sums over-estimate
        with pytest.raises(fm.FmcmcError, match="window"):
            m.store_ess(0, 3, [1, 0, 1], C)
    finally:
        m.close()
