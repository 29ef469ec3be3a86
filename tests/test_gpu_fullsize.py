"""BASELINE.json's FULL sizes, checked through size-independent properties (the CPU oracle needs minutes
per chain-step at these sizes, so it is not the checker here):

  * split additivity   f(full data) == f(first half) + f(second half) (+ the prior counted once)
  * permutation invariance of the observations
  * agreement of four independent device code paths (DFMA lane<->chain kernel, DMMA kernel, split-integer
    tcgen05 kernel, the plain one-CTA-per-theta logpost kernel)
  * bit-for-bit determinism, and independence of the result from how chains are sharded

Tolerance: 1e-12 relative on every log-posterior (north_star's FP64 band); decisions identical."""
import numpy as np
import pytest

from fmcmc_b200 import _abi as A

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _first_rows(model, spec, init, C, seed=5, rows=2, path=0, chain_offset=0):
    if path:
        model.set_path(path)
    out = model.run(spec, rows, C, initial=init, stream=A.marshal_stream(A.STREAM_PHILOX, seed=seed, run_index=0),
                    chain_offset=chain_offset)
    return out


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


@pytest.fixture(scope="module")
def cfg3_data():
    import bench
    return bench.make_data()                        # n = 1e6, p = 32: BASELINE configs[2]


def test_cfg3_logistic_full_size_properties(cfg3_data):
    import fmcmc_b200 as fm
    from fmcmc_b200.device import DeviceModel
    X, y = cfg3_data
    n, p = X.shape
    C = 1024
    rng = np.random.default_rng(3)
    init = rng.normal(0, 0.3, (C, p))
    spec = dict(type=A.KERNEL_NORMAL, k=p, mu=0.0, scale=0.01)
    prior = (init ** 2).sum(axis=1) / 8.0           # sum(beta^2) / (2 * 2^2)

    full = DeviceModel(fm.ll_logistic(X, y))
    o3 = _first_rows(full, spec, init, C, path=3)
    o2 = _first_rows(full, spec, init, C, path=2)
    o4 = _first_rows(full, spec, init, C, path=4)   # the default for > 128 chains (auto-selection: cfg5 test below)
    assert o3["report"].path == 3 and o2["report"].path == 2 and o4["report"].path == 4
    f3, f2 = o3["logpost"], o2["logpost"]           # [C][2]: f(initial), f(first proposal)
    assert _rel(f3, f2) <= RTOL                     # DMMA kernel == DFMA kernel
    assert _rel(o4["logpost"], f3) <= RTOL          # tcgen05 int8-slice kernel == DMMA kernel
    assert np.array_equal(o3["draws"], o2["draws"]) and np.array_equal(o4["draws"], o3["draws"])  # same Philox proposals
    # path 4: per-chain exponents and a fixed slice order => independent of how chains are grouped into CTAs / calls
    lo4 = _first_rows(full, spec, init[:300], 300, path=4)
    hi4 = _first_rows(full, spec, init[300:], C - 300, path=4, chain_offset=300)
    assert np.array_equal(np.concatenate([lo4["logpost"], hi4["logpost"]]), o4["logpost"])
    assert np.array_equal(np.concatenate([lo4["ans"], hi4["ans"]]), o4["ans"])
    direct = full.logpost(init[:64])                # third code path: logpost_kernel
    assert _rel(f3[:64, 0], direct) <= RTOL

    # determinism + sharding independence (Philox keyed by the global chain id)
    again = _first_rows(full, spec, init, C, path=3, rows=4)
    lo = _first_rows(full, spec, init[:512], 512, path=3, rows=4)
    hi = _first_rows(full, spec, init[512:], 512, path=3, rows=4, chain_offset=512)
    assert np.array_equal(again["ans"][:, :2], o3["ans"]) and np.array_equal(again["logpost"][:, :2], f3)
    assert np.array_equal(np.concatenate([lo["ans"], hi["ans"]]), again["ans"])
    assert np.array_equal(np.concatenate([lo["logpost"], hi["logpost"]]), again["logpost"])   # same summation tree
    # a shard small enough for the observation-split mapping sums in another (still fixed) order: same decisions
    # and samples, log-posteriors equal to the last ulps
    few = _first_rows(full, spec, init[100:108], 8, path=3, rows=4, chain_offset=100)
    assert np.array_equal(few["ans"], again["ans"][100:108])
    assert _rel(few["logpost"], again["logpost"][100:108]) <= 1e-14
    full.close()

    # split additivity: LL_A + LL_B - prior == (LL_A - prior) + (LL_B - prior) + prior
    h = n // 2 + 77                                  # ragged halves
    a = DeviceModel(fm.ll_logistic(X[:h], y[:h]))
    fa = _first_rows(a, spec, init, C, path=3)["logpost"][:, 0]
    a.close()
    b = DeviceModel(fm.ll_logistic(X[h:], y[h:]))
    fb = _first_rows(b, spec, init, C, path=3)["logpost"][:, 0]
    b.close()
    assert _rel(fa + fb + prior, f3[:, 0]) <= RTOL

    # permutation invariance
    perm = rng.permutation(n)
    pm = DeviceModel(fm.ll_logistic(np.asfortranarray(X[perm]), y[perm]))
    fp = _first_rows(pm, spec, init, C, path=3)["logpost"][:, 0]
    pm.close()
    assert _rel(fp, f3[:, 0]) <= RTOL


def test_cfg3_decisions_agree_between_device_kernels(cfg3_data):
    """20 MH rows of 1024 kernel_adapt chains at full n: the DFMA, DMMA and tcgen05 kernels make the same accept/reject
    decisions from the same Philox streams (a flipped decision would show as an O(1) difference)."""
    import fmcmc_b200 as fm
    from fmcmc_b200.device import DeviceModel
    X, y = cfg3_data
    p = X.shape[1]
    C, T = 1024, 21
    init = np.random.default_rng(4).normal(0, 0.1, (C, p))
    spec = dict(type=A.KERNEL_ADAPT, k=p, mu=0.0, warmup=5, freq=1, eps=1e-4)
    outs = []
    for path in (2, 3, 4):
        m = DeviceModel(fm.ll_logistic(X, y))
        outs.append(_first_rows(m, spec, init, C, rows=T, path=path))
        m.close()
    acc = [np.any(o["ans"][:, 1:] != o["ans"][:, :-1], axis=2) for o in outs]
    assert np.array_equal(acc[0], acc[1]) and np.array_equal(acc[0], acc[2])
    assert acc[0].mean() > 0.05
    scale = np.abs(outs[0]["ans"]).max(axis=(0, 1))
    for other in (1, 2):
        assert np.max(np.abs(outs[0]["ans"] - outs[other]["ans"]).max(axis=(0, 1)) / scale) <= 1e-11
        assert _rel(outs[other]["logpost"], outs[0]["logpost"]) <= RTOL


def test_cfg5_gaussian_full_size_properties():
    """BASELINE configs[4] shape: n = 1e7, 127 columns + sd, generated directly in HBM (10 GB).  512 of the
    GPU's 8192 chains keep the test short; the kernel, tile geometry and per-chain arithmetic are the same."""
    torch = pytest.importorskip("torch")
    from fmcmc_b200.device import DeviceModel
    from fmcmc_b200.families import DeviceFamily
    n, p, C = 10_000_000, 127, 512
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    Xd = torch.empty((p, n), dtype=torch.float64, device="cuda")
    Xd[0].fill_(1.0)
    for j in range(1, p):
        Xd[j].normal_(generator=g)
    beta = torch.randn(p, dtype=torch.float64, device="cuda", generator=g)
    yd = torch.matmul(beta, Xd) + 2.0 * torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    torch.cuda.synchronize()
    k = p + 1
    rng = np.random.default_rng(8)
    init = np.c_[beta.cpu().numpy() + rng.normal(0, 1e-3, (C, p)), rng.uniform(1.5, 2.5, C)]
    lb = np.full(k, -A.DBL_MAX); lb[-1] = 0.0
    spec = dict(type=A.KERNEL_NORMAL_REFLECTIVE, k=k, mu=0.0, scale=1e-4, lb=lb, ub=A.DBL_MAX)

    def model_of(Xt, yt):
        fam = DeviceFamily(A.FAMILY_GAUSSIAN_LM, yt.numel(), p_x=p, flags=A.MODEL_GUARD)
        return DeviceModel(fam, device_ptrs=(Xt.data_ptr(), yt.data_ptr(), None))

    full = model_of(Xd, yd)
    o = _first_rows(full, spec, init, C)
    assert o["report"].path == 4                    # > 128 chains: the split-integer tcgen05 kernel is the default
    f = o["logpost"][:, 0]
    again = _first_rows(full, spec, init, C)
    assert np.array_equal(again["logpost"], o["logpost"]) and np.array_equal(again["ans"], o["ans"])
    o3 = _first_rows(full, spec, init, C, path=3)   # the FP64 DMMA kernel on the same inputs
    assert o3["report"].path == 3
    assert np.array_equal(o3["ans"], o["ans"]) and _rel(o["logpost"], o3["logpost"]) <= RTOL
    full.close()
    # closed form from the sufficient statistics, FP64 on the device via torch (plumbing only)
    th = torch.from_numpy(init).cuda()
    XtX = Xd @ Xd.T
    Xty = Xd @ yd
    yty = yd @ yd
    bt = th[:, :p]
    ss = yty - 2.0 * (bt @ Xty) + ((bt @ XtX) * bt).sum(dim=1)
    sd = th[:, p]
    ref = -(n * (0.918938533204672741780329736406 + torch.log(sd)) + 0.5 * ss / sd ** 2)
    assert _rel(f, ref.cpu().numpy()) <= 1e-9       # the normal-equations form cancels ~3 digits; streaming is the truth

    h = n // 2                                       # even split keeps the 16-byte column alignment of borrowed pointers
    Xa, Xb = Xd[:, :h].contiguous(), Xd[:, h:].contiguous()
    ya, yb = yd[:h].contiguous(), yd[h:].contiguous()
    del Xd
    a = model_of(Xa, ya)
    fa = _first_rows(a, spec, init, C)["logpost"][:, 0]
    a.close()
    b = model_of(Xb, yb)
    fb = _first_rows(b, spec, init, C)["logpost"][:, 0]
    b.close()
    assert _rel(fa + fb, f) <= RTOL


def test_cfg3_full_n_against_the_oracle(oracle, cfg3_data):
    """BASELINE configs[2] at its full n = 1e6, p = 32 against the CPU oracle itself (not only device paths against each
    other): 160 chains (more than one 128-chain block, so the default split-integer tcgen05 path runs, 6 slices) x 5 fed-stream
    rows - every decision identical, log-posteriors and samples within 1e-12.  800 oracle chain-steps over 1e6 observations:
    seconds on the host threads."""
    import fmcmc_b200 as fm
    from gpu_util import assert_parity, run_both
    X, y = cfg3_data
    p = X.shape[1]
    C, T = 160, 5
    rng = np.random.default_rng(17)
    init = rng.normal(0, 0.1, (C, p))
    spec = dict(type=A.KERNEL_NORMAL, k=p, mu=0.0, scale=0.004)
    g, o, _ = run_both(oracle, fm.ll_logistic(X, y), spec, init, T, C, rng=rng)
    assert g[0]["report"].path == 4
    assert_parity(g[0], o[0], RTOL, "cfg3 full n")
    acc = np.any(g[0]["ans"][:, 1:] != g[0]["ans"][:, :-1], axis=2)
    assert 0.02 < acc.mean() < 0.98                 # the comparison saw both accepted and rejected proposals


def test_cfg3_full_n_near_the_mode_against_the_oracle(oracle, cfg3_data):
    """The same at the posterior mode, where the bench's chains end up: |eta| reaches 4.7 and the chains' bound on it 10.4 - the
    upper part of the bank-group-replicated cubic table (|eta| <= 11.15) is what the hot loop reads - with kernel_adapt adapting
    on every row (warm-up 3), 192 chains (two chain blocks: two observation slices per CTA) x 24 fed-stream rows = 4 416 oracle
    chain-steps over 1e6 observations: every decision identical, samples and log-posteriors within 1e-12."""
    import bench
    import fmcmc_b200 as fm
    from gpu_util import assert_parity, run_both
    X, y = cfg3_data
    p = X.shape[1]
    C, T = 192, 24
    beta = np.random.Generator(np.random.PCG64(bench.DATA_SEED + 7919)).standard_normal(p)   # the generating beta* (bench.make_data)
    rng = np.random.default_rng(29)
    init = beta + rng.normal(0, 2e-3, (C, p))
    spec = dict(type=A.KERNEL_ADAPT, k=p, mu=0.0, warmup=3, freq=1, eps=1e-6)
    g, o, _ = run_both(oracle, fm.ll_logistic(X, y, prior_sd=2.0), spec, init, T, C, rng=rng)
    assert g[0]["report"].path == 4
    assert_parity(g[0], o[0], RTOL, "cfg3 full n, mode, adapt")
    acc = np.any(g[0]["ans"][:, 1:] != g[0]["ans"][:, :-1], axis=2)
    assert 0.02 < acc.mean() < 0.98
    bound = np.linalg.norm(g[0]["ans"][:, -1, :], axis=1) * np.sqrt((X * X).sum(axis=1).max())
    assert 9.5 < bound.max() < 11.15                 # inside the replicated table, near its end


def test_cfg5_shape_against_the_oracle(oracle):
    """BASELINE configs[4]'s geometry (Gaussian, 127 columns + sd: four K blocks, Theta slices split between tensor and
    shared memory, 6 slices) at n = 300 000 against the CPU oracle: decisions identical, samples and log-posteriors 1e-12."""
    import fmcmc_b200 as fm
    from gpu_util import assert_parity, run_both
    rng = np.random.default_rng(23)
    n, p, C, T = 300_000, 127, 160, 4
    X = np.empty((n, p), order="F")
    X[:, 0] = 1.0
    for j in range(1, p):
        X[:, j] = rng.standard_normal(n)
    beta = rng.standard_normal(p)
    y = X @ beta + 2.0 * rng.standard_normal(n)
    fam = fm.ll_gaussian_lm(X, y, intercept=False, guard=True)
    k = p + 1
    lb = np.full(k, -A.DBL_MAX); lb[-1] = 0.0
    spec = dict(type=A.KERNEL_NORMAL_REFLECTIVE, k=k, mu=0.0, scale=2e-4, lb=lb, ub=A.DBL_MAX)
    init = np.c_[beta + rng.normal(0, 1e-3, (C, p)), rng.uniform(1.5, 2.5, C)]
    g, o, _ = run_both(oracle, fam, spec, init, T, C, rng=rng)
    assert g[0]["report"].path == 4
    assert_parity(g[0], o[0], RTOL, "cfg5 shape")
