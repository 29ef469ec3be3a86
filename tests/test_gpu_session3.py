"""Round 2, third session: the stepping-path variants that must not change a single bit.

* the few-chain head of kernel_adapt (one CTA per chain, right-looking Cholesky in registers: tiled.cuh
  tiled_head_adapt_cta_kernel) against the warp-per-chain head (left-looking loop, chol_lower_warp): every entry of the factor
  sees the same subtractions in the same order, so samples, log-posteriors and the kernel state are IDENTICAL - checked with the
  production Philox streams over the warm-up / first adapted row / steady adaptation, for k_f = 32, k_f < 32 with fixed
  parameters and bounds, freq > 1, and across two calls (state carried, factor cached);
* programmatic dependent launch on / off (the likelihood kernel's set-up overlapping the head kernel): identical outputs on
  every tiled path;
* and the head against the oracle with fed streams at 1 / 3 chains (the existing parity matrix runs 1 / 4 / 320 chains too);
* path 4 with one, two or four observation slices per CTA (FMCMC_I8_GSL): every slice is flushed into its own row of the partial
  sums whatever the CTA that walked it, so the outputs are identical bit for bit - logistic (three table levels) and Gaussian.
"""
import os

import numpy as np
import pytest

from fmcmc_b200 import _abi as A
from gpu_util import assert_parity, run_both
from test_gpu_parity import _logistic_family

pytestmark = pytest.mark.gpu


def _run(fam, spec, init, T, C, env, path=0, calls=1, seed=99):
    """The same Philox-stream job under an environment override read at model creation (FMCMC_HEAD_CTA / FMCMC_PDL)."""
    from fmcmc_b200.device import DeviceModel
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        m = DeviceModel(fam)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    k = spec["k"]
    kf = int((~np.broadcast_to(np.asarray(spec.get("fixed", False), dtype=bool), (k,))).sum())
    dlen = A.state_len(spec["type"], k, kf)
    ist, dst = np.zeros((C, A.ISTATE_LEN), dtype=np.int64), np.zeros((C, max(dlen, 1)))
    outs = []
    try:
        if path:
            m.set_path(path)
        for b in range(calls):
            g = m.run(spec, T, C, initial=init if b == 0 else None, stream=A.marshal_stream(A.STREAM_PHILOX, seed=seed, run_index=b),
                      istate=ist, dstate=dst if dlen else None)
            outs.append({n: g[n].copy() for n in ("ans", "draws", "logpost")} | {"path": g["report"].path})
    finally:
        m.close()
    return outs, ist.copy(), dst.copy()


def _same(a, b):
    (oa, ia, da), (ob, ib, db) = a, b
    for x, y in zip(oa, ob):
        assert x["path"] == y["path"]
        for n in ("ans", "draws", "logpost"):
            assert np.array_equal(x[n], y[n]), f"{n} differs: max |d| = {np.abs(x[n] - y[n]).max():.3e}"
    assert np.array_equal(ia, ib)
    assert np.array_equal(da, db), f"kernel state differs: max |d| = {np.abs(da - db).max():.3e}"


@pytest.mark.parametrize("case", ["k32", "k7_fixed_bounds", "freq3", "one_chain", "c140", "c300"])
def test_cta_head_is_bit_identical_to_the_warp_head(case):
    """(c140: nearly one CTA of chains per SM; c300: more chains than SMs, where both settings run the warp-per-chain head)"""
    rng = np.random.default_rng(5)
    p = 32 if case in ("k32", "freq3", "one_chain", "c140", "c300") else 7
    C = 1 if case == "one_chain" else (300 if case == "c300" else (140 if case == "c140" else 5))
    fam = _logistic_family(rng, 6000, p)
    spec = dict(type=A.KERNEL_ADAPT, k=p, mu=0.0, warmup=12, freq=3 if case == "freq3" else 1, eps=1e-4)
    if case == "k7_fixed_bounds":
        fixed = np.zeros(p, dtype=bool); fixed[2] = True
        spec.update(fixed=fixed, lb=np.full(p, -0.6), ub=np.full(p, 0.7))
    init = rng.normal(0, 0.1, (C, p))
    # (freq = 3: a continued call would adapt at its row 3 from rows (0, 1, 2) - the reference indexes row 0 there and both heads
    # refuse with FMCMC_EUNSUP - so that case is one longer call)
    calls, T = (1, 90) if case == "freq3" else (2, 45)
    a = _run(fam, spec, init, T, C, {"FMCMC_HEAD_CTA": "0"}, calls=calls)
    b = _run(fam, spec, init, T, C, {"FMCMC_HEAD_CTA": "1"}, calls=calls)
    assert a[0][0]["path"] in (2, 3, 4)
    assert a[1][0, 0] > spec["warmup"] + 40            # well past the warm-up: the covariance recurrence and the factorisation ran
    _same(a, b)


@pytest.mark.parametrize("path,C", [(2, 6), (3, 6), (3, 200), (4, 200)])
def test_programmatic_dependent_launch_changes_nothing(path, C):
    rng = np.random.default_rng(6)
    p = 16 if path == 2 else 32
    fam = _logistic_family(rng, 5000, p)
    spec = dict(type=A.KERNEL_ADAPT, k=p, mu=0.0, warmup=5, freq=1, eps=1e-4)
    init = rng.normal(0, 0.1, (C, p))
    a = _run(fam, spec, init, 40, C, {"FMCMC_PDL": "0"}, path=path)
    b = _run(fam, spec, init, 40, C, {"FMCMC_PDL": "1"}, path=path)
    assert a[0][0]["path"] == path
    _same(a, b)


@pytest.mark.parametrize("C", [1, 3])
def test_cta_head_against_the_oracle(oracle, C):
    """Fed streams, k_f = 32, adaptation from row 8 on: every decision identical, samples within 1e-12 (the head's arithmetic
    is bit-exact; the band is the likelihood kernels')."""
    rng = np.random.default_rng(7)
    p = 32
    fam = _logistic_family(rng, 3000, p)
    spec = dict(type=A.KERNEL_ADAPT, k=p, mu=0.0, warmup=6, freq=1, eps=1e-4)
    g, o, st = run_both(oracle, fam, spec, rng.normal(0, 0.1, (C, p)), 60, C, rng=rng, path=3, bulks=2)
    for gb, ob in zip(g, o):
        assert gb["report"].path == 3
        assert_parity(gb, ob, 1e-12, f"cta head C={C}")
    assert np.array_equal(st[0], st[2])
    assert np.allclose(st[1], st[3], rtol=1e-10, atol=1e-300)


@pytest.mark.parametrize("family,C", [("logistic", 512), ("logistic", 300), ("gaussian", 256)])
def test_slices_per_cta_change_nothing(family, C):
    """512 chains = 4 chain blocks (automatic: four slices per CTA), 300 = 3 blocks (one slice per CTA whatever is asked),
    256 = 2 blocks (two).  n is large enough for 148 slices with several tiles each, and not a multiple of the tile."""
    rng = np.random.default_rng(8)
    n, p = 148 * 128 * 3 + 77, 32
    if family == "logistic":
        fam = _logistic_family(rng, n, p)
        k = p
        init = rng.normal(0, 0.1, (C, k))
        init[5] *= 60.0          # one chain block beyond the replicated table, one chain beyond every table
        init[200] *= 400.0
    else:
        from fmcmc_b200 import ll_gaussian_lm
        X = rng.standard_normal((n, p))
        y = 1.0 + X @ rng.standard_normal(p) + rng.normal(0, 2.0, n)
        fam, k = ll_gaussian_lm(X, y, intercept=True, guard=True), p + 2
        init = np.c_[rng.normal(0, 0.1, (C, k - 1)), np.full(C, 3.0)]
    spec = dict(type=A.KERNEL_NORMAL, k=k, mu=0.0, scale=0.01)
    runs = [_run(fam, spec, init, 6, C, {"FMCMC_I8_GSL": g}, path=4) for g in ("1", "2", "4")]
    assert runs[0][0][0]["path"] == 4
    _same(runs[0], runs[1])
    _same(runs[0], runs[2])
