"""The split-integer (Ozaki) scheme of path 4 (fmcmc_b200/csrc/tiled_i8.cuh), restated in numpy with exact integer
arithmetic — no GPU needed.  Checks the claims the kernel's header makes:

  * the slices reconstruct x 2^-cexp (and theta 2^(cexp - eth)) to 2^(-7 NS + 1), every slice in [-64, 64] (int8);
  * the diagonal sums a_d fit int32 with the stated bounds, merged pairs fit int32, the merged group fits int64;
  * eta reassembled from the NS kept diagonals differs from the float64 dot product by <= the stated bound,
    relative to the largest column contribution max_j |theta_j| 2^cexp_j;
  * y eta - softplus(eta) summed over observations == theta . X'(y - 1/2) - sum(|eta| / 2 + log1p(exp(-|eta|))).
"""
import numpy as np
import pytest


def exponent(m):
    """smallest e with m < 2^e (0 for m == 0), as i8_exponent()"""
    m = np.asarray(m, dtype=np.float64)
    e = np.where(m > 0, np.frexp(m)[1], 0)          # frexp: m = f 2^e, f in [0.5, 1)  ->  m < 2^e
    return e.astype(np.int64)


def slices(u, ns):
    """|u| <= 1 -> integer slices s with u ~ sum_s s 2^(-6 - 7 s) (i8_slices())"""
    out = []
    r = u * 64.0
    for _ in range(ns):
        q = np.rint(r)
        out.append(q.astype(np.int64))
        r = (r - q) * 128.0
    return out


@pytest.mark.parametrize("ns", [6, 7])
@pytest.mark.parametrize("badly_scaled", [False, True])
def test_slices_reconstruct_and_bounds(ns, badly_scaled):
    rng = np.random.default_rng(ns)
    n, p, C = 400, 32, 24
    X = rng.standard_normal((n, p))
    X[:, 0] = 1.0
    th = rng.standard_normal((C, p))
    if badly_scaled:
        sc = np.logspace(-6, 3, p)
        X = X * sc
        th = th / sc
    cexp = exponent(np.abs(X).max(axis=0))
    Xs = np.ldexp(X, -cexp)                                  # (-1, 1)
    assert np.all(np.abs(Xs) < 1)
    tp = np.ldexp(th, cexp)                                  # theta'_j = theta_j 2^cexp_j
    eth = exponent(np.abs(tp).max(axis=1))
    ts = np.ldexp(tp, -eth[:, None])
    sx, st = slices(Xs, ns), slices(ts, ns)
    for s in sx + st:
        assert s.min() >= -64 and s.max() <= 64              # int8 with room to spare
    rx = sum(s * 2.0 ** (-6 - 7 * i) for i, s in enumerate(sx))
    assert np.max(np.abs(rx - Xs)) <= 2.0 ** (-7 * ns + 1)
    # diagonals, exactly, in Python integers
    a = [sum(sx[i] @ st[d - i].T for i in range(d + 1)) for d in range(ns)]          # [n][C] int64 each
    for d in range(ns):
        assert np.abs(a[d]).max() <= (d + 1) * p * 64 * 64 < 2 ** 24
    v = [a[2 * q] * 128 + a[2 * q + 1] for q in range(ns // 2)]
    assert all(np.abs(x).max() < 2 ** 31 for x in v)         # merged pairs fit int32
    t = np.zeros((n, C), dtype=object)
    for d in range(ns):
        t = t + a[d].astype(object) * (1 << (7 * (ns - 1 - d)))
    if ns <= 6:
        assert max(abs(int(x)) for x in t.ravel()) < 2 ** 63
    eta = np.array(t, dtype=np.float64) * np.ldexp(1.0, (eth - (12 + 7 * (ns - 1))).astype(int))[None, :]
    ref = X @ th.T
    scale = np.abs(tp).max(axis=1)[None, :]                  # largest column contribution bound (|x'| < 1)
    err = np.max(np.abs(eta - ref) / scale)
    assert err <= (ns + 1) * p * 2.0 ** (-7 * ns - 2), err   # dropped diagonals d >= NS: <= (d+1) p 2^(-7 d - 2) each
    assert err <= {6: 3e-11, 7: 3e-13}[ns]


def test_logistic_even_identity():
    rng = np.random.default_rng(1)
    n, p = 5000, 8
    X = rng.standard_normal((n, p))
    th = rng.standard_normal(p)
    y = (rng.random(n) < 0.4).astype(np.float64)
    eta = X @ th
    direct = np.sum(np.where(y == 1, -np.logaddexp(0, -eta), -np.logaddexp(0, eta)))
    sxy = (y - 0.5) @ X
    even = th @ sxy - np.sum(0.5 * np.abs(eta) + np.log1p(np.exp(-np.abs(eta))))
    assert abs(even - direct) <= 1e-12 * abs(direct)
