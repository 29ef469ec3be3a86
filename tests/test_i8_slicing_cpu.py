"""The split-integer (Ozaki) scheme of path 4 (fmcmc_b200/csrc/tiled_i8.cuh), restated in numpy with exact integer
arithmetic — no GPU needed.  Checks the claims the kernel's header makes, for the default 8-bit digits (NS = 5 / 6) and for
round 1's 7-bit digits (NS = 6 / 7, -DI8_DIGIT_BITS=7):

  * the digits reconstruct x 2^-cexp (and theta 2^(cexp - eth)) to half a unit of the last digit, every digit inside int8
    ([-128, 127] after the carry pass / [-64, 64]);
  * the diagonal sums a_d fit int32 with the stated bounds, merged pairs fit int32 (int64 for the last pair of 8-bit NS = 6 at
    K = 128), the merged group fits int64;
  * eta reassembled from the NS kept diagonals differs from the float64 dot product by <= the stated bound,
    relative to the largest column contribution max_j |theta_j| 2^cexp_j;
  * y eta - softplus(eta) summed over observations == theta . X'(y - 1/2) - sum(|eta| / 2 + log1p(exp(-|eta|))).
"""
import numpy as np
import pytest


def exponent(m, db):
    """smallest e with m < 2^e (0 for m == 0); 8-bit digits: with m 2^-e < 127 / 128 (i8_exponent())"""
    m = np.asarray(m, dtype=np.float64)
    e = np.where(m > 0, np.frexp(m)[1], 0).astype(np.int64)        # frexp: m = f 2^e, f in [0.5, 1)  ->  m < 2^e
    if db == 8:
        e = np.where((m > 0) & ~(np.ldexp(m, -e) < 127.0 / 128.0), e + 1, e)
    return e


def slices(u, ns, db):
    """u -> integer digits d_s with u ~ sum_s d_s 2^(-(db - 1) - db s) (i8_slices())"""
    out = []
    r = u * 2.0 ** (db - 1)
    for _ in range(ns):
        q = np.rint(r)
        out.append(q.astype(np.int64))
        r = (r - q) * 2.0 ** db
    if db == 8:                                               # balanced digits: +128 (+129) -> -128 (-127) and a carry upwards
        for i in range(ns - 1, 0, -1):
            over = out[i] >= 128
            out[i] = np.where(over, out[i] - 256, out[i])
            out[i - 1] = out[i - 1] + over
    return out


@pytest.mark.parametrize("db,ns", [(8, 5), (8, 6), (7, 6), (7, 7)])
@pytest.mark.parametrize("badly_scaled", [False, True])
def test_slices_reconstruct_and_bounds(db, ns, badly_scaled):
    rng = np.random.default_rng(ns + db)
    n, p, C = 400, 32, 24
    X = rng.standard_normal((n, p))
    X[:, 0] = 1.0
    X[3, 5] = np.abs(X[:, 5]).max() * 2.0 ** 0.999            # a column whose maximum sits just below a power of two
    X[5:9, 7] = [0.5, -0.5, 0.498046875, 127.0 / 256]         # exact ties of the first remainder
    th = rng.standard_normal((C, p))
    if badly_scaled:
        sc = np.logspace(-6, 3, p)
        X = X * sc
        th = th / sc
    cexp = exponent(np.abs(X).max(axis=0), db)
    Xs = np.ldexp(X, -cexp)
    lim = 127.0 / 128.0 if db == 8 else 1.0
    assert np.all(np.abs(Xs) < lim)
    tp = np.ldexp(th, cexp)                                  # theta'_j = theta_j 2^cexp_j
    eth = exponent(np.abs(tp).max(axis=1), db)
    ts = np.ldexp(tp, -eth[:, None])
    sx, st = slices(Xs, ns, db), slices(ts, ns, db)
    lo, hi = (-128, 127) if db == 8 else (-64, 64)
    for s in sx + st:
        assert s.min() >= lo and s.max() <= hi               # int8
    rx = sum(s * 2.0 ** (-(db - 1) - db * i) for i, s in enumerate(sx))
    assert np.max(np.abs(rx - Xs)) <= 2.0 ** (-(db - 1) - db * (ns - 1) - 1)   # half a unit of the last digit
    # diagonals, exactly, in Python integers
    a = [sum(sx[i] @ st[d - i].T for i in range(d + 1)) for d in range(ns)]          # [n][C] int64 each
    unit = p * (1 << (2 * db - 2))
    for d in range(ns):
        assert np.abs(a[d]).max() <= (d + 1) * unit
    v = [a[2 * q] * (1 << db) + a[2 * q + 1] for q in range(ns // 2)]
    for q, x in enumerate(v):                                # merged pairs fit int32 - except the last pair of 8-bit NS = 6 at K = 128
        assert np.abs(x).max() < 2 ** 31
        bound = (2 * q + 1) * 4 * unit * (1 << db) + (2 * q + 2) * 4 * unit          # the same pair at K = 128 (4 K blocks)
        assert bound < 2 ** 31 or (db == 8 and ns == 6 and q == 2)
    t = np.zeros((n, C), dtype=object)
    for d in range(ns):
        t = t + a[d].astype(object) * (1 << (db * (ns - 1 - d)))
    assert max(abs(int(x)) for x in t.ravel()) < 2 ** 63
    shift = 2 * (db - 1) + db * (ns - 1)
    eta = np.array(t, dtype=np.float64) * np.ldexp(1.0, (eth - shift).astype(int))[None, :]
    ref = X @ th.T
    scale = np.abs(tp).max(axis=1)[None, :]                  # largest column contribution bound (|x'| < 1)
    err = np.max(np.abs(eta - ref) / scale)
    assert err <= (ns + 1) * p * 2.0 ** (-db * ns - 2) * (4 if db == 8 else 1), err   # dropped diagonals d >= NS
    assert err <= {(8, 5): 2e-10, (8, 6): 8e-13, (7, 6): 3e-11, (7, 7): 3e-13}[(db, ns)]


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
