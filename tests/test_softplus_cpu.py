"""csrc/softplus.h (the logistic epilogue's branch-free log(1+exp(-a))) compiled for the HOST and checked
against mpmath: the algorithm and its coefficients are validated without a GPU."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

mpmath = pytest.importorskip("mpmath")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ulp_err(got, a):
    from mpmath import exp, log1p, mp, mpf
    mp.dps = 60
    out = []
    for g, x in zip(got, a):
        r = log1p(exp(-mpf(float(x))))
        u = np.spacing(abs(float(r)))
        out.append(abs(float((mpf(float(g)) - r) / mpf(float(u)))))
    return np.array(out)


def sample_points(rng):
    return np.concatenate([rng.uniform(0, 40, 4000), rng.uniform(0, 2, 4000), 10 ** rng.uniform(-300, 2.8, 1500),
                           [0.0, 708.0, 1e-17, 36.7, 37.0, 0.34657, 0.34658, 745.0, 1e300]])


def test_softplus_host_build_vs_mpmath():
    src = '#include "%s/fmcmc_b200/csrc/softplus.h"\n' % ROOT + \
          'extern "C" void sp_eval(const double* a, double* o, long n) { for (long i = 0; i < n; i++) o[i] = fm_softplus_neg(a[i]); }\n'
    with tempfile.TemporaryDirectory() as td:
        cpp, so = os.path.join(td, "sp.cpp"), os.path.join(td, "libsp.so")
        open(cpp, "w").write(src)
        subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, cpp], check=True)
        L = C.CDLL(so)
        a = sample_points(np.random.default_rng(0))
        o = np.empty_like(a)
        L.sp_eval(a.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), C.c_long(a.size))
    e = ulp_err(o, np.minimum(a, 708.0))
    assert e.max() < 2.5 and e.mean() < 0.6, (e.max(), e.mean())


def test_softplus_fine_table_vs_mpmath():
    """The 128-per-unit table + degree-4 Taylor core of the split-integer kernel's epilogue (softplus.h, FM_SP4_*)."""
    src = '#include "%s/fmcmc_b200/csrc/softplus.h"\n' % ROOT + \
          'extern "C" void sp4_eval(const double* a, double* o, long n) { static double tab[2 * FM_SP4_ENTRIES]; ' \
          'fm_softplus_table4_fill(tab); for (long i = 0; i < n; i++) o[i] = fm_softplus_tab4(a[i], tab); }\n'
    with tempfile.TemporaryDirectory() as td:
        cpp, so = os.path.join(td, "sp4.cpp"), os.path.join(td, "libsp4.so")
        open(cpp, "w").write(src)
        subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, cpp], check=True)
        L = C.CDLL(so)
        rng = np.random.default_rng(3)
        a = np.concatenate([rng.uniform(0, 39.9, 6000), rng.uniform(0, 2, 3000), 10 ** rng.uniform(-300, 1.5, 1000),
                            np.arange(0, 5120) / 128.0 + 1.0 / 256.0 - 1e-9, [0.0, 1e-17, 36.7, 39.99]])
        o = np.empty_like(a)
        L.sp4_eval(a.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), C.c_long(a.size))
        e = ulp_err(o, a)
        assert e.max() < 1.5 and e.mean() < 0.45, (e.max(), e.mean())
        big = np.array([40.0, 41.5, 64.0, 700.0, 1e300, np.inf])            # clamped: absolute error below 4.3e-18
        ob = np.empty_like(big)
        L.sp4_eval(big.ctypes.data_as(C.c_void_p), ob.ctypes.data_as(C.c_void_p), C.c_long(big.size))
        assert np.all(np.abs(ob - np.log1p(np.exp(-big))) < 4.3e-18)


def test_softplus_finest_table_vs_mpmath():
    """The 256-per-unit table + cubic near-minimax cores (softplus.h, FM_SP8_*) of the split-integer kernel at p <= 64."""
    src = '#include "%s/fmcmc_b200/csrc/softplus.h"\n' % ROOT + \
          'extern "C" void sp8_eval(const double* a, double* o, long n) { static double tab[2 * FM_SP8_ENTRIES]; ' \
          'fm_softplus_table8_fill(tab); for (long i = 0; i < n; i++) o[i] = fm_softplus_tab8(a[i], tab); }\n'
    with tempfile.TemporaryDirectory() as td:
        cpp, so = os.path.join(td, "sp8.cpp"), os.path.join(td, "libsp8.so")
        open(cpp, "w").write(src)
        subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, cpp], check=True)
        L = C.CDLL(so)
        rng = np.random.default_rng(4)
        a = np.concatenate([rng.uniform(0, 39.9, 6000), rng.uniform(0, 2, 3000), 10 ** rng.uniform(-300, 1.5, 1000),
                            np.arange(0, 10240) / 256.0 + 1.0 / 512.0 - 1e-9, np.arange(0, 10240) / 256.0 - 1.0 / 512.0 + 1e-9,
                            [0.0, 1e-17, 36.7, 39.99]])
        a = np.abs(a)
        o = np.empty_like(a)
        L.sp8_eval(a.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), C.c_long(a.size))
        e = ulp_err(o, a)
        assert e.max() < 2.0 and e.mean() < 0.5, (e.max(), e.mean())


def test_lcosh_table_vs_mpmath():
    """log(2 cosh(a / 2)) = a / 2 + log1p(exp(-a)) on the 256-per-unit (tau, T) table with the degree-4 Taylor core whose
    coefficients are polynomials in tau (softplus.h, fm_lcosh_*): the binary-logistic epilogue of the split-integer kernel.
    Absolute error against mpmath below 4e-16 for a < 2 (values ~0.7 .. 1.1), below 1.5 ulp of the value everywhere on [0, 40]."""
    from mpmath import exp, log1p, mp, mpf
    mp.dps = 60
    src = '#include "%s/fmcmc_b200/csrc/softplus.h"\n' % ROOT + \
          'extern "C" void lc_eval(const double* a, double* o, long n) { static double tab[2 * FM_SP8_ENTRIES]; ' \
          'fm_lcosh_table8_fill(tab); for (long i = 0; i < n; i++) o[i] = fm_lcosh_tab8(a[i], tab); }\n' \
          'extern "C" double lc_tau_last() { static double tab[2 * FM_SP8_ENTRIES]; fm_lcosh_table8_fill(tab); ' \
          'return tab[2 * (FM_SP8_ENTRIES - 1)]; }\n'
    with tempfile.TemporaryDirectory() as td:
        cpp, so = os.path.join(td, "lc.cpp"), os.path.join(td, "liblc.so")
        open(cpp, "w").write(src)
        subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, cpp], check=True)
        L = C.CDLL(so)
        L.lc_tau_last.restype = C.c_double
        assert L.lc_tau_last() == 0.5          # u = 1/4 - tau^2 = 0 exactly at the last entry
        rng = np.random.default_rng(5)
        a = np.concatenate([rng.uniform(0, 39.9, 6000), rng.uniform(0, 2, 3000), 10 ** rng.uniform(-300, 1.5, 1000),
                            np.arange(0, 10240) / 256.0 + 1.0 / 512.0 - 1e-9,
                            np.abs(np.arange(0, 10240) / 256.0 - 1.0 / 512.0 + 1e-9), [0.0, 1e-17, 36.7, 39.99, 40.0]])
        o = np.empty_like(a)
        L.lc_eval(a.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), C.c_long(a.size))
    err = np.empty_like(a)
    ulp = np.empty_like(a)
    for i, (g, x) in enumerate(zip(o, a)):
        r = mpf(float(x)) / 2 + log1p(exp(-mpf(float(x))))
        err[i] = float(mpf(float(g)) - r)
        ulp[i] = abs(err[i]) / np.spacing(float(r))
    assert ulp.max() < 1.5 and ulp.mean() < 0.45, (ulp.max(), ulp.mean())
    assert np.abs(err[a < 2]).max() < 4e-16


def test_lcosh_degree3_on_the_mean_corrected_table_vs_mpmath():
    """The hot loop's degree-3 core (softplus.h, fm_lcosh_tab8m_*): what it leaves out, u (1/24 - u/4) d^4, is even in the
    remainder d and its mean over d is folded into the table's T.  Against mpmath: every evaluation within 8e-14 (the bound at
    a = 0, |d| = 1/512), and - what a log-posterior sees - the error of a SUM over uniformly spread arguments is zero-mean:
    |mean error| < 1e-15 per evaluation on every unit interval that matters, i.e. < 1.5e-15 relative to the values summed."""
    from mpmath import exp, log1p, mp, mpf
    mp.dps = 50
    src = '#include "%s/fmcmc_b200/csrc/softplus.h"\n' % ROOT + \
          'extern "C" void lc3_eval(const double* a, double* o, long n) { static double tab[2 * FM_SP8_ENTRIES]; ' \
          'fm_lcosh_table8m_fill(tab); for (long i = 0; i < n; i++) o[i] = fm_lcosh_tab8m(a[i], tab); }\n'
    rng = np.random.default_rng(11)
    with tempfile.TemporaryDirectory() as td:
        cpp, so = os.path.join(td, "lc3.cpp"), os.path.join(td, "liblc3.so")
        open(cpp, "w").write(src)
        subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, cpp], check=True)
        L = C.CDLL(so)
        # uniformly spread arguments per unit interval (the remainder d is then uniform on its cell), plus the cell edges
        a = np.concatenate([rng.uniform(lo, lo + 1, 4000) for lo in (0, 1, 2, 4, 8, 16, 32)] +
                           [np.arange(0, 2560) / 256.0 + 1.0 / 512.0 - 1e-12, [0.0, 39.9]])
        o = np.empty_like(a)
        L.lc3_eval(a.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), C.c_long(a.size))
    err = np.array([float(mpf(float(g)) - (mpf(float(x)) / 2 + log1p(exp(-mpf(float(x)))))) for g, x in zip(o, a)])
    assert np.abs(err).max() < 8e-14, np.abs(err).max()
    for i in range(7):
        seg = err[4000 * i:4000 * (i + 1)]
        assert abs(seg.mean()) < 1e-15, (i, seg.mean())
        assert seg.std() < 2.5e-14, (i, seg.std())
    # a log-likelihood-sized sum: 28 000 terms of ~0.7 .. 16, error of the sum relative to the sum
    assert abs(err[:28000].sum()) / o[:28000].sum() < 1e-16


def test_lcosh_replicated_cubic_table_vs_mpmath():
    """The bank-group-replicated cubic table of the hot loop (softplus.h, fm_lcosh_table6r_fill / fm_lcosh_tab6r): 64 points per
    unit up to a = 11.25, every point 256 bytes - eight identical copies of (c1, c0), then eight of (c2, c3) - holding the
    LEAST-SQUARES cubic of h on its cell.  Against mpmath: every evaluation within 4.6e-12 (the cell edges at a = 0), rms
    <= 1.6e-12, and - what a log-posterior sees - the error is orthogonal to constants on every cell: |mean| < 3e-14 on every
    unit interval, a log-likelihood-sized sum within 5e-15 relative."""
    from mpmath import exp, log1p, mp, mpf
    mp.dps = 50
    src = '#include "%s/fmcmc_b200/csrc/softplus.h"\n' % ROOT + \
          'static double tab[FM_LC6_ENTRIES_MAX * FM_LC6_POINT_BYTES / 8];\n' \
          'extern "C" const double* lc6_table() { fm_lcosh_table6r_fill(tab); return tab; }\n' \
          'extern "C" void lc6_eval(const double* a, double* o, long n, int r) { fm_lcosh_table6r_fill(tab); ' \
          'for (long i = 0; i < n; i++) o[i] = fm_lcosh_tab6r(a[i], tab, r < 0 ? (int)(i & 7) : r); }\n' \
          'extern "C" int lc6_entries() { return FM_LC6_ENTRIES_MAX; }\n'
    rng = np.random.default_rng(12)
    with tempfile.TemporaryDirectory() as td:
        cpp, so = os.path.join(td, "lc6.cpp"), os.path.join(td, "liblc6.so")
        open(cpp, "w").write(src)
        subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, cpp], check=True)
        L = C.CDLL(so)
        L.lc6_table.restype = C.POINTER(C.c_double)
        ne = L.lc6_entries()
        tab = np.ctypeslib.as_array(L.lc6_table(), shape=(ne, 2, 8, 2)).copy()
        amax = (ne - 1) / 64.0
        a = np.concatenate([rng.uniform(lo, lo + 1, 4000) for lo in (0, 1, 2, 4, 8, 10)] +
                           [np.arange(0, ne - 1) / 64.0 + 1.0 / 128.0 - 1e-12, [0.0, amax, amax - 1e-9]])
        o = np.empty_like(a)
        L.lc6_eval(a.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), C.c_long(a.size), C.c_int(-1))
        o3 = np.empty_like(a)
        L.lc6_eval(a.ctypes.data_as(C.c_void_p), o3.ctypes.data_as(C.c_void_p), C.c_long(a.size), C.c_int(3))
    assert amax == 11.25
    for r in range(1, 8):                              # the copies are identical: which one a lane reads cannot matter
        assert np.array_equal(tab[:, :, r, :], tab[:, :, 0, :])
    assert np.array_equal(o, o3)
    err = np.array([float(mpf(float(g)) - (mpf(float(x)) / 2 + log1p(exp(-mpf(float(x)))))) for g, x in zip(o, a)])
    assert np.abs(err).max() < 4.6e-12, np.abs(err).max()
    for i in range(6):
        seg = err[4000 * i:4000 * (i + 1)]
        assert abs(seg.mean()) < 3e-14, (i, seg.mean())
        assert seg.std() < 1.6e-12, (i, seg.std())
    assert abs(err[:24000].sum()) / o[:24000].sum() < 5e-15
    # the fit really is the least-squares one: on a dense uniform grid of ONE cell the error is orthogonal to 1, d, d^2, d^3
    k = 40
    d = (np.arange(2000) + 0.5) / 2000 * (1 / 64.0) - 1 / 128.0
    c1, c0 = tab[k, 0, 0]
    c2, c3 = tab[k, 1, 0]
    e = np.array([float(mpf(c0) + mpf(float(x)) * (mpf(c1) + mpf(float(x)) * (mpf(c2) + mpf(float(x)) * mpf(c3)))
                        - ((mpf(k) / 64 + mpf(float(x))) / 2 + log1p(exp(-(mpf(k) / 64 + mpf(float(x))))))) for x in d])
    scale = np.abs(e).max()
    for pw in range(4):
        assert abs(np.mean(e * (d * 128) ** pw)) < 2e-3 * scale, (pw, np.mean(e * (d * 128) ** pw), scale)


@pytest.mark.gpu
def test_softplus_device_vs_mpmath():
    import fmcmc_b200 as fm
    a = sample_points(np.random.default_rng(1))
    o = np.empty_like(a)
    err = C.create_string_buffer(256)
    dp = C.POINTER(C.c_double)
    rc = fm.lib().fmcmc_test_softplus(0, a.size, a.ctypes.data_as(dp), o.ctypes.data_as(dp), err, 256)
    assert rc == 0, err.value
    e = ulp_err(o, np.minimum(a, 708.0))
    assert e.max() < 2.5 and e.mean() < 0.6, (e.max(), e.mean())
