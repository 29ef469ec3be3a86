"""Committed golden vectors (tests/golden/cases.npz, made by tests/golden/make_golden.py from the oracle
after it was pinned on the reference's published outputs; tests/golden/lifeexpect.npz is the reference's
own data fixture, data/lifeexpect.rda).

CPU: the oracle still reproduces every vector bit for bit (guards the checker itself).
GPU: the CUDA path, through the C ABI, reproduces them — every accept/reject decision identical, every
sample within 1e-12 — without the oracle or /root/reference being present at run time."""
import os

import numpy as np
import pytest

from fmcmc_b200 import _abi as A

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL = 1e-12


def _load():
    z = np.load(os.path.join(HERE, "golden", "cases.npz"), allow_pickle=False)
    cases = {}
    for name in z["__names__"]:
        name = str(name)
        pre = name + "/"
        c = dict(fam={}, spec={}, out={})
        for key in z.files:
            if not key.startswith(pre):
                continue
            rest = key[len(pre):]
            v = z[key]
            if rest.startswith("fam/"):
                c["fam"][rest[4:]] = v
            elif rest.startswith("spec/"):
                c["spec"][rest[5:]] = v if v.ndim else v.item()
            elif rest.startswith("out/"):
                c["out"][rest[4:]] = v
            else:
                c[rest] = v if v.ndim else v.item()
        cases[name] = c
    return cases


CASES = _load()


def _family(c):
    import fmcmc_b200 as fm
    kw = {a: (v if v.ndim else v.item()) for a, v in c["fam"].items()}
    if "gamma_bounds" in kw:
        kw["gamma_bounds"] = tuple(float(x) for x in kw["gamma_bounds"])
    return {"gaussian_lm": fm.ll_gaussian_lm, "logistic": fm.ll_logistic, "hier_normal": fm.ll_hier_normal}[c["family"]](**kw)


def _free(spec):
    k = spec["k"]
    return np.where(~np.broadcast_to(np.asarray(spec.get("fixed", False), dtype=bool), (k,)))[0]


def test_lifeexpect_fixture_matches_reference_doc():
    """R/data.R:1-37, data-raw/lifeexpect.R:3-28: 1000 rows; smoke / female are 0-1 indicators;
    age = -10 smoke + 5.4 female + N(80.1, 2^2)."""
    le = np.load(os.path.join(HERE, "golden", "lifeexpect.npz"))
    assert le["age"].shape == le["smoke"].shape == le["female"].shape == (1000,)
    assert set(np.unique(le["smoke"])) == {0, 1} and set(np.unique(le["female"])) == {0, 1}
    assert abs(le["age"].mean() - 77.864) < 1e-3 and abs(le["age"].std(ddof=1) - 5.899) < 1e-3
    X = np.c_[np.ones(1000), le["smoke"], le["female"]]
    coef = np.linalg.lstsq(X, le["age"], rcond=None)[0]
    assert np.allclose(coef, [80.1, -10.0, 5.4], atol=0.3)
    assert abs(np.std(le["age"] - X @ coef, ddof=3) - 2.0) < 0.1


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(oracle, name):
    c = CASES[name]
    out = oracle.run(_family(c).marshal(), c["spec"], c["init"], c["T"], nchains=c["C"],
                     stream=A.marshal_stream(A.STREAM_FED, logu=c["logu"], z=c["z"]))
    for key in ("ans", "draws", "logpost"):
        assert np.array_equal(out[key], c["out"][key], equal_nan=True), key
    assert np.array_equal(out["istate"], c["out"]["istate"])
    if "mpsrf" in c["out"]:
        T = c["T"]
        psrf, mpsrf, rc = oracle.gelman(out["ans"][:, T // 2:, :][:, :, _free(c["spec"])])
        assert rc == int(c["out"]["gelman_rc"])
        np.testing.assert_array_equal(psrf, c["out"]["psrf"])
        assert mpsrf == float(c["out"]["mpsrf"]) or (np.isnan(mpsrf) and np.isnan(c["out"]["mpsrf"]))


def _paths(c):
    fam = c["family"]
    if fam == "hier_normal":
        return [0]
    p_x = np.atleast_2d(c["fam"]["X"]).shape[1] if c["fam"]["X"].ndim > 1 else 1
    return [0, 1, 3, 4] + ([2] if p_x <= 32 else [])     # 4: the split-integer tcgen05 kernel, forced (auto picks it above 128 chains)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_reproduces_golden(name):
    from fmcmc_b200.device import DeviceModel
    from gpu_util import assert_parity
    c = CASES[name]
    fam = _family(c)
    spec = c["spec"]
    kf = len(_free(spec))
    dlen = A.state_len(spec["type"], spec["k"], kf)
    for path in _paths(c):
        m = DeviceModel(fam)
        if path:
            m.set_path(path)
        ist = np.zeros((c["C"], A.ISTATE_LEN), dtype=np.int64)
        dst = np.zeros((c["C"], max(dlen, 1)))
        append = A.RUN_APPEND if "mpsrf" in c["out"] else 0
        if append:
            m.store_reset(c["C"], c["T"])
        g = m.run(spec, c["T"], c["C"], initial=c["init"], istate=ist, dstate=dst if dlen else None, flags=append,
                  stream=A.marshal_stream(A.STREAM_FED, logu=c["logu"], z=c["z"]))
        assert_parity(g, c["out"], RTOL, f"{name} path {path}")
        assert np.array_equal(ist, c["out"]["istate"]), "integer kernel state"
        if dlen:
            ref = c["out"]["dstate"]
            np.testing.assert_allclose(dst, ref, rtol=1e-10, atol=1e-12 * max(np.abs(ref).max(), 1e-300))
        if append and int(c["out"]["gelman_rc"]) == 0:
            T = c["T"]
            free = np.zeros(spec["k"], dtype=np.uint8)
            free[_free(spec)] = 1
            xb, s2, ws = m.gelman_partials(T // 2, T, free, c["C"])
            psrf, mpsrf = m.gelman_finish(T - T // 2, c["C"], kf, xb, s2, ws)
            np.testing.assert_allclose(psrf, c["out"]["psrf"], rtol=1e-9)
            np.testing.assert_allclose(mpsrf, float(c["out"]["mpsrf"]), rtol=1e-9)
        m.close()
