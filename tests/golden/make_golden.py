#!/usr/bin/env python
"""Generates the committed golden fixtures of tests/golden/ (run in the BUILD container only).

  lifeexpect.npz   the reference's only data fixture, data/lifeexpect.rda (R/data.R:1-37,
                   data-raw/lifeexpect.R:3-28), converted from its gzip'd XDR serialisation.  Reads
                   /root/reference, which does not exist on the GPU box — hence the committed copy.
  cases.npz        fed-stream input/output vectors for every kernel x family on the hot path, produced by
                   the CPU oracle AFTER it was pinned on the reference's published outputs
                   (tests/test_oracle_readme_golden.py: README.md:183-201, 315-339, 388-412).  R is not
                   installed, so the reference itself cannot generate them; the streams of the README-model
                   cases come from the restated R RNG (oracle/r_rng.c) in the serial path's order.

Tests: tests/test_golden.py checks (CPU) that the oracle still reproduces cases.npz bit for bit and
(GPU) that the CUDA path reproduces it through the C ABI (decisions identical, samples <= 1e-12).

usage: python tests/golden/make_golden.py [--reference /root/reference]
"""
import argparse
import gzip
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


# ----------------------------------------------------------------------------------- lifeexpect.rda
def read_rda_dataframe(path):
    """Minimal reader of R's XDR serialisation (format 2/3): enough for a data.frame of INTSXP/REALSXP."""
    raw = gzip.open(path, "rb").read()
    assert raw[:5] == b"RDX2\n" or raw[:5] == b"RDX3\n", raw[:8]
    pos = 5
    assert raw[pos:pos + 2] == b"X\n"
    pos += 2

    def i32():
        nonlocal pos
        v = struct.unpack(">i", raw[pos:pos + 4])[0]
        pos += 4
        return v

    version = i32(); i32(); i32()
    assert version == 2, f"serialisation version {version} not handled (lifeexpect.rda is version 2)"

    def item():
        nonlocal pos
        flags = i32()
        typ = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if typ == 254:                         # NILVALUE_SXP
            return None
        if typ == 2:                           # LISTSXP (pairlist)
            attr = item() if has_attr else None
            tag = item() if has_tag else None
            car = item()
            cdr = item()
            return [("pair", tag, car)] + (cdr if isinstance(cdr, list) else [])
        if typ == 1:                           # SYMSXP
            s = item()
            syms.append(s)
            return s
        if typ == 255:                         # REFSXP
            return syms[(flags >> 8) - 1]
        if typ == 9:                           # CHARSXP
            n = i32()
            if n == -1:
                return None
            s = raw[pos:pos + n].decode()
            pos += n
            return s
        if typ == 13:                          # INTSXP
            n = i32()
            v = np.frombuffer(raw, dtype=">i4", count=n, offset=pos).astype(np.int32)
            pos += 4 * n
        elif typ == 14:                        # REALSXP
            n = i32()
            v = np.frombuffer(raw, dtype=">f8", count=n, offset=pos).astype(np.float64)
            pos += 8 * n
        elif typ == 16:                        # STRSXP
            n = i32()
            v = [item() for _ in range(n)]
        elif typ == 19:                        # VECSXP
            n = i32()
            v = [item() for _ in range(n)]
        else:
            raise ValueError(f"unsupported SEXP type {typ} at {pos}")
        attr = item() if has_attr else None
        return {"value": v, "attr": attr} if attr else v

    syms = []
    top = item()                               # pairlist: name -> data.frame
    (_, name, df), = [t for t in top if t[0] == "pair"]
    names = [a[2] for a in df["attr"] if a[1] == "names"][0]
    cols = {}
    for nm, col in zip(names, df["value"]):
        cols[nm] = col["value"] if isinstance(col, dict) else col
    return name, cols


def make_lifeexpect(ref):
    name, cols = read_rda_dataframe(os.path.join(ref, "data", "lifeexpect.rda"))
    assert name == "lifeexpect" and set(cols) == {"smoke", "female", "age"}, (name, list(cols))
    smoke, female, age = cols["smoke"], cols["female"], cols["age"]
    assert smoke.size == female.size == age.size == 1000
    np.savez_compressed(os.path.join(HERE, "lifeexpect.npz"), smoke=smoke, female=female, age=age)
    print(f"lifeexpect.npz: n={age.size} mean(age)={age.mean():.3f} sd={age.std(ddof=1):.3f} "
          f"mean(smoke)={smoke.mean():.3f} mean(female)={female.mean():.3f}")


# ----------------------------------------------------------------------------------------- cases.npz
def golden_cases():
    """name -> dict(family=(kind, kwargs), spec, init, T, C, stream kind, seed).  Everything a test needs
    to rebuild the inputs is stored in the npz, so neither numpy's generators nor the R RNG layer have to
    reproduce anything at test time."""
    from fmcmc_b200 import _abi as A
    from oracle import oracle as O
    R = O.RRng
    cases = {}

    # README model, README.md:112-115 (R RNG, seed 78845)
    R.set_seed(78845)
    n = 1000
    Xr = R.rnorm(n)
    yr = 3.0 + 2.0 * Xr + R.rnorm(n, 0.0, 4.0)
    sd_y = R.sd(yr)
    readme = ("gaussian_lm", dict(X=Xr.reshape(-1, 1), y=yr, intercept=True, guard=True))
    lb3 = np.array([-A.DBL_MAX, -A.DBL_MAX, 0.0])

    def r_stream(C, T, kd, kind="normal"):
        from helpers import r_fed_stream
        return r_fed_stream(R, C, T, kd, kind)

    # config 1: kernel_normal(scale = .1), 1 chain  (README.md:161-166 with the BASELINE scale)
    R.set_seed(1215)
    logu, z = r_stream(1, 400, 3)
    cases["cfg1_readme_normal"] = dict(family=readme, spec=dict(type=A.KERNEL_NORMAL, k=3, mu=0.0, scale=0.1),
                                       init=np.array([[0, 0, sd_y]]), T=400, C=1, logu=logu, z=z)
    # config 2: 4 chains, kernel_normal_reflective(lb on sd), Gelman on the second half
    R.set_seed(1215)
    logu, z = r_stream(4, 300, 3)
    init = np.tile([0, 0, sd_y], (4, 1)) + np.arange(4)[:, None] * np.array([0.5, -0.3, 0.1])
    cases["cfg2_readme_reflective_gelman"] = dict(
        family=readme, spec=dict(type=A.KERNEL_NORMAL_REFLECTIVE, k=3, mu=0.0, scale=0.1, lb=lb3, ub=A.DBL_MAX),
        init=init, T=300, C=4, logu=logu, z=z, gelman=True)
    R.set_seed(7)
    logu, z = r_stream(2, 250, 3, "unif")
    cases["readme_unif_reflective"] = dict(
        family=readme, spec=dict(type=A.KERNEL_UNIF_REFLECTIVE, k=3, min_=-0.4, max_=0.4, lb=lb3, ub=8.0),
        init=np.tile([1.0, 1.0, sd_y], (2, 1)), T=250, C=2, logu=logu, z=z)
    R.set_seed(8)
    logu, z = r_stream(2, 250, 1)
    cases["readme_normal_ordered_fixed"] = dict(
        family=readme, spec=dict(type=A.KERNEL_NORMAL, k=3, mu=0.0, scale=0.3, scheme=A.SCHEME_ORDERED,
                                 fixed=[False, True, False]),
        init=np.tile([1.0, 2.0, sd_y], (2, 1)), T=250, C=2, logu=logu, z=z)
    R.set_seed(9)
    logu, z = r_stream(3, 300, 3, "unif")
    cases["readme_umirror"] = dict(
        family=readme, spec=dict(type=A.KERNEL_UMIRROR, k=3, mu=np.array([3.0, 2.0, 4.0]), scale=0.4, warmup=120,
                                 arate=0.4, lb=lb3, ub=A.DBL_MAX, nadapt=np.array([30, 60, 90, 120])),
        init=np.tile([2.0, 1.0, sd_y], (3, 1)), T=300, C=3, logu=logu, z=z)

    # config 3 (small): logistic + kernel_adapt, numpy streams
    rng = np.random.default_rng(20260317)
    n, p = 700, 8
    X = rng.standard_normal((n, p)) / np.sqrt(p)
    X[:, 0] = 1.0
    y = (rng.random(n) < 1 / (1 + np.exp(-X @ rng.standard_normal(p)))).astype(np.float64)
    logistic = ("logistic", dict(X=X, y=y, prior_sd=2.0))
    C, T = 3, 260
    cases["cfg3_logistic_adapt"] = dict(
        family=logistic, spec=dict(type=A.KERNEL_ADAPT, k=p, mu=0.0, warmup=60, freq=1, eps=1e-4),
        init=rng.normal(0, 0.1, (C, p)), T=T, C=C, logu=np.log(rng.random((C, T))),
        z=rng.standard_normal((C, T, p)))
    cases["logistic_adapt_freq4"] = dict(
        family=logistic, spec=dict(type=A.KERNEL_ADAPT, k=p, mu=0.0, warmup=40, freq=4, eps=1e-4),
        init=rng.normal(0, 0.1, (C, p)), T=T, C=C, logu=np.log(rng.random((C, T))),
        z=rng.standard_normal((C, T, p)))

    # config 4: lifeexpect, (a) the documented regression with kernel_ram (vignettes/advanced-features.Rmd:113-124),
    #           (b) hierarchical normal on smoke x female cells (SURVEY §8d reading)
    le = np.load(os.path.join(HERE, "lifeexpect.npz"))
    age, smoke, female = le["age"], le["smoke"].astype(np.float64), le["female"].astype(np.float64)
    sd_age = float(np.std(age, ddof=1))
    C, T = 3, 300
    lifereg = ("gaussian_lm", dict(X=np.c_[smoke, female], y=age, intercept=True, guard=True))
    cases["cfg4_lifeexpect_regression_ram"] = dict(
        family=lifereg, spec=dict(type=A.KERNEL_RAM, k=4, warmup=0, freq=1, eps=1e-4, arate=0.234,
                                  lb=np.array([-A.DBL_MAX] * 3 + [0.001]), ub=A.DBL_MAX),
        init=np.tile([70, 0, 0, sd_age], (C, 1)), T=T, C=C, logu=np.log(rng.random((C, T))),
        z=rng.standard_t(4, size=(C, T, 4)))
    grp = (2 * le["smoke"] + le["female"]).astype(np.int32)
    hier = ("hier_normal", dict(y=age, group=grp, n_groups=4, gamma_bounds=(0.0, 150.0), estimate_scales=True))
    cases["cfg4_lifeexpect_hier_ram"] = dict(
        family=hier, spec=dict(type=A.KERNEL_RAM, k=7, warmup=0, freq=1, eps=1e-2, arate=0.234,
                               lb=np.array([-A.DBL_MAX] * 5 + [1e-3, 1e-3]), ub=A.DBL_MAX),
        init=np.tile([75, 75, 75, 75, 75, 5, 5], (C, 1)) + rng.normal(0, 0.1, (C, 7)), T=T, C=C,
        logu=np.log(rng.random((C, T))), z=rng.standard_t(7, size=(C, T, 7)))
    # the playground model itself (hierarchical-bayes.Rmd:28-51): unit variances, gamma ~ U(-1, 1)
    N, Nc = 1000, 20
    g20 = (np.arange(N) % Nc).astype(np.int32)
    th = rng.normal(0.3, 1, Nc)
    yh = rng.normal(th[g20], 1.0)
    k = Nc + 1
    cases["playground_hier_normal_reflective"] = dict(
        family=("hier_normal", dict(y=yh, group=g20, n_groups=Nc, gamma_bounds=(-1.0, 1.0), estimate_scales=False)),
        spec=dict(type=A.KERNEL_NORMAL_REFLECTIVE, k=k, mu=0.0, scale=0.05,
                  lb=np.r_[np.full(Nc, -A.DBL_MAX), -1.0], ub=np.r_[np.full(Nc, A.DBL_MAX), 1.0]),
        init=np.zeros((2, k)), T=200, C=2, logu=np.log(rng.random((2, 200))), z=rng.standard_normal((2, 200, k)))

    # config 5 (small): Gaussian LM, p_x = 40 (> 32: the DMMA kernel's wide tier when the tiled path is forced),
    # kernel_nmirror with lb on sd
    n, p = 600, 40
    X = rng.standard_normal((n, p))
    beta = rng.standard_normal(p)
    y = 1.0 + X @ beta + rng.normal(0, 2.0, n)
    k = p + 2
    lb = np.full(k, -A.DBL_MAX); lb[-1] = 0.0
    C, T = 4, 200
    centre = np.r_[1.0, beta, 2.0]
    cases["cfg5_gaussian_nmirror"] = dict(
        family=("gaussian_lm", dict(X=X, y=y, intercept=True, guard=True)),
        spec=dict(type=A.KERNEL_NMIRROR, k=k, mu=centre, scale=0.02, warmup=100, arate=0.4, lb=lb, ub=A.DBL_MAX,
                  nadapt=np.array([25, 50, 75, 100])),
        init=centre + rng.normal(0, 0.05, (C, k)), T=T, C=C,
        logu=np.log(rng.random((C, T))), z=rng.standard_normal((C, T, k)), gelman=True)
    return cases


def build_family(fam):
    import fmcmc_b200 as fm
    kind, kw = fam
    return {"gaussian_lm": fm.ll_gaussian_lm, "logistic": fm.ll_logistic, "hier_normal": fm.ll_hier_normal}[kind](**kw)


def run_oracle_case(case):
    from fmcmc_b200 import _abi as A
    from oracle import oracle as O
    fam = build_family(case["family"])
    out = O.run(fam.marshal(), case["spec"], case["init"], case["T"], nchains=case["C"],
                stream=A.marshal_stream(A.STREAM_FED, logu=case["logu"], z=case["z"]))
    res = dict(ans=out["ans"], draws=out["draws"], logpost=out["logpost"], istate=out["istate"], dstate=out["dstate"])
    if case.get("gelman"):
        T = case["T"]
        psrf, mpsrf, rc = O.gelman(out["ans"][:, T // 2:, :][:, :, free_cols(case["spec"])])
        res["psrf"], res["mpsrf"], res["gelman_rc"] = psrf, np.array(mpsrf), np.array(rc)   # rc 5: chol(W) failed
    return res


def free_cols(spec):
    k = spec["k"]
    fixed = np.broadcast_to(np.asarray(spec.get("fixed", False), dtype=bool), (k,))
    return np.where(~fixed)[0]


def flatten(cases, results):
    flat = {"__names__": np.array(sorted(cases))}
    for name, c in cases.items():
        kind, kw = c["family"]
        flat[f"{name}/family"] = np.array(kind)
        for a, v in kw.items():
            flat[f"{name}/fam/{a}"] = np.asarray(v)
        for a, v in c["spec"].items():
            flat[f"{name}/spec/{a}"] = np.asarray(v)
        for a in ("init", "T", "C", "logu", "z"):
            flat[f"{name}/{a}"] = np.asarray(c[a])
        for a, v in results[name].items():
            flat[f"{name}/out/{a}"] = np.asarray(v)
    return flat


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    make_lifeexpect(args.reference)
    cases = golden_cases()
    results = {name: run_oracle_case(c) for name, c in cases.items()}
    flat = flatten(cases, results)
    np.savez_compressed(os.path.join(HERE, "cases.npz"), **flat)
    sz = os.path.getsize(os.path.join(HERE, "cases.npz"))
    for name in sorted(cases):
        a = results[name]["ans"]
        moved = np.any(a[:, 1:] != a[:, :-1], axis=2).mean()
        print(f"{name:40s} ans {a.shape} accept {moved:.3f}" +
              (f" mpsrf {float(results[name]['mpsrf']):.4f}" if "mpsrf" in results[name] else ""))
    print(f"cases.npz: {len(cases)} cases, {sz / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
