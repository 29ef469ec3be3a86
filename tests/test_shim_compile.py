"""The reference-side binding (integration/src/shim.c, the .Call layer of INTEGRATION.md) goes through a compiler and is
driven from C.  R is not installed here, so it is compiled against integration/rstub/ (a stand-in for the subset of R's
C API the shim uses) with -Wall -Wextra -Werror, linked against libfmcmcb200.so, and called by integration/test/drive_shim.c,
which builds the same lists integration/R/device.R builds.

CPU: compiles, links, registration table complete, errors come back through Rf_error with a balanced PROTECT stack.
GPU: the shim's outputs (R-layout arrays, kernel state carried through R objects between bulks, the device Gelman check)
are bit-identical to the same runs made through the Python mirror."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INTEG = os.path.join(ROOT, "integration")
DRIVER = os.path.join(INTEG, "_build", "drive_shim")


@pytest.fixture(scope="module")
def driver():
    import fmcmc_b200
    fmcmc_b200.build()                                         # the shim links against the product library
    r = subprocess.run(["make", "-C", INTEG], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return DRIVER


def test_shim_compiles_and_propagates_errors(driver):
    r = subprocess.run([driver, "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "7 routines registered" in r.stdout and "PROTECT stack balanced" in r.stdout


def test_shim_calls_every_entry_point_with_the_headers_signature():
    """Every fmcmc_* call in the shim names a function include/fmcmc_b200.h declares (the compiler checked the arguments);
    every routine R/device.R .Call()s is registered by the shim."""
    hdr = open(os.path.join(ROOT, "include", "fmcmc_b200.h")).read()
    shim = open(os.path.join(INTEG, "src", "shim.c")).read()
    glue = open(os.path.join(INTEG, "R", "device.R")).read()
    declared = set(re.findall(r"\b(fmcmc_[a-z0-9_]+)\s*\(", hdr))
    used = set(re.findall(r"\b(fmcmc_[a-z0-9_]+)\s*\(", shim))
    assert used <= declared, used - declared
    assert {"fmcmc_model_create", "fmcmc_model_free", "fmcmc_run", "fmcmc_gelman", "fmcmc_store_reset"} <= used
    registered = set(re.findall(r'\{"(C_fmcmc_[a-z_]+)"', shim))
    called = set(re.findall(r"\.Call\((C_fmcmc_[a-z_]+)", glue))
    assert called <= registered, called - registered


@pytest.mark.gpu
def test_shim_matches_the_python_mirror(driver, readme_data, tmp_path):
    import fmcmc_b200 as fm
    from fmcmc_b200 import _abi as A
    from fmcmc_b200.device import DeviceModel
    n, C, bulk, k = readme_data["n"], 6, 150, 3
    rng = np.random.default_rng(3)
    init = np.tile([0.5, 0.5, readme_data["sd_y"]], (C, 1)) + np.abs(rng.normal(0, 0.3, (C, k)))
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("4i", n, C, bulk, k))
        f.write(np.ascontiguousarray(readme_data["X"], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(readme_data["y"], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(init).tobytes())            # [C][k] row-major == t(initial) in R
    r = subprocess.run([driver, str(inp), str(outp)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "burnin" in r.stdout                                  # the reference's message came back through Rf_error
    got = np.fromfile(outp)
    pos = 0

    def take(*shape):
        nonlocal pos
        cnt = int(np.prod(shape))
        v = got[pos:pos + cnt].reshape(shape)
        pos += cnt
        return v

    fam = fm.ll_gaussian_lm(readme_data["X"], readme_data["y"], intercept=True, guard=True)
    m = DeviceModel(fam)
    try:
        # scenario A: kernel_normal_reflective, 3 bulks appended to the store, Gelman on the device
        spec = fm.kernel_normal_reflective(scale=0.05, lb=[-5.0, 0.0, 0.0], ub=5.0).to_spec(k)
        m.store_reset(C, 3 * bulk)
        for b in range(3):
            o = m.run(spec, bulk, C, initial=init if b == 0 else None, flags=A.RUN_APPEND,
                      stream=A.marshal_stream(A.STREAM_PHILOX, seed=11, run_index=b))
            ans_r = take(C, k, bulk)                              # R's [row, param, chain] array, seen from C order
            assert np.array_equal(ans_r.transpose(0, 2, 1), o["ans"]), f"bulk {b}"
            assert np.array_equal(take(C, bulk), o["logpost"])
        psrf, mpsrf, used = m.gelman(np.ones(k, dtype=np.uint8))
        assert np.array_equal(take(k), psrf) and take(1)[0] == mpsrf and take(1)[0] == used == 225
        # scenario B: kernel_adapt, state through R objects
        spec = fm.kernel_adapt(warmup=20, lb=[-5.0, 0.0, 0.0], ub=5.0).to_spec(k)
        ist = np.zeros((C, A.ISTATE_LEN), dtype=np.int64)
        dst = np.zeros((C, A.state_len(A.KERNEL_ADAPT, k, k)))
        for b in range(2):
            o = m.run(spec, bulk, C, initial=init if b == 0 else None, istate=ist, dstate=dst,
                      stream=A.marshal_stream(A.STREAM_PHILOX, seed=12, run_index=b))
            assert np.array_equal(take(C, k, bulk).transpose(0, 2, 1), o["ans"]), f"adapt bulk {b}"
        assert np.array_equal(take(C, A.ISTATE_LEN), ist.astype(np.float64))
        assert np.array_equal(take(C, dst.shape[1]), dst)
        assert np.allclose(take(2), [0.7, 0.4], rtol=0, atol=1e-15)   # reflect_on_boundaries
        assert pos == got.size
    finally:
        m.close()
