"""Observation sharding across GPUs (include/fmcmc_b200.h, fmcmc_shard_*): every rank holds a row slice of X / y
and runs the same chains; partial log-likelihood sums are exchanged inside the kernels over NVLink peer memory.
Needs >= 2 GPUs (run with `gpurun --gpus 2`); skipped otherwise."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from fmcmc_b200 import _abi as A

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import fmcmc_b200 as fm
    return fm.lib().fmcmc_device_count()


def _data(family, n, p, rng):
    import fmcmc_b200 as fm
    if family == "logistic":
        X = rng.standard_normal((n, p)) / np.sqrt(p)
        X[:, 0] = 1.0
        y = (rng.random(n) < 1 / (1 + np.exp(-X @ rng.standard_normal(p)))).astype(np.float64)
        return X, y, (lambda Xs, ys: fm.ll_logistic(Xs, ys, prior_sd=2.0)), p
    X = rng.standard_normal((n, p))
    y = 1.0 + X @ rng.standard_normal(p) + rng.normal(0, 2.0, n)
    return X, y, (lambda Xs, ys: fm.ll_gaussian_lm(Xs, ys, intercept=True, guard=True)), p + 2


CASES = {
    "logistic_normal": ("logistic", 32, 4, lambda k: dict(type=A.KERNEL_NORMAL, k=k, mu=0.0, scale=0.02)),
    "logistic_adapt": ("logistic", 20, 3, lambda k: dict(type=A.KERNEL_ADAPT, k=k, mu=0.0, warmup=10, freq=1, eps=1e-4)),
    "gaussian_ram": ("gaussian", 6, 2, lambda k: dict(type=A.KERNEL_RAM, k=k, warmup=0, freq=1, eps=1e-3, arate=0.234,
                                                      lb=np.r_[np.full(k - 1, -A.DBL_MAX), 0.0], ub=A.DBL_MAX)),
    "gaussian_wide_nmirror": ("gaussian", 100, 2, lambda k: dict(type=A.KERNEL_NORMAL_REFLECTIVE, k=k, mu=0.0, scale=0.01,
                                                                 lb=np.r_[np.full(k - 1, -A.DBL_MAX), 0.0], ub=A.DBL_MAX)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_two_gpus_one_process(name):
    """Two models on cuda:0 / cuda:1 in one process (peer access), driven from two host threads: both ranks return
    bit-identical results, equal to the un-sharded run on the full data (same decisions, 1e-12)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    from fmcmc_b200.device import DeviceModel
    from gpu_util import assert_parity
    family, p, C, mk = CASES[name]
    rng = np.random.default_rng(77)
    n = 2 * 20_000 + 6
    X, y, make, k = _data(family, n, p, rng)
    spec = mk(k)
    T = 40
    init = rng.normal(0, 0.05, (C, k))
    if family == "gaussian":
        init[:, -1] = 3.0
    stream = lambda: A.marshal_stream(A.STREAM_PHILOX, seed=99, run_index=0)
    full = DeviceModel(make(X, y), device=0)
    full.set_path(3)
    ref = full.run(spec, T, C, initial=init, stream=stream())
    full.close()

    h = (n // 2) & ~1
    parts = [(X[:h], y[:h]), (X[h:], y[h:])]
    models = [DeviceModel(make(np.asfortranarray(Xs), ys), device=d) for d, (Xs, ys) in enumerate(parts)]
    handles = [m.shard_alloc(2, 2 * C, n) for m in models]
    for r, m in enumerate(models):
        m.shard_attach(r, 2, handles)
    outs, errs = [None, None], [None, None]

    def work(r):
        try:
            outs[r] = models[r].run(spec, T, C, initial=init, stream=stream())
            outs[r + 0] = outs[r]
        except Exception as e:  # noqa: BLE001
            errs[r] = e

    th = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join(timeout=120) for t in th]
    assert not any(t.is_alive() for t in th), "sharded run hung"
    assert errs == [None, None], errs
    for key in ("ans", "draws", "logpost"):
        assert np.array_equal(outs[0][key], outs[1][key], equal_nan=True), f"ranks disagree on {key}"
    assert_parity(outs[0], ref, 1e-12, name)
    # a second bulk continues from the device state on both ranks (flags stay monotonic)
    def work2(r):
        try:
            outs[r] = models[r].run(spec, 15, C, initial=None, stream=A.marshal_stream(A.STREAM_PHILOX, seed=99, run_index=1))
        except Exception as e:  # noqa: BLE001
            errs[r] = e
    th = [threading.Thread(target=work2, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join(timeout=120) for t in th]
    assert errs == [None, None], errs
    assert np.array_equal(outs[0]["ans"], outs[1]["ans"]) and np.array_equal(outs[0]["logpost"], outs[1]["logpost"])
    for m in models:
        m.close()


def test_two_processes_ipc_through_mcmc():
    """torchrun, one process per GPU: MCMC(..., shard="observations") exchanges CUDA IPC handles through
    torch.distributed once; rank 0 checks the result against the un-sharded run."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "shard_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "SHARD_OK" in r.stdout
