"""N > 1 path on CPU: world_size-2 gloo runs of everything convergence_gelman does across ranks -
all_gather of per-chain means / variances, all_reduce of the within-chain scatter, the pooled variance of rm_invariant, the
replicated finish and the checker's decision - against the ORACLE's gelman.diag on the un-sharded samples.

The CUDA statistics kernels are replaced by a numpy stand-in with the same interface (this suite has no GPU); its finish is
the whole of coda's formula (psrf with the df adjustment, mpsrf through chol(W)), not a shortcut, so what
is compared with the oracle is the complete result of the sharded path.  The same exchange with the real kernels and
NCCL runs in tests/test_gpu_multi.py."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyStatsModel:
    """Same methods as fmcmc_b200.device.DeviceModel's store / Gelman part, on host arrays."""

    def __init__(self, samples):          # [rows][C_local][k]
        self.s = samples

    def store_rows(self):
        return self.s.shape[0]

    def store_pooled(self, free_mask):
        x = self.s[:, :, np.asarray(free_mask, dtype=bool)].ravel()
        return np.array([x.size, x.mean(), ((x - x.mean()) ** 2).sum()])

    def gelman_partials(self, row_begin, row_end, free_mask, nlocal, out=None):
        x = self.s[row_begin:row_end][:, :, np.asarray(free_mask, dtype=bool)]
        xbar = x.mean(axis=0)
        s2 = x.var(axis=0, ddof=1)
        ws = sum(np.atleast_2d(np.cov(x[:, c, :].T)) for c in range(x.shape[1]))
        return xbar, s2, np.asfortranarray(np.atleast_2d(ws))

    def gelman_finish(self, niter, nchains_total, kf, xbar, s2, wsum, dev_in=False):
        """coda::gelman.diag from the gathered statistics, all of it."""
        N, m = float(niter), float(nchains_total)
        xbar, s2 = np.asarray(xbar).reshape(-1, kf), np.asarray(s2).reshape(-1, kf)
        W = np.asarray(wsum).reshape(kf, kf) / m
        B = N * np.atleast_2d(np.cov(xbar.T))
        mpsrf = float("nan")
        if kf > 1:
            L = np.linalg.cholesky(W)
            M = np.linalg.solve(L, np.linalg.solve(L, B).T)
            emax = np.linalg.eigvalsh((M + M.T) / 2).max()
            mpsrf = float(np.sqrt((1 - 1 / N) + (1 + 1 / kf) * emax / N))
        w, b = np.diag(W), np.diag(B)
        muhat = xbar.mean(axis=0)
        var_w = s2.var(axis=0, ddof=1) / m
        var_b = 2 * b * b / (m - 1)
        cov = lambda u, v: ((u - u.mean(axis=0)) * (v - v.mean(axis=0))).sum(axis=0) / (m - 1)   # noqa: E731
        cov_wb = (N / m) * (cov(s2, xbar ** 2) - 2 * muhat * cov(s2, xbar))
        V = (N - 1) / N * w + (1 + 1 / m) * b / N
        var_V = ((N - 1) ** 2 * var_w + (1 + 1 / m) ** 2 * var_b + 2 * (N - 1) * (1 + 1 / m) * cov_wb) / N ** 2
        df_adj = (2 * V * V / var_V + 3) / (2 * V * V / var_V + 1)
        return np.sqrt(df_adj * ((N - 1) / N + (1 + 1 / m) / N * b / w)), mpsrf


def _worker(rank, world, port, counts, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fmcmc_b200 as fm
    from fmcmc_b200.dist import ChainSharding, combine_pooled
    total, k, rows = sum(counts), 3, 60
    rng = np.random.default_rng(5)
    allx = rng.standard_normal((rows, total, k)).cumsum(axis=0) * 0.1 + rng.standard_normal((1, total, k))
    sh = ChainSharding(total)
    assert sh.counts == counts and sh.local == counts[rank] and sh.offset == sum(counts[:rank])
    local = allx[:, sh.offset:sh.offset + sh.local, :]
    free = np.array([1, 0, 1], dtype=np.uint8)
    tm = {}
    psrf, mpsrf = sh.gelman(NumpyStatsModel(local), rows // 2, rows, free, sh.local, 2, rows - rows // 2, timings=tm)
    assert set(tm) >= {"stats_ms", "all_gather_ms", "all_reduce_ms", "finish_ms"}
    pooled = sh.pooled_variance(NumpyStatsModel(local), free)
    # the checker itself, driven the way MCMC() drives it (R/mcmc.R:968): decision + message on every rank
    chk = fm.convergence_gelman(freq=rows, threshold=1.5)
    x = fm.McmcList([fm.Mcmc(local[:, c, :][:, [0, 2]], start=1, end=rows, thin=1) for c in range(sh.local)])
    chk._device_ctx = (NumpyStatsModel(local), sh.local, free, sh)
    decision = chk(x)
    msg = fm.convergence_msg_get()
    # degenerate store: every element equal -> rm_invariant drops column 1 (D9); one free column left -> psrf decides
    const = np.full((rows, sh.local, k), 2.5) + 1e-9 * rng.standard_normal((rows, sh.local, k))
    chk2 = fm.convergence_gelman(freq=rows)
    chk2._device_ctx = (NumpyStatsModel(const), sh.local, free, sh)
    xc = fm.McmcList([fm.Mcmc(const[:, c, :][:, [0, 2]], start=1, end=rows, thin=1) for c in range(sh.local)])
    chk2(xc)
    q.put((rank, np.asarray(psrf), mpsrf, pooled, bool(decision), msg, chk2._last_kf,
           combine_pooled([[3, 1.0, 2.0]]) if rank == 0 else None))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("counts", [[4, 4], [3, 2]])
def test_gelman_exchange_world2_gloo(counts, oracle):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, p0, m0, v0, d0, msg0, kf0, one), (_, p1, m1, v1, d1, msg1, kf1, _) = res
    assert np.array_equal(p0, p1) and m0 == m1 and v0 == v1            # every rank sees the same numbers ...
    assert d0 == d1 and msg0 == msg1                                   # ... and takes the same decision
    assert kf0 == kf1 == 1                                             # rm_invariant removed the first free column
    assert one == 1.0                                                  # M2 / (n - 1)
    # the un-sharded truth: the oracle's restatement of coda::gelman.diag on all chains (free columns, second half)
    total, k, rows = sum(counts), 3, 60
    rng = np.random.default_rng(5)
    allx = rng.standard_normal((rows, total, k)).cumsum(axis=0) * 0.1 + rng.standard_normal((1, total, k))
    win = np.ascontiguousarray(allx[rows // 2:, :, :][:, :, [0, 2]].transpose(1, 0, 2))
    rp, rm, rc = oracle.gelman(win)
    assert rc == 0
    np.testing.assert_allclose(p0, rp, rtol=1e-10)
    np.testing.assert_allclose(m0, rm, rtol=1e-10)
    np.testing.assert_allclose(v0, np.var(allx[:, :, [0, 2]].ravel(), ddof=1), rtol=1e-12)
    assert msg0 == "Gelman-Rubin's R: %.4f." % rm and d0 == (rm < 1.5)
