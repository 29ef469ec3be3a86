"""N > 1 path on CPU: world_size-2 gloo run of the Gelman exchange (all_gather of per-chain means /
variances + all_reduce of the within-chain scatter).  The CUDA statistics kernels are replaced by a numpy
stand-in with the same interface so only the sharding / collective logic is exercised here."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyStatsModel:
    """Same methods as fmcmc_b200.device.DeviceModel's Gelman part, on host arrays."""

    def __init__(self, samples):          # [rows][C_local][k]
        self.s = samples

    def gelman_partials(self, row_begin, row_end, free_mask, nlocal, out=None):
        x = self.s[row_begin:row_end][:, :, np.asarray(free_mask, dtype=bool)]
        xbar = x.mean(axis=0)
        s2 = x.var(axis=0, ddof=1)
        ws = sum(np.cov(x[:, c, :].T) for c in range(x.shape[1]))
        return xbar, s2, np.asfortranarray(np.atleast_2d(ws))

    def gelman_finish(self, niter, nchains_total, kf, xbar, s2, wsum, dev_in=False):
        W = np.asarray(wsum).reshape(kf, kf) / nchains_total
        B = niter * np.atleast_2d(np.cov(np.asarray(xbar).T))
        L = np.linalg.cholesky(W)
        M = np.linalg.solve(L, np.linalg.solve(L, B).T)
        emax = np.linalg.eigvalsh((M + M.T) / 2).max()
        return np.sqrt(np.diag(B) / np.diag(W)), float(np.sqrt((1 - 1 / niter) + (1 + 1 / kf) * emax / niter))


def _worker(rank, world, port, counts, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fmcmc_b200.dist import ChainSharding
    total, k, rows = sum(counts), 3, 60
    rng = np.random.default_rng(5)
    allx = rng.standard_normal((rows, total, k)).cumsum(axis=0) * 0.1 + rng.standard_normal((1, total, k))
    sh = ChainSharding(total)
    assert sh.counts == counts and sh.local == counts[rank] and sh.offset == sum(counts[:rank])
    local = allx[:, sh.offset:sh.offset + sh.local, :]
    free = np.array([1, 0, 1], dtype=np.uint8)
    psrf, mpsrf = sh.gelman(NumpyStatsModel(local), rows // 2, rows, free, sh.local, 2, rows - rows // 2)
    ref_psrf, ref_mpsrf = None, None
    if rank == 0:
        m = NumpyStatsModel(allx)
        xb, s2, ws = m.gelman_partials(rows // 2, rows, free, total)
        ref_psrf, ref_mpsrf = m.gelman_finish(rows - rows // 2, total, 2, xb, s2, ws)
    q.put((rank, np.asarray(psrf), mpsrf, ref_psrf, ref_mpsrf))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("counts", [[4, 4], [3, 2]])
def test_gelman_exchange_world2_gloo(counts):
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
    (_, p0, m0, rp, rm), (_, p1, m1, _, _) = res
    assert np.array_equal(p0, p1) and m0 == m1            # every rank takes the same decision
    np.testing.assert_allclose(p0, rp, rtol=1e-12)
    np.testing.assert_allclose(m0, rm, rtol=1e-12)
