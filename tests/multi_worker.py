"""Worker of tests/test_gpu_multi.py (launched by torchrun, one rank per GPU): chain sharding with the R-hat exchange over
NCCL.  Rank 0 re-runs everything un-sharded on its own GPU and compares."""
import io
import os
import re
import sys
from contextlib import redirect_stderr

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import fmcmc_b200 as fm
    from fmcmc_b200 import _abi as A
    from fmcmc_b200.device import DeviceModel
    from fmcmc_b200.dist import ChainSharding
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()

    # ---- 1. ChainSharding.gelman over NCCL == fmcmc_gelman on one GPU holding all chains ---------------------------
    rng = np.random.default_rng(21)                    # same data on every rank
    n, p = 3000, 12
    X = rng.standard_normal((n, p)) / np.sqrt(p)
    X[:, 0] = 1.0
    y = (rng.random(n) < 1 / (1 + np.exp(-X @ rng.standard_normal(p)))).astype(np.float64)
    fam = fm.ll_logistic(X, y)
    C_tot, T = 16 * world + 3, 240                     # uneven shards on purpose
    init = rng.normal(0, 0.3, (C_tot, p))
    spec = dict(type=A.KERNEL_NORMAL, k=p, mu=0.0, scale=0.08)
    sh = ChainSharding(C_tot)
    free = np.ones(p, dtype=np.uint8)
    free[3] = 0
    kf = int(free.sum())
    m = DeviceModel(fam, device=local)
    m.store_reset(sh.local, T)
    m.run(spec, T, sh.local, initial=init[sh.offset:sh.offset + sh.local], flags=A.RUN_APPEND, chain_offset=sh.offset,
          stream=A.marshal_stream(A.STREAM_PHILOX, seed=77), nchains_total=C_tot, outputs=False)
    first = T // 2
    tm = {}
    psrf, mpsrf = sh.gelman(m, first, T, free, sh.local, kf, T - first, timings=tm)
    pooled = sh.pooled_variance(m, free)
    m.close()
    if rank == 0:
        m1 = DeviceModel(fam, device=local)
        m1.store_reset(C_tot, T)
        m1.run(spec, T, C_tot, initial=init, flags=A.RUN_APPEND, stream=A.marshal_stream(A.STREAM_PHILOX, seed=77), outputs=False)
        p1, mp1, used = m1.gelman(free)
        from fmcmc_b200.dist import combine_pooled
        v1 = combine_pooled([m1.store_pooled(free)])
        m1.close()
        assert used == T - first
        e1 = float(np.max(np.abs(psrf - p1) / p1))
        e2 = abs(mpsrf - mp1) / mp1
        e3 = abs(pooled - v1) / v1
        assert e1 <= 1e-12 and e2 <= 1e-12 and e3 <= 1e-12, (e1, e2, e3)
        print(f"GELMAN_NCCL_OK world={world} chains={C_tot} psrf_err={e1:.1e} mpsrf_err={e2:.1e} pooled_err={e3:.1e} "
              f"stats={tm['stats_ms']:.2f}ms gather={tm['all_gather_ms']:.2f}ms reduce={tm['all_reduce_ms']:.2f}ms "
              f"finish={tm['finish_ms']:.2f}ms")

    # ---- 2. MCMC(conv_checker = convergence_gelman()) under torchrun stops where one GPU stops --------------------------
    rng = np.random.default_rng(22)
    nn = 1000
    Xr = rng.standard_normal(nn)
    yr = 3.0 + 2.0 * Xr + rng.normal(0, 4.0, nn)
    ll = fm.ll_gaussian_lm(Xr.reshape(-1, 1), yr, intercept=True, guard=True)
    C2 = 4 * world
    init2 = np.tile([0.0, 0.0, float(np.std(yr, ddof=1))], (C2, 1)) + np.abs(rng.normal(0, 3.0, (C2, 3)))   # dispersed starts

    def go(**kw):
        buf = io.StringIO()
        with redirect_stderr(buf):
            a = fm.MCMC(init2, ll, 6000, nchains=C2, seed=5, kernel=fm.kernel_normal_reflective(scale=0.04, lb=[-19.0, -19.0, 0.0], ub=19.0),
                        conv_checker=fm.convergence_gelman(300, threshold=1.03), **kw)
        return a, [float(v) for v in re.findall(r"Gelman-Rubin's R: ([0-9.]+)\.", buf.getvalue())], buf.getvalue()

    a, trace, text = go()
    mine = a.as_array()                                # this rank's chains
    assert mine.shape[0] == C2 // world
    rows = [None] * world
    dist.all_gather_object(rows, (mine.shape[1], trace))
    assert all(r == rows[0] for r in rows), rows       # every rank stopped at the same bulk with the same trace
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        b, trace1, text1 = go(device=local)            # torch.distributed is gone: all chains on this GPU
        ref = b.as_array()
        assert ref.shape[1] == mine.shape[1], (ref.shape, mine.shape)
        assert len(trace) >= 2 and len(trace) == len(trace1), (trace, trace1)
        assert trace == trace1, (trace, trace1)         # 4-decimal R-hat values of every bulk
        assert np.array_equal(ref[:mine.shape[0]], mine), "rank 0's chains differ from the un-sharded run"
        assert ("Convergence has been reached" in text) == ("Convergence has been reached" in text1)
        print(f"MCMC_NCCL_OK world={world} stopped_at={mine.shape[1]} checks={len(trace)} last_R={trace[-1]}")


if __name__ == "__main__":
    main()
