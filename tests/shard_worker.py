"""Worker of tests/test_gpu_shard.py::test_two_processes_ipc_through_mcmc (launched by torchrun, 2 ranks)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import fmcmc_b200 as fm
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank()
    rng = np.random.default_rng(5)                     # same data on every rank
    n, p, C, T = 60_000, 16, 3, 60
    X = rng.standard_normal((n, p)) / np.sqrt(p)
    X[:, 0] = 1.0
    y = (rng.random(n) < 1 / (1 + np.exp(-X @ rng.standard_normal(p)))).astype(np.float64)
    fam = fm.ll_logistic(X, y)
    init = rng.normal(0, 0.05, (C, p))
    ans = fm.MCMC(init, fam, T, nchains=C, seed=3, kernel=fm.kernel_adapt(warmup=20), shard="observations", path=3)
    a = ans.as_array()
    gathered = [None, None]
    dist.all_gather_object(gathered, a.tobytes())
    assert gathered[0] == gathered[1], "ranks returned different chains"
    if rank == 0:
        dist.barrier()
    else:
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0:                                      # un-sharded reference on this GPU alone
        ref = fm.MCMC(init, fam, T, nchains=C, seed=3, kernel=fm.kernel_adapt(warmup=20), device=local, path=3).as_array()
        moved_a = np.any(a[:, 1:] != a[:, :-1], axis=2)
        moved_r = np.any(ref[:, 1:] != ref[:, :-1], axis=2)
        assert np.array_equal(moved_a, moved_r), "decisions differ from the un-sharded run"
        err = np.max(np.abs(a - ref).max(axis=(0, 1)) / np.abs(ref).max(axis=(0, 1)))
        assert err <= 1e-12, err
        print(f"SHARD_OK accept={moved_a.mean():.3f} max_rel_err={err:.2e}")


if __name__ == "__main__":
    main()
