"""Chain sharding on >= 2 GPUs (R/mcmc.R:593-627 -> one rank per GPU): the R-hat exchange over NCCL gives the numbers one GPU
holding every chain gives, and MCMC(conv_checker = convergence_gelman()) under torchrun stops at the same bulk with the same
R-hat trace and the same samples.  Runs whenever the box shows >= 2 devices (`gpurun --gpus 2`), with as many ranks as
there are devices (up to 8); skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nccl_gelman_and_autostop_match_one_gpu():
    import fmcmc_b200 as fm
    ngpu = fm.lib().fmcmc_device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(ngpu, 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "multi_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-6000:]
    assert "GELMAN_NCCL_OK" in r.stdout and "MCMC_NCCL_OK" in r.stdout, r.stdout[-2000:]
    print(r.stdout[-600:])
