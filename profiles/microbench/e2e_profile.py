"""cProfile of the public call fm.MCMC(..., conv_checker=convergence_gelman(11)) at bench.py's default workload (cfg3, 100 MH steps
in 10 bulks): where the host spends the ~1.3 ms per bulk that the end-to-end figure loses against the device-timed one."""
import cProfile, io, os, pstats, sys, time, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import fmcmc_b200 as fm

X, y = bench.make_data()
fam = fm.ll_logistic(X, y, prior_sd=2.0)
rng = np.random.default_rng(1000)
C, k = 1024, 32
init = rng.normal(0, 0.1, (C, k))
kern = fm.kernel_adapt()
with open(os.devnull, "w") as dn, contextlib.redirect_stderr(dn):
    fm.MCMC(init, fam, 560, nchains=C, kernel=kern, seed=1)            # past the kernel's warm-up; model resident
    def call():
        return fm.MCMC(init, fam, 110, nchains=C, kernel=kern, conv_checker=fm.convergence_gelman(freq=11, threshold=0.0), seed=2)
    call()
    t0 = time.perf_counter(); call(); t1 = time.perf_counter()
    pr = cProfile.Profile(); pr.enable(); call(); pr.disable()
print("wall of one call: %.1f ms" % (1e3 * (t1 - t0)))
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:6000])
# the driver's --steps 20: two bulks, kernel state carried in from the host (as bench.py's e2e does)
from fmcmc_b200 import _abi as A
ist = np.zeros((C, A.ISTATE_LEN), dtype=np.int64); dst = np.zeros((C, A.state_len(A.KERNEL_ADAPT, k, k)))
with open(os.devnull, "w") as dn, contextlib.redirect_stderr(dn):
    k0 = fm.kernel_adapt(); fm.MCMC(init, fam, 560, nchains=C, kernel=k0, seed=1)
    i0, d0 = k0.state_arrays(C, k)
    def call20():
        kq = fm.kernel_adapt(); kq.load_state(i0, d0, C, k)
        t0 = time.perf_counter()
        a = fm.MCMC(init, fam, 22, nchains=C, kernel=kq, conv_checker=fm.convergence_gelman(freq=11, threshold=0.0), seed=3)
        return time.perf_counter() - t0
    call20(); print("22-row call: %.1f ms" % (1e3 * call20()))
    pr = cProfile.Profile(); pr.enable(); call20(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(16); print(s.getvalue()[:5000])
