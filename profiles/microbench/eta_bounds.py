import os, sys, contextlib
sys.path.insert(0, os.getcwd())
import numpy as np, bench, fmcmc_b200 as fm
X, y = bench.make_data()
rmax = np.sqrt((X * X).sum(axis=1).max()); cmax = np.abs(X).max(axis=0)
fam = fm.ll_logistic(X, y, prior_sd=2.0)
rng = np.random.default_rng(1000)
C, k = 1024, 32
init = rng.normal(0, 0.1, (C, k))
with open(os.devnull, "w") as dn, contextlib.redirect_stderr(dn):
    ans = fm.MCMC(init, fam, 900, nchains=C, kernel=fm.kernel_adapt(), seed=20260317)
A = ans.as_array()          # [C][T][k]
for t in (1, 100, 400, 500, 510, 520, 550, 600, 700, 899):
    th = A[:, t, :]
    b = np.minimum(np.abs(th) @ cmax, np.linalg.norm(th, axis=1) * rmax)
    blk = b.reshape(8, 128).max(axis=1)
    print(t, "bound quantiles", np.round(np.quantile(b, [0, .25, .5, .75, .9, .99, 1]), 2), "chains<=11.15:", int((b <= 11.15).sum()), "CTAs all<=11.15:", int((blk <= 11.15).sum()),
          "true max|eta| (chain 0):", round(float(np.abs(X @ th[0]).max()), 2))
