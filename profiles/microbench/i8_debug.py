import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import fmcmc_b200 as fm
from fmcmc_b200 import _abi as A
from fmcmc_b200.device import DeviceModel
rng = np.random.default_rng(1)
for (n, p, C) in [(128, 4, 4), (1000, 7, 70), (5000, 32, 300)]:
    X = rng.standard_normal((n, p)) / np.sqrt(p); X[:, 0] = 1.0
    beta = rng.standard_normal(p)
    y = (rng.random(n) < 1 / (1 + np.exp(-X @ beta))).astype(float)
    fam = fm.ll_logistic(X, y)
    th = rng.normal(0, 0.5, (C, p))
    eta = X @ th.T
    ref = (np.where(y[:, None] == 1, -np.logaddexp(0, -eta), -np.logaddexp(0, eta))).sum(0) - (th**2).sum(1) / 8
    spec = dict(type=A.KERNEL_NORMAL, k=p, mu=0.0, scale=0.01)
    for path in (3, 4):
        m = DeviceModel(fam); m.set_path(path)
        try:
            g = m.run(spec, 3, C, initial=th, stream=A.marshal_stream(A.STREAM_PHILOX, seed=1))
            lp = g["logpost"][:, 0]
            print(f"n={n} p={p} C={C} path {g['report'].path}: max rel err vs numpy {np.max(np.abs(lp-ref)/np.abs(ref)):.3e}  lp[:3]={lp[:3]} ref[:3]={ref[:3]}", flush=True)
        except Exception as e:
            print("path", path, "FAILED:", e, flush=True)
        m.close()
