"""Where does the e2e call spend its time beyond the kernels?  (cfg3 shape, 100 rows, 1024 chains)"""
import sys, time, numpy as np
sys.path.insert(0, '.')
import bench, fmcmc_b200 as fm
from fmcmc_b200 import _abi as A
from fmcmc_b200.device import DeviceModel
import torch
X, y = bench.make_data()
fam = fm.ll_logistic(X, y)
m = DeviceModel(fam)
C, k, K = 1024, 32, 100
kern = fm.kernel_adapt(); spec = kern.to_spec(k)
dlen = A.state_len(spec["type"], k, k)
ist = np.zeros((C, A.ISTATE_LEN), dtype=np.int64); ist[:, 0] = 501
dst = np.zeros((C, dlen))
init = np.random.default_rng(0).normal(0, 0.1, (C, k))
st = lambda i: A.marshal_stream(A.STREAM_PHILOX, seed=1, run_index=i)
m.run(spec, 3, C, initial=init, stream=st(0), istate=ist, dstate=dst, outputs=False)
for name, kw in [("device-resident state, no outputs", dict(initial=None, flags=A.RUN_DEVICE_STATE, outputs=False)),
                 ("host state in/out, no outputs", dict(initial=init, istate=ist.copy(), dstate=dst.copy(), outputs=False)),
                 ("host state + ans/logpost (no draws)", dict(initial=init, istate=ist.copy(), dstate=dst.copy(), outputs=True, want_draws=False)),
                 ("host state + ans/draws/logpost", dict(initial=init, istate=ist.copy(), dstate=dst.copy(), outputs=True, want_draws=True))]:
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        o = m.run(spec, K + 1, C, stream=st(1 + rep), **kw)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
    r = o["report"]
    print(f"{name:45s} wall {1e3*t:8.2f} ms  device {r.device_ms:8.2f} ms  h2d {r.h2d_bytes/1e6:6.1f} MB  d2h {r.d2h_bytes/1e6:6.1f} MB")
a = torch.empty(53_000_000 // 8, dtype=torch.float64, device="cuda")
for name, mk in [("fresh np.empty", lambda: np.empty(a.numel())), ("touched np.zeros+1", lambda: np.zeros(a.numel()) + 1.0)]:
    h = mk(); ht = torch.from_numpy(h)
    torch.cuda.synchronize(); t0 = time.perf_counter(); ht.copy_(a); torch.cuda.synchronize(); t = time.perf_counter() - t0
    print(f"D2H 53 MB into {name:22s}: {1e3*t:7.2f} ms")
hp = torch.empty(a.numel(), dtype=torch.float64).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter(); hp.copy_(a); torch.cuda.synchronize(); t = time.perf_counter() - t0
print(f"D2H 53 MB into pinned: {1e3*t:7.2f} ms")
t0 = time.perf_counter(); h = np.empty(a.numel()); h[:] = hp.numpy(); t = time.perf_counter() - t0
print(f"memcpy pinned -> fresh np.empty: {1e3*t:7.2f} ms")
m.close()
