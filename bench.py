#!/usr/bin/env python
"""bench.py — MH chain-steps/s of the multi-chain Metropolis-Hastings hot path on B200.

Workload (BASELINE.json configs[2], the configuration the north_star target is quoted on):
synthetic logistic regression n = 1e6, p = 32 (column 1 == 1), N(0, 2^2) prior, 1024 chains per GPU,
kernel_adapt() (Haario AM, warmup 500, freq 1, recursive covariance).  One "step" = one MH row for
every chain on the GPU = 1024 chain-steps = 1.024e9 chain-step x observation evaluations.

  python bench.py [--gpus N] [--steps K] [--warmup W]        our CUDA path
  python bench.py --impl reference ...                        the reference's CPU path (oracle port, all host threads)

Under torchrun (N > 1) chains are sharded over ranks (weak scaling: 1024 chains per GPU, X replicated), no data-path
collective.  The timed region is the reference's bulk loop (R/mcmc.R:901-968) in small: after every `--check-every` steps
(default 10; the reference's default freq is 1000) the kept rows go to the sample store and ONE Gelman-Rubin check runs over
all chains of all GPUs (statistics kernels + NCCL all_gather / all_reduce + finish), so `value` and the 1 -> 8 scaling curve
contain the collective; `stepping_only` is the same region without the checks' time.
`e2e` is the public call: fm.MCMC(initial, family, nsteps, nchains, kernel, conv_checker = convergence_gelman(...)) with host
buffers (the family's X / y already resident in HBM like the data of an R closure; model creation reported apart).
For N > 1 the line also carries `strong_scaling` (the literal BASELINE configs[2]: 1 024 chains in TOTAL, 1 024 / N per GPU);
with the default workload it carries `cfg5` too, one GPU's share of BASELINE configs[4] (FMCMC_BENCH_CFG5=0 skips it).
Prints ONE JSON line on rank 0.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_OBS, P_X, CHAINS_PER_GPU = 1_000_000, 32, 1024
DATA_SEED = 20260317
KERNEL_WARMUP = 500
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed ncu --set full
# capture of this command (profiles/): a profiler figure, so it is a constant here, never measured in the timed run
TRAFFIC_PER_LAUNCH = {("cfg3", 4): 168.4e6,   # profiles/r02_v13_cfg3_tiled_i8_ncu.txt: 162.5 MB read + 5.9 MB written
                      ("cfg5", 4): 9.31e9,    # profiles/r02_v13_cfg5_tiled_i8_ncu.txt: 9.276 GB read + 0.034 GB written (the 64 chain-block CTAs of a slice drift apart in L2)
                      ("cfg3", 3): 278.7e6, ("cfg3", 2): 269.7e6}


def make_data(n=N_OBS, p=P_X, seed=DATA_SEED, block=0):
    """X[:,0] = 1, X[:,1:] ~ N(0,1)/sqrt(p), beta* ~ N(0,1), y ~ Bernoulli(plogis(X beta*)).
    block > 0: another block of rows of the same model (same beta*), for observation-sharded runs."""
    beta = np.random.Generator(np.random.PCG64(seed + 7919)).standard_normal(p)
    rng = np.random.Generator(np.random.PCG64(seed + 104729 * block))
    X = np.empty((n, p), order="F")
    X[:, 0] = 1.0
    for j in range(1, p):
        X[:, j] = rng.standard_normal(n) / np.sqrt(p)
    eta = X @ beta
    y = (rng.random(n) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
    return X, y


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nme, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_port_run(X, y, chains, rows, threads, seed=1):
    """The oracle (C restatement of the reference's loop + kernel_adapt + logistic closure), one chain per
    host thread at a time (the PSOCK decomposition, R/mcmc.R:593-627).  Returns seconds."""
    from fmcmc_b200 import _abi as A
    from oracle import oracle as O
    p = X.shape[1]
    model = A.marshal_model(A.FAMILY_LOGISTIC, X.shape[0], p_x=p, X=X, y=y, hyper=(2.0, 0, 0, 0))
    spec = dict(type=A.KERNEL_ADAPT, k=p, mu=0.0, warmup=0, freq=1, eps=1e-4)   # adapting from row 3 on
    rng = np.random.default_rng(seed)
    init = rng.normal(0, 0.1, (chains, p))
    ist = np.zeros((chains, A.ISTATE_LEN), dtype=np.int64)
    ist[:, 0] = 2     # abs_iter primed past warmup so the covariance recurrence + Cholesky run every step
    O.lib()
    t0 = time.perf_counter()
    O.run(model, spec, init, rows, nchains=chains, stream=A.marshal_stream(A.STREAM_PHILOX, seed=seed),
          istate=ist, threads=threads)
    return time.perf_counter() - t0


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path.  R is not installed (and the
    reference has no native code to compile), so this is the C port in oracle/ with every host thread;
    no interpreter overhead => an optimistic stand-in for R's PSOCK path."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    X, y = make_data()
    chains = threads                      # bounded sample: one chain per host thread, full n
    rows_w, rows_k = args.warmup + 1, args.steps + 1
    if args.warmup > 0:
        cpu_port_run(X, y, chains, max(rows_w, 3), threads)
    sec = cpu_port_run(X, y, chains, max(rows_k, 3), threads)
    steps = max(rows_k, 3) - 1
    val = chains * steps / sec
    line = {
        "impl": "reference", "metric": "MH chain-steps/sec", "value": val, "unit": "chain-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"logistic n={N_OBS} p={P_X} kernel_adapt (BASELINE configs[2])",
                   "chains_in_sample": chains, "n": N_OBS, "p": P_X, "kernel": "kernel_adapt"},
        "evals_per_s": val * N_OBS,
        "cpu_baseline": {"value": val, "unit": "chain-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{chains} chains x {steps} MH steps over the full n={N_OBS} (one chain per "
                                   "host thread; C restatement of the reference, R itself is not installed)"},
        "e2e": {"value": val, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def ess_pooled(ans):
    """ESS of each parameter, pooled over chains: per-chain Geyer initial-positive-sequence estimator on the
    FFT autocovariance, summed over chains.  ans: [C][T][k].  (coda::effectiveSize is third-party and
    unpinned; this estimator is only a reported figure.)"""
    C, T, k = ans.shape
    x = ans - ans.mean(axis=1, keepdims=True)
    nfft = 1 << int(np.ceil(np.log2(2 * T)))
    f = np.fft.rfft(x, n=nfft, axis=1)
    acov = np.fft.irfft(f * np.conj(f), n=nfft, axis=1)[:, :T, :] / T
    var = acov[:, :1, :]
    rho = np.where(var > 0, acov / np.where(var > 0, var, 1.0), 0.0)
    npair = T // 2
    pairs = rho[:, 0:2 * npair:2, :] + rho[:, 1:2 * npair:2, :]
    pos = np.cumprod(pairs > 0, axis=1).astype(bool)             # stop at the first non-positive pair
    tau = -1.0 + 2.0 * np.sum(np.where(pos, pairs, 0.0), axis=1)
    tau = np.maximum(tau, 1.0 / T)
    ess = np.where(var[:, 0, :] > 0, T / tau, 0.0)
    return ess.sum(axis=0)


class Workload:
    """One BASELINE.json config as a synthetic bench workload."""

    def __init__(self, key, args):
        self.key = key
        self.bound = "tensor"
        if key == "cfg3":            # BASELINE configs[2]: the configuration the north_star target is quoted on
            self.family, self.n, self.p_x, self.k = "logistic", N_OBS, P_X, P_X
            self.chains = args.chains or CHAINS_PER_GPU
            self.kernel_name = "kernel_adapt(warmup=500, freq=1), timed rows are post-warm-up (adapting every row)"
            self.kwarm = KERNEL_WARMUP
            self.flops_per_eval, self.transc_per_eval, self.fp64_instr_per_eval = 2 * P_X + 6, 2, 51
            self.label = f"logistic n={self.n} p={self.p_x} x {self.chains} chains/GPU, kernel_adapt (BASELINE configs[2])"
        elif key == "cfg5":          # BASELINE configs[4]: per-GPU share (8192 chains) of the 65536-chain config
            self.family, self.n, self.p_x, self.k = "gaussian", args.n or 10_000_000, 127, 128
            self.chains = args.chains or 8192
            self.kernel_name = "kernel_nmirror(warmup=500, nadapt=4, lb sd = 0), timed rows are post-warm-up"
            self.kwarm = 500
            self.flops_per_eval, self.transc_per_eval, self.fp64_instr_per_eval = 2 * 127 + 4, 0, 127 + 2
            self.label = (f"gaussian_lm n={self.n} k=128 (127 columns + sd) x {self.chains} chains/GPU, kernel_nmirror "
                          "(BASELINE configs[4], one GPU's share)")
        elif key == "few":           # the reference's typical usage: a handful of chains on a large n (README: 1-4 chains)
            self.family, self.n, self.p_x, self.k = "logistic", args.n or N_OBS, P_X, P_X
            self.chains = args.chains or 4
            self.kernel_name = "kernel_adapt(warmup=500, freq=1), timed rows are post-warm-up"
            self.kwarm = KERNEL_WARMUP
            self.flops_per_eval, self.transc_per_eval, self.fp64_instr_per_eval = 2 * P_X + 6, 2, 51
            self.bound = "hbm"
            self.label = (f"logistic n={self.n} p={self.p_x} x {self.chains} chains/GPU, kernel_adapt (few-chain regime: "
                          "observation-split DMMA mapping, HBM-bound)")
        elif key == "cfg4":          # BASELINE configs[3]: lifeexpect hierarchical normal, 4096 chains, kernel_ram
            self.family, self.n, self.p_x, self.k = "hier", 1000, 0, 7
            self.chains = args.chains or 4096
            self.kernel_name = "kernel_ram() defaults (warmup 0: adapting every row, 2nd likelihood when reflected)"
            self.kwarm = 0
            self.flops_per_eval, self.transc_per_eval, self.fp64_instr_per_eval = 3, 0, 2
            self.label = (f"lifeexpect hier_normal (smoke x female cells, k=7) n=1000 x {self.chains} chains/GPU, kernel_ram "
                          "(BASELINE configs[3]); on-chip data, latency-bound")
        elif key == "cfg2":          # BASELINE configs[1]: README model, 4 chains, kernel_normal_reflective, convergence_gelman
            self.family, self.n, self.p_x, self.k = "gaussian", 1000, 1, 3
            self.chains = args.chains or 4
            self.kernel_name = "kernel_normal_reflective(scale=.1, lb=(-5, 0, 0), ub=5) (README.md:378-382)"
            self.kwarm = 0
            self.flops_per_eval, self.transc_per_eval, self.fp64_instr_per_eval = 6, 0, 3
            self.label = (f"README Gaussian LM shape n=1000 k=3 x {self.chains} chains, kernel_normal_reflective + "
                          "convergence_gelman(200) auto-stop (BASELINE configs[1]); on-chip data, latency-bound")
        elif key == "cfg1":          # BASELINE configs[0]: README model shape, 1 chain, kernel_normal(scale=.1)
            self.family, self.n, self.p_x, self.k = "gaussian", 1000, 1, 3
            self.chains = args.chains or 1
            self.kernel_name = "kernel_normal(scale=0.1)"
            self.kwarm = 0
            self.flops_per_eval, self.transc_per_eval, self.fp64_instr_per_eval = 6, 0, 3
            self.label = (f"README Gaussian LM shape n=1000 k=3 x {self.chains} chain(s), kernel_normal(scale=.1) "
                          "(BASELINE configs[0]); on-chip data, latency-bound (serial steps)")
        else:
            raise SystemExit(f"unknown workload {key}")

    def make(self, fm, A, torch, local, rank, kernel_only=False):
        """Returns (family, device_ptrs or None, kernel, init[C][k], host X / y or None); kernel_only: a fresh kernel object."""
        C, k = self.chains, self.k
        if kernel_only:
            if self.key in ("cfg3", "few"):
                return fm.kernel_adapt()
            if self.key == "cfg4":
                return fm.kernel_ram(lb=[np.nan] * 5 + [1e-3, 1e-3])
            if self.key == "cfg2":
                return fm.kernel_normal_reflective(scale=0.1, lb=[-5.0, 0.0, 0.0], ub=5.0)
            if self.key == "cfg1":
                return fm.kernel_normal(scale=0.1)
            return self._kernel5()
        rng = np.random.default_rng(1000 + rank)
        if self.key in ("cfg3", "few"):
            if getattr(self, "obs_shard", None):                 # this rank's rows only (same beta* on every rank)
                world = self.obs_shard
                per = (self.n // world) & ~1
                nloc = per if rank < world - 1 else self.n - per * (world - 1)
                X, y = make_data(n=nloc, block=rank + 1)
                rng = np.random.default_rng(1000)                # identical chains on every rank
            else:
                X, y = make_data(n=self.n)
            fam = fm.ll_logistic(X, y, prior_sd=2.0)
            return fam, None, fm.kernel_adapt(), rng.normal(0, 0.1, (C, k)), (X, y)
        if self.key == "cfg4":
            le = np.load(os.path.join(ROOT, "tests", "golden", "lifeexpect.npz"))
            grp = (2 * le["smoke"] + le["female"]).astype(np.int32)
            fam = fm.ll_hier_normal(le["age"], grp, n_groups=4, gamma_bounds=(0.0, 150.0), estimate_scales=True)
            kern = fm.kernel_ram(lb=[np.nan] * 5 + [1e-3, 1e-3])
            init = np.tile([75.0] * 5 + [5.0, 5.0], (C, 1)) + rng.normal(0, 0.5, (C, k))
            return fam, None, kern, init, None
        if self.key == "cfg2":
            rng = np.random.default_rng(1000)
            X = rng.standard_normal(self.n)
            y = 3.0 + 2.0 * X + rng.normal(0, 4.0, self.n)
            fam = fm.ll_gaussian_lm(X.reshape(-1, 1), y, intercept=True, guard=True)
            kern = fm.kernel_normal_reflective(scale=0.1, lb=[-5.0, 0.0, 0.0], ub=5.0)
            init = np.tile([0.0, 0.0, float(np.std(y, ddof=1))], (C, 1)) + np.abs(rng.normal(0, 0.2, (C, k)))
            return fam, None, kern, np.clip(init, [-5, 0, 0.01], 5.0), None
        if self.key == "cfg1":
            X = rng.standard_normal(self.n)
            y = 3.0 + 2.0 * X + rng.normal(0, 4.0, self.n)
            fam = fm.ll_gaussian_lm(X.reshape(-1, 1), y, intercept=True, guard=True)
            return fam, None, fm.kernel_normal(scale=0.1), np.tile([0.0, 0.0, float(np.std(y, ddof=1))], (C, 1)), None
        # cfg5: 10 GB of X generated directly in HBM (torch is plumbing: device memory + RNG for SYNTHETIC data)
        g = torch.Generator(device="cuda")
        g.manual_seed(DATA_SEED)
        n, p = self.n, self.p_x
        Xd = torch.empty((p, n), dtype=torch.float64, device="cuda")      # column-major n x p (R layout)
        Xd[0].fill_(1.0)
        for j in range(1, p):
            Xd[j].normal_(generator=g)
        beta = torch.randn(p, dtype=torch.float64, device="cuda", generator=g)
        yd = torch.matmul(beta, Xd) + 2.0 * torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
        torch.cuda.synchronize()
        from fmcmc_b200.families import DeviceFamily
        fam = DeviceFamily(A.FAMILY_GAUSSIAN_LM, n, p_x=p, flags=A.MODEL_GUARD)
        self._keep = (Xd, yd)
        centre = np.r_[beta.cpu().numpy(), 2.0]
        lb = np.full(k, np.nan); lb[-1] = 0.0
        self._kernel5 = lambda: fm.kernel_nmirror(mu=centre, scale=2.0 / np.sqrt(n) * 0.3, lb=lb)
        kern = self._kernel5()
        init = centre + rng.normal(0, 2.0 / np.sqrt(n), (C, k))
        return fam, (Xd.data_ptr(), yd.data_ptr(), None), kern, init, None


def _never():
    """convergence_gelman whose threshold cannot be met: every bulk runs, every check is computed."""
    import fmcmc_b200 as fm
    return fm.convergence_gelman(freq=1, threshold=0.0)


def measure(wl, args, env, brief=False, chains_total=None):
    """W warm-up + K timed steps of workload `wl` on this rank's GPU.  Returns the dict of measurements (rank-local; the caller
    reduces over ranks).  brief: device-timed leg only (no e2e / ESS / CPU baseline)."""
    torch, dist, fm, A = env["torch"], env["dist"], env["fm"], env["A"]
    from fmcmc_b200.device import DeviceModel
    world, rank, local = env["world"], env["rank"], env["local"]
    barrier = env["barrier"]
    obs_shard = getattr(wl, "obs_shard", None)
    C, k = wl.chains, wl.k
    K, W = wl.steps, args.warmup
    ce = max(1, min(args.check_every, K)) if args.check_every > 0 else 0
    t0 = time.perf_counter()
    fam, dev_ptrs, kern, init0, host_data = wl.make(fm, A, torch, local, rank)
    t_data = time.perf_counter() - t0
    t0 = time.perf_counter()
    model = DeviceModel(fam, device=local, device_ptrs=dev_ptrs)      # X, y -> HBM once
    fam.__dict__.setdefault("_models", {})[local] = model             # MCMC() finds the resident copy on the family object
    torch.cuda.synchronize()
    t_model = time.perf_counter() - t0
    chain_offset = rank * C
    ntot = chains_total or (C if obs_shard else C * world)
    if obs_shard:
        from fmcmc_b200.dist import ObservationSharding
        ObservationSharding().attach(model, wl.n, 2 * C)
        chain_offset = 0
    spec = kern.to_spec(k)
    dlen = A.state_len(spec["type"], k, k)
    istate = np.zeros((C, A.ISTATE_LEN), dtype=np.int64)
    dstate = np.zeros((C, max(dlen, 1)))
    init = torch.empty((C, k), dtype=torch.float64).pin_memory().numpy()
    init[:] = init0
    seed = 20260317
    run_idx = [0]

    def stream():
        run_idx[0] += 1
        return A.marshal_stream(A.STREAM_PHILOX, seed=seed, run_index=run_idx[0])

    def run(rows, **kw):
        kw.setdefault("chain_offset", chain_offset)
        kw.setdefault("nchains_total", ntot)
        return model.run(spec, rows, C, stream=stream(), **kw)

    # ---- setup (untimed): the kernel's own warm-up so the timed rows do the full adaptive step
    # (kernel_adapt: covariance recurrence + Cholesky + mvn proposal every row) ---------------------------------
    if args.skip_kernel_warmup or wl.key == "cfg5":
        # cfg5: 500 warm-up rows cost minutes of GPU time; the post-warm-up mirror step does the same work per
        # row as a warm-up one (the adaptation is O(k) per chain), so abs_iter is primed past the warm-up instead
        istate[:, 0] = wl.kwarm + 1
        if wl.key == "cfg5":
            istate[:, 1] = A.STATE_INIT | (2 << A.STATE_OBS_SHIFT)
            dstate[:, :k] = np.asarray(spec["mu"]); dstate[:, k:2 * k] = np.asarray(spec["scale"]); dstate[:, 2 * k:] = 0.4
        run(3, initial=init, istate=istate, dstate=dstate, outputs=False)
    else:
        run(wl.kwarm + 3, initial=init, istate=istate, dstate=dstate, outputs=False)
    assert wl.kwarm == 0 or istate[0, 0] > wl.kwarm
    path = None
    free = np.ones(k, dtype=np.uint8)
    sh = None
    if world > 1 and not obs_shard and ce:
        from fmcmc_b200.dist import ChainSharding
        sh = ChainSharding(ntot)
        assert sh.local == C and sh.offset == chain_offset, (sh.local, C, sh.offset, chain_offset)

    def region(nsteps, timings=None):
        """The bulk loop in small: bulks of `ce` steps appended to the store, one R-hat check over ALL chains after each
        (R/mcmc.R:901-968).  Returns (device ms over everything, [run reports], mpsrf of the last check, checks, check ms)."""
        nb = (nsteps + ce - 1) // ce if ce else 1
        if ce:
            model.store_reset(C, nsteps + nb)
        reps, mps, nchk, chk_ms = [], None, 0, 0.0
        model.mark(0)
        done = 0
        while done < nsteps:
            s = min(ce, nsteps - done) if ce else nsteps
            o = run(s + 1, initial=None, outputs=False, flags=A.RUN_DEVICE_STATE | (A.RUN_APPEND if ce else 0))
            reps.append(o["report"])
            done += s
            if ce:
                rows = model.store_rows()
                first = rows // 2
                model.mark(2)
                try:
                    if sh is not None:
                        _, mps = sh.gelman(model, first, rows, free, C, k, rows - first, timings=timings)
                    else:       # one GPU: what MCMC(conv_checker = convergence_gelman()) calls - coda's window + statistics + finish in ONE library call
                        _, mps, _ = model.gelman(free)
                except fm.FmcmcError as e:          # chol(W) may fail on a very short window: the reference warns and goes on
                    mps = f"unavailable: {e}"
                model.mark(3)
                chk_ms += model.elapsed_ms(2, 3)
                nchk += 1
        model.mark(1)
        return model.elapsed_ms(0, 1), reps, mps, nchk, chk_ms

    # ---- W untimed warm-up steps, then K timed steps: inputs resident in HBM ---------------------------------
    region(W)
    if not ce and wl.key != "cfg5" and K > W:
        # without checks the timed region is ONE call of K + 1 rows: let the library size its row buffers for it once, untimed
        # (cudaMalloc of the larger ans / draws / logpost buffers otherwise lands inside the first timed region)
        region(K)
    barrier()
    sampler = ClockSampler(local) if (rank == 0 and not brief) else None
    barrier()
    t_wall0 = time.perf_counter()
    tm = {}
    dev_ms, reps, mpsrf, nchk, chk_ms = region(K, timings=tm)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = float(sum(r.device_ms for r in reps))
    hot_l = int(sum(r.hot_launches for r in reps))
    hot_ms = float(sum(r.hot_ms for r in reps)) / hot_l if hot_l else step_ms / K      # path 1: one fused launch per bulk
    hot_series = [float(r.hot_ms) / r.hot_launches for r in reps if r.hot_launches]    # per call: its timed launches' mean
    launches = int(sum(r.n_launches for r in reps)) + nchk * 6                          # + the 3 statistics + 3 finish kernels of a check
    accept = int(sum(r.n_accept for r in reps)) / (C * K)
    path = int(reps[0].path)
    repeats = []
    for _ in range(0 if brief else 2):                 # run-to-run spread of the same region (reported, not used for `value`)
        barrier()
        rr = region(K)
        repeats.append(rr[0] / K)
        hot_series += [float(q.hot_ms) / q.hot_launches for q in rr[1] if q.hot_launches]
    # ESS of the timed rows on the DEVICE (fmcmc_store_ess: autocovariances + Geyer's initial positive sequence per chain and
    # parameter over the sample store the region just filled), summed over this rank's chains
    ess_dev = None
    if ce and not brief:
        try:
            e, trunc = model.store_ess(0, model.store_rows(), free, C)
            ess_dev = {"per_param_sum_over_chains": e.sum(axis=0), "truncated": trunc, "rows": int(model.store_rows())}
        except fm.FmcmcError as ex:
            ess_dev = {"error": str(ex)}
    res = dict(hot_series=hot_series, ess_dev=ess_dev, dev_ms=dev_ms, step_ms=step_ms, hot_ms=hot_ms, launches=launches, accept=accept, path=path, t_wall=t_wall,
               mpsrf=mpsrf, checks=nchk, check_ms=chk_ms / max(nchk, 1), check_timings=tm, repeats=repeats,
               t_data=t_data, t_model=t_model, C=C, k=k, K=K, ce=ce, ntot=ntot, fam=fam, host_data=host_data,
               model=model, sampler=sampler)
    if brief:
        return res

    # ---- e2e: the PUBLIC call, fm.MCMC(), with HOST buffers: H2D of initial + kernel state, D2H of ans / draws / logpost /
    # state per bulk and the R-hat check after every bulk inside the timed region; X / y stay resident on the family
    # object like the data an R closure captures (model creation: t_model above) ----------------------------------------
    init_all = np.tile(init0, (1 if obs_shard else world, 1))[:ntot] if world > 1 else init0
    if world > 1 and not obs_shard:                     # every rank's own initial rows, in rank order
        parts = [None] * world
        dist.all_gather_object(parts, np.asarray(init0))
        init_all = np.concatenate(parts, axis=0)

    def public_call(nsteps_rows, freq_rows):
        kq = wl.make(fm, A, torch, local, rank, kernel_only=True)
        kq.load_state(istate, dstate, C, k)             # the adaptation continues from the warm-up state
        chk = fm.convergence_gelman(freq=freq_rows, threshold=0.0) if ce else None
        t0 = time.perf_counter()
        with open(os.devnull, "w") as devnull, contextlib.redirect_stderr(devnull):
            a = fm.MCMC(init_all, fam, nsteps_rows, nchains=ntot, kernel=kq, conv_checker=chk, seed=seed + 1, device=local,
                        shard="observations" if obs_shard else "chains")
        torch.cuda.synchronize()
        return a, time.perf_counter() - t0

    # K steps = bulks of ce steps = ce + 1 rows each (the first row of a bulk repeats the last state, quirk D2)
    nb = (K + ce - 1) // ce if ce else 1
    rows_call, freq_rows = (K + nb, ce + 1) if ce else (K + 1, 0)
    e2e_runs, h2d, d2h, ans = [], 0, 0, None
    if obs_shard:
        # observation sharding: MCMC(shard="observations") slices the family itself; here every rank already holds its rows, so
        # the host-buffer call is fmcmc_run through the Python mirror (initial + kernel state in, ans / draws / logpost / state out)
        run(W + 1, initial=init, istate=istate.copy(), dstate=dstate.copy(), outputs=True, want_draws=True)
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            m2 = run(K + 1, initial=init, istate=istate.copy(), dstate=dstate.copy(), outputs=True, want_draws=True)
            torch.cuda.synchronize()
            e2e_runs.append(time.perf_counter() - t0)
        h2d, d2h = int(m2["report"].h2d_bytes), int(m2["report"].d2h_bytes)
        res.update(e2e_sec=float(np.median(e2e_runs)), e2e_runs=e2e_runs, h2d=h2d, d2h=d2h, e2e_rows=K + 1, e2e_freq=0, e2e_bulks=1)
        res["clocks"] = sampler.stop() if sampler else None
        res["ess"] = ess_pooled(m2["ans"][:, 1:, :]) if rank == 0 else None
        return res
    public_call(rows_call, freq_rows)                   # untimed: first-use allocations (output staging, sample store, R-hat scratch)
    for _ in range(1 if wl.key == "cfg5" else 3):       # median of 3 host-side timings; cfg5: 1 (seconds each)
        barrier()
        ans, sec = public_call(rows_call, freq_rows)
        e2e_runs.append(sec)
        h2d = int(sum(r.h2d_bytes for r in fm.MCMC_OUTPUT.reports))
        d2h = int(sum(r.d2h_bytes for r in fm.MCMC_OUTPUT.reports))
    res.update(e2e_sec=float(np.median(e2e_runs)), e2e_runs=e2e_runs, h2d=h2d, d2h=d2h,
               e2e_rows=rows_call, e2e_freq=freq_rows, e2e_bulks=len(fm.MCMC_OUTPUT.reports))
    res["clocks"] = sampler.stop() if sampler else None
    res["ess"] = ess_pooled(ans.as_array()[:, 1:, :] if hasattr(ans, "as_array") else ans.data[None, 1:, :]) if rank == 0 else None
    return res


def rhat_at_scale(env, C=8192, k=128, rows=1000):
    """convergence_gelman at BASELINE configs[4]'s size: C chains x k parameters per GPU, a 500-row window (the second half of
    1 000 accumulated rows, coda's autoburnin).  The rows come from a cheap on-chip model of the same C and k (Gaussian LM,
    80 observations) - the statistics kernels only see the sample store, whatever produced it.  Returns the timings of one check."""
    torch, dist, fm, A = env["torch"], env["dist"], env["fm"], env["A"]
    from fmcmc_b200.device import DeviceModel
    world, rank, local = env["world"], env["rank"], env["local"]
    rng = np.random.default_rng(77 + rank)
    n, p = 80, k - 1
    X = rng.standard_normal((n, p))
    y = X @ rng.standard_normal(p) * 0.1 + rng.normal(0, 1.0, n)
    fam = fm.ll_gaussian_lm(X, y, intercept=False, guard=True)
    lb = np.full(k, np.nan); lb[-1] = 0.0
    spec = fm.kernel_normal_reflective(scale=0.01, lb=lb).to_spec(k)
    init = np.c_[rng.normal(0, 0.1, (C, p)), np.full(C, 1.0) + np.abs(rng.normal(0, 0.05, C))]
    m = DeviceModel(fam, device=local)
    try:
        m.store_reset(C, rows)
        m.run(spec, rows, C, initial=init, flags=A.RUN_APPEND, outputs=False, chain_offset=rank * C, nchains_total=C * world,
              stream=A.marshal_stream(A.STREAM_PHILOX, seed=5))
        free = np.ones(k, dtype=np.uint8)
        first = rows // 2
        out = {}
        for rep in range(2):                            # the second call: buffers allocated, kernels loaded
            tm = {}
            env["barrier"]()
            m.mark(4)
            if world > 1:
                from fmcmc_b200.dist import ChainSharding
                _, mps = ChainSharding(C * world).gelman(m, first, rows, free, C, k, rows - first, timings=tm)
            else:
                t0 = time.perf_counter()
                xb, s2, ws = m.gelman_partials(first, rows, free, C)
                t1 = time.perf_counter()
                _, mps = m.gelman_finish(rows - first, C, k, xb, s2, ws)
                tm = {"stats_ms": 1e3 * (t1 - t0), "finish_ms": 1e3 * (time.perf_counter() - t1)}
            m.mark(5)
            out = {"ms": m.elapsed_ms(4, 5), "breakdown": tm, "mpsrf": mps, "chains_total": C * world, "k": k,
                   "window_rows": rows - first,
                   "note": "one convergence_gelman check: per-chain moments + centred SYRK on the device, all_gather(means, variances) + "
                           "all_reduce(W) over NCCL when N > 1, cross-chain moments on the device, chol(W) + top eigenvalue on the host"}
        return out
    finally:
        m.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg5", "cfg4", "cfg2", "cfg1", "few"],
                    help="cfg3 = BASELINE configs[2] (default, the metric's configuration); cfg5 = configs[4] per-GPU share")
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU")
    ap.add_argument("--nobs", dest="n", type=int, default=None, help="observations (cfg5: default 1e7; few: default 1e6)")
    ap.add_argument("--shard", default="chains", choices=["chains", "observations"],
                    help="multi-GPU mode: chains (default, weak scaling) or observations (workload few: rows of X split over "
                         "ranks, every rank runs all chains, partial sums exchanged over NVLink inside the kernels; strong scaling)")
    ap.add_argument("--check-every", type=int, default=10,
                    help="steps between Gelman-Rubin checks inside the timed region (0 = none; the reference's default freq is 1000)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-kernel-warmup", action="store_true",
                    help="prime abs_iter instead of running the kernel's 500 warm-up rows (profiling runs)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = {"cfg3": 100, "cfg5": 5, "cfg4": 1000, "cfg2": 5000, "cfg1": 10000, "few": 1000}[args.workload]
    args.steps = max(args.steps, 2)
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return reference_arm(args)
    # stdout carries exactly ONE JSON line: anything a library writes to file descriptor 1 meanwhile (NCCL prints its version
    # banner there) is sent to stderr; the descriptor is restored for the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    import fmcmc_b200 as fm
    from fmcmc_b200 import _abi as A

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: fmcmc_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    env = dict(torch=torch, dist=dist, fm=fm, A=A, world=world, rank=rank, local=local, barrier=barrier)

    def reduce_max(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    wl = Workload(args.workload, args)
    wl.steps = args.steps
    obs_shard = args.shard == "observations" and world > 1
    if obs_shard:
        if args.workload != "few":
            raise SystemExit("--shard observations is for --workload few")
        wl.obs_shard = world
    if wl.key in ("cfg1", "cfg2", "cfg4") and args.check_every == 10:
        args.check_every = {"cfg1": 0, "cfg2": 200, "cfg4": 100}[wl.key]    # 1 chain has no R-hat; README's own freq for cfg2
    r = measure(wl, args, env)
    C, k, K, W, path = r["C"], r["k"], r["K"], args.warmup, r["path"]
    fam, host_data, model = r["fam"], r["host_data"], r["model"]
    dev_ms, step_ms, hot_ms, e2e_sec, t_wall = reduce_max(r["dev_ms"], r["step_ms"], r["hot_ms"], r["e2e_sec"], r["t_wall"])
    clocks, ess = r["clocks"], r["ess"]

    # ---- cfg2: the whole public call with the convergence checker (bulks of 200 rows, R-hat on the device after each) ----
    autostop = None
    if wl.key == "cfg2" and rank == 0 and world == 1:
        chk = fm.convergence_gelman(200)
        a0 = time.perf_counter()
        with open(os.devnull, "w") as devnull, contextlib.redirect_stderr(devnull):
            res = fm.MCMC(wl.make(fm, A, torch, local, rank)[3], fam, 5000, nchains=C, kernel=wl.make(fm, A, torch, local, rank, kernel_only=True),
                          conv_checker=chk, seed=11)
        a_sec = time.perf_counter() - a0
        rows = res.niter()
        autostop = {"rows_per_chain_until_converged": rows, "wall_ms": 1e3 * a_sec,
                    "chain_steps_per_s": C * rows / a_sec, "threshold": 1.1, "freq": 200,
                    "call": "MCMC(initial, ll_gaussian_lm, 5000, nchains=4, kernel_normal_reflective, conv_checker=convergence_gelman(200))"}

    # ---- strong scaling: the literal BASELINE configs[2] - 1 024 chains in TOTAL, 1 024 / N per GPU (N > 1 only) --------
    strong = None
    if world > 1 and wl.key == "cfg3" and not obs_shard and not args.chains and CHAINS_PER_GPU % world == 0:
        model.close()
        fam.release()
        ws = Workload("cfg3", args)
        ws.chains, ws.steps = CHAINS_PER_GPU // world, K
        rs = measure(ws, args, env, brief=True, chains_total=CHAINS_PER_GPU)
        s_dev, s_step, s_hot = reduce_max(rs["dev_ms"], rs["step_ms"], rs["hot_ms"])
        rs["model"].close()
        strong = {"chains_total": CHAINS_PER_GPU, "chains_per_gpu": ws.chains, "value": CHAINS_PER_GPU * K / (s_dev * 1e-3),
                  "unit": "chain-steps/s", "ms_per_step": s_dev / K, "stepping_only_ms_per_step": s_step / K,
                  "hot_kernel_ms": s_hot, "path": rs["path"], "rhat_checks_in_region": rs["checks"],
                  "hbm_floor_ms": 1e3 * (N_OBS * 6 * 32) / (load_peaks()[0] * 1e9),
                  "note": "same n, p, kernel; the whole job is 1 024 chains, so each GPU runs 1 024 / N of them on the SAME stepping "
                          "path as one GPU would (fmcmc_run_spec.nchains_total); X is replicated, every GPU still streams all of it "
                          "per step, which is what bounds this mode (hbm_floor_ms = the int8 slices of X / measured HBM bandwidth)"}

    # ---- one GPU's share of BASELINE configs[4] next to the headline (weak over ranks; FMCMC_BENCH_CFG5=0 skips it) ---------
    cfg5 = None
    if wl.key == "cfg3" and os.environ.get("FMCMC_BENCH_CFG5", "1") != "0" and not obs_shard and not args.chains:
        try:
            model.close()
            fam.release()
            a5 = argparse.Namespace(**vars(args))
            a5.chains, a5.n = None, None
            w5 = Workload("cfg5", a5)
            w5.steps = 5
            r5 = measure(w5, a5, env, brief=True)
            d5, st5, h5 = reduce_max(r5["dev_ms"], r5["step_ms"], r5["hot_ms"])
            r5["model"].close()
            ev5 = float(w5.n) * w5.chains
            macs5 = ev5 * 15 * 32 * 4
            sm_clock5 = (clocks or {}).get("sm_mhz") or 1965.0
            t_t5 = macs5 / (148 * 7710.0 * sm_clock5 * 1e6)
            t_f5 = ev5 * 3 / 32.0 / (148 * 2 * sm_clock5 * 1e6)
            cfg5 = {"workload": w5.label, "value": w5.chains * world * w5.steps / (d5 * 1e-3), "unit": "chain-steps/s",
                    "steps": w5.steps, "warmup": W, "ms_per_step": d5 / w5.steps, "stepping_only_ms_per_step": st5 / w5.steps,
                    "evals_per_s": w5.chains * world * w5.steps / (d5 * 1e-3) * w5.n, "path": r5["path"],
                    "rhat_checks_in_region": r5["checks"], "rhat_check_ms": r5["check_ms"], "rhat_check_timings": r5["check_timings"],
                    "mpsrf": r5["mpsrf"], "chains_per_gpu": w5.chains, "chains_total": w5.chains * world,
                    "model_create_s": r5["t_model"], "data_s": r5["t_data"],
                    "rhat_500_row_window": rhat_at_scale(env),
                    "roofline": {"bound": "tensor", "kernel": "tiled_loglik_i8_kernel", "launch_ms": h5,
                                 "floor_ms": 1e3 * (t_t5 + t_f5), "frac": (t_t5 + t_f5) / (h5 * 1e-3),
                                 "note": "two-engine floor: 15 slice pairs x 128 int8 MACs per eval / measured int8 peak + 3 FP64 slots "
                                         "per eval / FP64 pipe rate (time-additive on B200)"}}
        except Exception as e:                          # never lose the headline line over the secondary workload
            cfg5 = {"unavailable": f"{type(e).__name__}: {e}"}

    if rank == 0:
        hbm_peak, peak_src = load_peaks()
        cw = 1 if obs_shard else world                                        # observation sharding: the SAME chains on every rank
        total_chain_steps = C * cw * K
        value = total_chain_steps / (dev_ms * 1e-3)
        n, p_x = wl.n, wl.p_x
        if obs_shard:
            n = int(fam.n)                                                      # rank 0's rows: what ITS kernel streams per launch
        alg_bytes = 8.0 * n * (p_x + 1) + 8.0 * C * (3 * k + 2)               # SURVEY §8d, per launch (= per step per GPU)
        i8_ns = 5                                                             # 8-bit int8 slices per operand of path 4 (6 for kernel_ram / n < 65 536)
        i8_kb = 1 if p_x <= 32 else (2 if p_x <= 64 else 4)
        if path == 4:   # path 4 streams NS int8 slices of X (K padded to 32 / 64 / 128) instead of FP64 X; y only for the Gaussian
            alg_bytes = float(n) * (i8_ns * 32 * i8_kb + (8 if wl.family == "gaussian" else 0)) + 8.0 * C * (3 * k + 2)
        hbm_achieved = alg_bytes / (hot_ms * 1e-3) / 1e9
        evals = float(n) * C
        flops = evals * wl.flops_per_eval                                     # 2 p_x + 6 (+ 2 transcendentals, apart)
        fp64_peak = None
        try:
            import ctypes as Ct
            v = Ct.c_double()
            err = Ct.create_string_buffer(256)
            if fm.lib().fmcmc_measure_fp64_peak(local, Ct.byref(v), err, 256) == 0:
                fp64_peak = v.value
        except Exception:
            pass
        peak_tf = fp64_peak or 37.1
        achieved_tf = flops / (hot_ms * 1e-3) / 1e12
        sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
        pipe_slots = 148 * 2 * sm_clock * 1e6                                   # FP64 warp-instructions / s, whole GPU
        fp64_instr = wl.fp64_instr_per_eval if path != 4 else (7 if wl.family == "logistic" else 3)   # round 2: 10, round 1: 12
        pipe_util = evals * fp64_instr / 32.0 / (hot_ms * 1e-3) / pipe_slots
        kname = {2: "tiled_loglik_kernel", 3: "tiled_loglik_mma_kernel", 4: "tiled_loglik_i8_kernel", 1: "mh_resident_kernel"}[path]
        line = {
            "metric": "MH chain-steps/sec", "value": value, "unit": "chain-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True,
            "scaling": "strong" if obs_shard else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.label, "n": n, "p": p_x, "chains_per_gpu": C, "chains_total": C * cw,
                       "sharding": "observations (rows of X split over ranks; partial sums exchanged over NVLink peer memory "
                                   "inside the kernels)" if obs_shard else "chains",
                       "kernel": wl.kernel_name, "stream": "Philox4x32-10", "path": path,
                       "rhat_check_every_steps": r["ce"],
                       "l2": (f"X ({8e-6 * n * p_x:.0f} MB) is larger than L2 (126 MB) and streamed every step: no flush needed"
                              if path != 1 else "data staged once into shared memory by TMA: on-chip by construction")},
            "evals_per_s": value * wl.n,
            "accept_rate": r["accept"],
            "timed_region": {"what": f"{K} MH steps in bulks of {r['ce']} appended to the sample store + one Gelman-Rubin check over all "
                                     f"{C * cw} chains after every bulk" if r["ce"] else f"{K} MH steps (no R-hat: one chain)",
                             "device_ms": dev_ms, "stepping_only_ms": step_ms, "rhat_checks": r["checks"],
                             "rhat_check_ms": r["check_ms"], "rhat_check_breakdown": r["check_timings"] or None,
                             "mpsrf_last_check": r["mpsrf"],
                             "repeat_ms_per_step": r["repeats"],
                             "hot_launch_ms_series": [round(v, 4) for v in r["hot_series"][:64]],
                             "hot_launch_ms_series_note": "the timed launches of the dominant kernel, call after call (timed region first, "
                                                          "then the two repeats): its run-to-run spread as the chains move"},
            "stepping_only": {"value": total_chain_steps / (step_ms * 1e-3), "ms_per_step": step_ms / K,
                              "note": "the same timed region counting only the stepping kernels' CUDA-event time (what round 1 reported as `value`)"},
            "ess_per_s": (float(r["ess_dev"]["per_param_sum_over_chains"].min()) * cw / (dev_ms * 1e-3)
                          if r.get("ess_dev") and "error" not in r["ess_dev"] else
                          (float(ess.min()) * cw / e2e_sec if ess is not None else None)),
            "ess": ({"min_over_params": float(r["ess_dev"]["per_param_sum_over_chains"].min()),
                     "median_over_params": float(np.median(r["ess_dev"]["per_param_sum_over_chains"])),
                     "rows_per_chain": r["ess_dev"]["rows"], "chains": C, "truncated": r["ess_dev"]["truncated"],
                     "method": "fmcmc_store_ess on the device: per (chain, parameter) autocovariances + Geyer's initial positive sequence "
                               "over the rows of the timed region (bulk boundaries repeat a row, quirk D2), summed over rank 0's chains "
                               "(x n_gpus in ess_per_s); time = the timed region"}
                    if r.get("ess_dev") and "error" not in r["ess_dev"] else
                    ({"min_over_params": float(ess.min()), "median_over_params": float(np.median(ess)),
                      "rows_per_chain": r["e2e_rows"] - 1, "chains": C,
                      "method": "host numpy, per-chain Geyer initial-positive-sequence on the e2e call's output; time = the e2e call"}
                     if ess is not None else None)),
            "e2e": {"value": C * cw * K / e2e_sec, "unit": "chain-steps/s",
                    "h2d_bytes_per_step": r["h2d"] / K, "d2h_bytes_per_step": r["d2h"] / K,
                    "call": (f"fm.MCMC(initial, family, nsteps={r['e2e_rows']}, nchains={C * cw}, kernel=<state past warm-up>, "
                             f"conv_checker=convergence_gelman(freq={r['e2e_freq']}, never met))" if r["ce"] else
                             f"fm.MCMC(initial, family, nsteps={r['e2e_rows']}, nchains={C * cw}, kernel)") +
                            f": {r['e2e_bulks']} bulk(s) = {K} MH steps with host numpy buffers - H2D initial + kernel state, D2H ans + "
                            "draws + logpost + kernel state per bulk, R-hat check after every bulk; X / y resident on the family "
                            "object (model_create_s apart); median of the timed calls",
                    "timed_calls_s": r["e2e_runs"], "model_create_s": r["t_model"]},
            "gpu_launches": r["launches"],
            "wall_ms_per_step": 1e3 * t_wall / K,
            "roofline": None, "roofline_other": None,
            "strong_scaling": strong,
            "cfg5": cfg5,
            "autostop": autostop,
            "clocks": clocks,
        }
        # The many-chain kernel is bound by the FP64 pipe (DFMA / DMMA share one 64-lane datapath per SM; tcgen05 has
        # no FP64 kind), not by HBM: arithmetic intensity ~ C/4 flop/B vs a ridge of ~6.  With a handful
        # of chains (workload "few") the same kernel family is HBM-bound instead.
        fp64_roof = {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf, "traffic": TRAFFIC_PER_LAUNCH.get((wl.key, path)),
                     "kernel": kname, "launch_ms": hot_ms,
                     "peak_source": "FP64 DFMA/DMMA peak measured live by fmcmc_measure_fp64_peak (MEASURED_PEAKS.json "
                                    "has no FP64 figure; the bf16 tcgen05 peak does not apply to FP64)",
                     "algorithmic_flops_per_launch": flops,
                     "flops_per_eval": wl.flops_per_eval, "transcendentals_per_eval_not_counted": wl.transc_per_eval,
                     "fp64_pipe_util": pipe_util,
                     "fp64_pipe_util_note": f"{fp64_instr} FP64-pipe instruction slots per eval (DMMA = 8 slots) x evals / "
                                            "(148 SMs x 2 warp-instr/clk x SM clock)"}
        hbm_roof = {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                    "traffic": TRAFFIC_PER_LAUNCH.get((wl.key, path)), "kernel": kname, "launch_ms": hot_ms,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                    "note": "8 n (p_x + 1) + 8 C (3k + 2) bytes per launch (SURVEY §8d).  With many chains this fraction is "
                            "structurally ~1 %: X is read once per step and shared by every resident chain"}
        line["roofline"], line["roofline_other"] = (hbm_roof, fp64_roof) if wl.bound == "hbm" else (fp64_roof, hbm_roof)
        if path == 4:
            # Path 4 runs the contraction on the int8 tensor pipe (exact integer slices) and only the family epilogue on the
            # FP64 pipe; on B200 the two share a datapath and do not overlap (profiles/r01_i8_findings.md), so the floor of a
            # launch is the SUM of both.  `roofline` above keeps the algorithmic FP64 flops against the FP64 peak (comparable
            # with the FP64 kernels; it may exceed 1 because the dot product is no longer on that pipe); this is the two-engine floor.
            int8_peak = 148 * 7710.0 * sm_clock * 1e6                              # MAC/s, measured (i8mma_vs_fp64.cu: 7710 MAC/clk/SM)
            macs = evals * (i8_ns * (i8_ns + 1) / 2) * 32 * i8_kb
            t_tensor = macs / int8_peak
            t_fp64 = evals * fp64_instr / 32.0 / pipe_slots
            line["roofline_two_engine"] = {
                "int8_mac_per_launch": macs, "int8_peak_mac_per_s": int8_peak, "tensor_floor_ms": 1e3 * t_tensor,
                "fp64_slots_per_eval": fp64_instr, "fp64_floor_ms": 1e3 * t_fp64, "floor_ms": 1e3 * (t_tensor + t_fp64),
                "frac": (t_tensor + t_fp64) / (hot_ms * 1e-3),
                "note": "floor = int8 MACs / measured int8 peak + FP64 epilogue slots / FP64 pipe rate (time-additive on B200).  The floor "
                        "counts THIS kernel's own slice pairs and FP64 slots, so it falls whenever the kernel sheds work: see floor_round1_ms",
                # the same launch against ROUND 1's floor (21 slice pairs, 12 FP64 slots per evaluation: 0.97 ms at cfg3), the yardstick
                # VERDICT r1 set its target on (>= 0.70, i.e. <= 1.4 ms per launch at cfg3)
                "floor_round1_ms": 1e3 * (evals * 21 * 32 * i8_kb / int8_peak + evals * (12 if wl.family == "logistic" else 3) / 32.0 / pipe_slots),
                "frac_vs_round1_floor": (evals * 21 * 32 * i8_kb / int8_peak + evals * (12 if wl.family == "logistic" else 3) / 32.0 / pipe_slots) / (hot_ms * 1e-3)}
            # the same launch against a roof that does NOT depend on how this kernel splits its operands: the algorithm's own
            # p_x multiply-adds per eval on the int8 tensor pipe at its measured peak + the algorithm's epilogue flops (6 logistic /
            # 4 Gaussian, transcendentals not counted) at the measured FP64 peak.  Adding slices or instructions cannot raise it.
            t_alg = evals * p_x / int8_peak + evals * (wl.flops_per_eval - 2 * p_x) / (peak_tf * 1e12)
            line["roofline_algorithmic"] = {
                "floor_ms": 1e3 * t_alg, "frac": t_alg / (hot_ms * 1e-3),
                "fp64_flops_over_fp64_peak": achieved_tf / peak_tf,
                "note": "implementation-independent: p_x MACs per eval / measured int8 peak + (flops_per_eval - 2 p_x) epilogue flops per "
                        "eval / measured FP64 peak (transcendentals not counted); fp64_flops_over_fp64_peak = all algorithmic flops as if "
                        "on the FP64 pipe (exceeds 1: the contraction has left that pipe)"}
            if wl.bound != "hbm":
                # The headline `roofline` of a path-4 launch: the algorithmic FP64 flops against the rate at which they would
                # complete with BOTH engines at their measured peaks (int8 tensor pipe for the exact slice products, FP64 pipe
                # for the epilogue), time-additive as measured on B200.  The FP64-peak-only view (which exceeds 1 because the
                # dot product has left that pipe) stays in `roofline_fp64_pipe`.
                line["roofline_fp64_pipe"] = fp64_roof
                peak2 = flops / (t_tensor + t_fp64) / 1e12
                line["roofline"] = dict(fp64_roof, peak=peak2, frac=achieved_tf / peak2,
                                        peak_source="two-engine floor: int8 MACs / measured tcgen05 int8 peak (7710 MAC/clk/SM, "
                                                    "profiles/microbench/i8mma_vs_fp64.cu) + FP64 epilogue slots / FP64 pipe rate "
                                                    "(measured live), time-additive; algorithmic FP64 flops / that time")
            if wl.family == "logistic":
                # every instruction of the epilogue runs on a 16-lane datapath (FP64, IMAD / IMAD.WIDE, SHF, I2F, LDS): the
                # scheduler issues one warp instruction per 2 clk whatever the pipe (ncu: issue-active 47 % of cycles with
                # not-selected warps waiting), so the instruction count, not the FP64 count alone, is what bounds the kernel
                instr = 15                                                      # 7 FP64 + 2 IMAD + LEA.HI.SX32 + IMAD.WIDE + I2F + IMAD (address) + 2 LDS.128
                t_issue = evals * instr / 32.0 * 2.0 / (148 * 4 * sm_clock * 1e6)
                line["roofline_two_engine"].update({
                    "instr_per_eval": instr, "issue_floor_ms": 1e3 * (t_tensor + t_issue),
                    "issue_frac": (t_tensor + t_issue) / (hot_ms * 1e-3),
                    "issue_note": "tensor time + (warp instructions per eval x 2 clk per instruction per scheduler)"})
        if not args.no_cpu_baseline and world == 1 and host_data is not None:
            X, y = host_data
            threads = os.cpu_count() or 1
            chains_s = 2 * threads
            probe = cpu_port_run(X, y, chains_s, 3, threads)                   # 2 MH steps: sizes the sample
            rows_s = 1 + max(2, int(round(20.0 / max(probe / 2.0, 1e-3))))     # 10-20 s of CPU work
            sec = cpu_port_run(X, y, chains_s, rows_s, threads)
            line["cpu_baseline"] = {"value": chains_s * (rows_s - 1) / sec, "unit": "chain-steps/s", "cores": threads,
                                    "kind": "port",
                                    "sample": f"{chains_s} chains x {rows_s - 1} MH steps over the full n={n}, "
                                              f"{sec:.1f} s wall on {threads} host threads (C restatement of the "
                                              "reference loop; R is not installed)"}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
