#!/usr/bin/env python
"""bench.py — MH chain-steps/s of the multi-chain Metropolis-Hastings hot path on B200.

Workload (BASELINE.json configs[2], the configuration the north_star target is quoted on):
synthetic logistic regression n = 1e6, p = 32 (column 1 == 1), N(0, 2^2) prior, 1024 chains per GPU,
kernel_adapt() (Haario AM, warmup 500, freq 1, recursive covariance).  One "step" = one MH row for
every chain on the GPU = 1024 chain-steps = 1.024e9 chain-step x observation evaluations.

  python bench.py [--gpus N] [--steps K] [--warmup W]        our CUDA path
  python bench.py --impl reference ...                        the reference's CPU path (oracle port, all host threads)

Under torchrun (N > 1) chains are sharded over ranks (weak scaling: 1024 chains per GPU, X replicated),
no data-path collective; one Gelman-Rubin check over NCCL runs after the timed region (untimed).
Prints ONE JSON line on rank 0.
"""
import argparse
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


def make_data(n=N_OBS, p=P_X, seed=DATA_SEED):
    """X[:,0] = 1, X[:,1:] ~ N(0,1)/sqrt(p), beta* ~ N(0,1), y ~ Bernoulli(plogis(X beta*))."""
    rng = np.random.Generator(np.random.PCG64(seed))
    X = np.empty((n, p), order="F")
    X[:, 0] = 1.0
    for j in range(1, p):
        X[:, j] = rng.standard_normal(n) / np.sqrt(p)
    beta = rng.standard_normal(p)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU, help="chains per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-kernel-warmup", action="store_true",
                    help="prime abs_iter instead of running kernel_adapt's 500 warm-up rows (profiling runs)")
    args = ap.parse_args()
    args.steps = max(args.steps, 2)
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist

    import fmcmc_b200 as fm
    from fmcmc_b200 import _abi as A
    from fmcmc_b200.device import DeviceModel

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

    C = args.chains
    K, W = args.steps, args.warmup
    X, y = make_data()
    fam = fm.ll_logistic(X, y, prior_sd=2.0)
    model = DeviceModel(fam, device=local)          # X, y -> HBM once (264 MB)
    k = fam.k
    chain_offset = rank * C
    kern = fm.kernel_adapt()                        # reference defaults: warmup 500, freq 1, eps 1e-4
    spec = kern.to_spec(k)
    dlen = A.state_len(A.KERNEL_ADAPT, k, k)
    istate = np.zeros((C, A.ISTATE_LEN), dtype=np.int64)
    dstate = np.zeros((C, dlen))
    rng = np.random.default_rng(1000 + rank)
    init = torch.empty((C, k), dtype=torch.float64).pin_memory().numpy()
    init[:] = rng.normal(0, 0.1, (C, k))
    seed = 20260317

    def stream(run_index):
        return A.marshal_stream(A.STREAM_PHILOX, seed=seed, run_index=run_index)

    # ---- setup (untimed): kernel_adapt's own warm-up so the timed rows do the full adaptive step
    # (covariance recurrence + Cholesky + mvn proposal) -----------------------------------------------------
    run_idx = 0
    if args.skip_kernel_warmup:
        istate[:, 0] = KERNEL_WARMUP + 1
        model.run(spec, 3, C, initial=init, stream=stream(run_idx), istate=istate, dstate=dstate,
                  chain_offset=chain_offset, outputs=False)
    else:
        model.run(spec, KERNEL_WARMUP + 3, C, initial=init, stream=stream(run_idx), istate=istate, dstate=dstate,
                  chain_offset=chain_offset, outputs=False)
    run_idx += 1
    assert istate[0, 0] > KERNEL_WARMUP

    # ---- W untimed warm-up steps, then K timed steps: inputs resident in HBM ---------------------------------
    model.run(spec, W + 1, C, initial=None, stream=stream(run_idx), chain_offset=chain_offset, outputs=False,
              flags=A.RUN_DEVICE_STATE)
    run_idx += 1
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    t_wall0 = time.perf_counter()
    out = model.run(spec, K + 1, C, initial=None, stream=stream(run_idx), chain_offset=chain_offset, outputs=False,
                    flags=A.RUN_DEVICE_STATE)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    run_idx += 1
    rep = out["report"]
    dev_ms = float(rep.device_ms)                   # CUDA events on the library's launch stream
    hot_ms = float(rep.hot_ms) / max(int(rep.hot_launches), 1)
    launches = int(rep.n_launches)
    accept = int(rep.n_accept) / (C * K)

    # ---- e2e: the public call MCMC() with HOST buffers (H2D of initial + kernel state, D2H of ans / draws /
    # logpost / state inside the timed region); data X stays cached on the device like the closure's data ----
    last = None
    barrier()
    e2e_t0 = time.perf_counter()
    m2 = model.run(spec, K + 1, C, initial=init, stream=stream(run_idx), istate=istate.copy(), dstate=dstate.copy(),
                   chain_offset=chain_offset, outputs=True, want_draws=True)
    torch.cuda.synchronize()
    e2e_sec = time.perf_counter() - e2e_t0
    run_idx += 1
    last = m2["ans"][:, -1, :]
    h2d, d2h = int(m2["report"].h2d_bytes), int(m2["report"].d2h_bytes)

    clocks = sampler.stop() if sampler else None

    # ---- one Gelman-Rubin check across all chains / GPUs (untimed; NCCL all_gather + all_reduce) ------------
    gel_ms, mpsrf = None, None
    try:
        model.store_reset(C, K + 1)
        model.run(spec, K + 1, C, initial=last, stream=stream(run_idx), istate=m2["istate"], dstate=m2["dstate"],
                  chain_offset=chain_offset, outputs=False, flags=A.RUN_APPEND)
        free = np.ones(k, dtype=np.uint8)
        torch.cuda.synchronize()
        g0 = time.perf_counter()
        if world > 1:
            from fmcmc_b200.dist import ChainSharding
            sh = ChainSharding(C * world)
            _, mpsrf = sh.gelman(model, (K + 1) // 2, K + 1, free, C, k, K + 1 - (K + 1) // 2)
        else:
            xb, s2, ws = model.gelman_partials((K + 1) // 2, K + 1, free, C)
            _, mpsrf = model.gelman_finish(K + 1 - (K + 1) // 2, C, k, xb, s2, ws)
        gel_ms = 1e3 * (time.perf_counter() - g0)
    except Exception as e:  # the R-hat of a 20-row window may be degenerate; never fail the bench on it
        mpsrf = f"unavailable: {e}"

    # ---- max over ranks ------------------------------------------------------------------------------------------
    t = torch.tensor([dev_ms, hot_ms, e2e_sec, t_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, hot_ms, e2e_sec, t_wall = [float(v) for v in t.cpu()]

    if rank == 0:
        hbm_peak, peak_src = load_peaks()
        total_chain_steps = C * world * K
        value = total_chain_steps / (dev_ms * 1e-3)
        alg_bytes = 8.0 * N_OBS * (P_X + 1) + 8.0 * C * (3 * k + 2)          # SURVEY §8d, per launch (= per step per GPU)
        achieved = alg_bytes / (hot_ms * 1e-3) / 1e9
        evals = float(N_OBS) * C
        flops = evals * (2 * P_X + 6)                                         # + 2 transcendentals per eval (reported apart)
        fp64_peak = None
        try:
            import ctypes as Ct
            v = Ct.c_double()
            err = Ct.create_string_buffer(256)
            if fm.lib().fmcmc_measure_fp64_peak(local, Ct.byref(v), err, 256) == 0:
                fp64_peak = v.value
        except Exception:
            pass
        line = {
            "metric": "MH chain-steps/sec", "value": value, "unit": "chain-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"logistic n={N_OBS} p={P_X} x {C} chains/GPU, kernel_adapt (BASELINE configs[2])",
                       "n": N_OBS, "p": P_X, "chains_per_gpu": C, "chains_total": C * world,
                       "kernel": "kernel_adapt(warmup=500, freq=1), timed rows are post-warm-up (adapting every row)",
                       "stream": "Philox4x32-10", "path": int(rep.path),
                       "l2": "X (264 MB) is larger than L2 (126 MB) and streamed once per step: no flush needed"},
            "evals_per_s": value * N_OBS,
            "accept_rate": accept,
            "e2e": {"value": C * world * K / e2e_sec, "unit": "chain-steps/s",
                    "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
                    "call": "fmcmc_run via the Python mirror with host numpy buffers (initial pinned): H2D initial + "
                            "kernel state + spec, D2H ans + draws + logpost + kernel state, per bulk of K rows"},
            "gpu_launches": launches,
            "wall_ms_per_step": 1e3 * t_wall / K,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": None, "kernel": "tiled_loglik_kernel<logistic,32>",
                         "peak_source": peak_src, "launch_ms": hot_ms,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "FP64-pipe bound, not HBM bound (SURVEY §8d: AI ~ C/4 flop/B >> ridge ~6): see fp64"},
            "fp64": {"dfma_peak_tflops_measured": fp64_peak,
                     "achieved_tflops_dot_plus_epilogue": flops / (hot_ms * 1e-3) / 1e12,
                     "frac": (flops / (hot_ms * 1e-3) / 1e12 / fp64_peak) if fp64_peak else None,
                     "flops_per_eval": 2 * P_X + 6, "transcendentals_per_eval": 2},
            "gelman": {"mpsrf": mpsrf, "ms": gel_ms, "chains": C * world},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            chains_s = 2 * threads
            probe = cpu_port_run(X, y, chains_s, 3, threads)                   # 2 MH steps: sizes the sample
            rows_s = 1 + max(2, int(round(15.0 / max(probe / 2.0, 1e-3))))     # ~15 s of CPU work
            sec = cpu_port_run(X, y, chains_s, rows_s, threads)
            line["cpu_baseline"] = {"value": chains_s * (rows_s - 1) / sec, "unit": "chain-steps/s", "cores": threads,
                                    "kind": "port",
                                    "sample": f"{chains_s} chains x {rows_s - 1} MH steps over the full n={N_OBS}, "
                                              f"{sec:.1f} s wall on {threads} host threads (C restatement of the "
                                              "reference loop; R is not installed)"}
        print(json.dumps(line))
    model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
