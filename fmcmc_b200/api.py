"""MCMC(): the reference's entry point (R/mcmc.R:325-479) over the CUDA C ABI.

Signature, argument meaning, return types (coda-like mcmc / mcmc.list), error messages and the
bulk / convergence loop (R/mcmc.R:841-1019) follow the reference; the per-step work runs in
libfmcmcb200.so.  `fun` must be a DeviceFamily: closures are a TypeError, never a CPU fallback."""
from __future__ import annotations

import sys
import threading
import warnings

import numpy as np

from . import _abi as A
from . import mcmc_info as info
from .coda import Mcmc, McmcList, append_chains, append_mcpar
from .convergence import (AutoChecker, GelmanChecker, convergence_data_flush, convergence_msg_get, convergence_msg_set)
from .device import DeviceModel
from .dist import current_sharding
from .families import DeviceFamily
from .kernels import FmcmcKernel, kernel_normal


def _message(*parts):
    print("".join(str(p) for p in parts), file=sys.stderr)


def check_initial(initial, nchains):
    """R/checks.R:22-58"""
    if isinstance(initial, McmcList):
        return np.stack([m.data[-1] for m in initial]), initial.varnames
    names = None
    if isinstance(initial, Mcmc):
        names = initial.varnames
        initial = initial.data[-1]
    if isinstance(initial, dict):
        names, initial = list(initial.keys()), list(initial.values())
    a = np.asarray(initial, dtype=np.float64) if not isinstance(initial, str) else None
    if a is None or a.dtype == object:
        raise TypeError("When `initial` is not a numeric vector, it should be a matrix. Right now it is an "
                        f"object of class `{type(initial).__name__}`.")
    if a.ndim <= 1:
        a = a.reshape(-1)
        if nchains > 1:
            warnings.warn("While using multiple chains, a single initial point has been passed via `initial`: c("
                          + ", ".join(f"{v:g}" for v in a) + "). The values will be recycled. Ideally you would "
                          "want to start each chain from different locations.")
        if a.size == 0:
            raise ValueError("The `initial` vector is of length zero.")
        a = np.tile(a, (nchains, 1))
    elif a.ndim == 2:
        if a.shape[0] != nchains:
            raise ValueError(f"The number of rows of `initial` ({a.shape[0]}) must coincide with the number of "
                             f"chains ({nchains}).")
    else:
        raise TypeError("When `initial` is not a numeric vector, it should be a matrix.")
    if names is None:
        names = [f"par{i + 1}" for i in range(a.shape[1])]
    return np.ascontiguousarray(a), names


class FedStream:
    """Verification mode: the host's own draws (R's runif / rnorm in the serial path's order,
    SURVEY App. B).  logu: [nchains][nsteps], z: [nchains][nsteps][kdraw]; one per bulk."""

    def __init__(self, logu, z):
        self.logu, self.z = np.asarray(logu, dtype=np.float64), np.asarray(z, dtype=np.float64)


def _run_bulk(model, initial, nsteps, nchains, burnin, thin, kernel, seed, run_index, fed, names,
              sharding, append, want_draws=True, resident_state=False, into=None, keep_state=False):
    """MCMC_without_conv_checker, R/mcmc.R:485-838 (chains fan out on the device, not in a loop).
    resident_state: the kernel state of the previous bulk is still on the device and stays there (no upload, no download,
    no write-back into the kernel objects: the caller fetches it once after its last bulk)."""
    if nchains < 1:
        raise ValueError("`nchains` must be an integer greater than 1.")
    if burnin >= nsteps:
        raise ValueError(f"-burnin- ({burnin}) cannot be >= than -nsteps- ({nsteps}).")
    if thin >= nsteps:
        raise ValueError(f"-thin- ({thin}) cannot be > than -nsteps- ({nsteps}).")
    if thin < 1:
        raise ValueError("-thin- should be >= 1.")
    total_chains = sharding.total if sharding else nchains
    if total_chains > 1 and not kernel.is_list:                 # R/mcmc.R:526-527 rep_kernel
        kernel._replicate(nchains)
    elif kernel.is_list and len(kernel) != nchains:
        raise ValueError(f"The passed kernel is a list of {len(kernel)} kernels, but this run has {nchains} chains"
                         + (" on this rank" if sharding else "") + ". Pass a fresh kernel or one replicated for "
                         "the same number of chains.")
    elif total_chains == 1 and kernel.is_list:
        raise ValueError("The passed kernel is for MCMC with more than one chain. Right now, -kernel- is of "
                         f"length {len(kernel)}")
    k = model.k
    if initial is not None and initial.shape[1] != k:
        raise ValueError(f"Incorrect length of -initial-: the family has {k} parameters, got {initial.shape[1]}.")
    spec = kernel.to_spec(k)
    istate, dstate = kernel.state_arrays(nchains, k)
    if fed is not None:
        stream = A.marshal_stream(A.STREAM_FED, logu=fed.logu, z=fed.z)
    else:
        stream = A.marshal_stream(A.STREAM_PHILOX, seed=seed, run_index=run_index)
    obs = getattr(info.MCMC_OUTPUT, "obs_sharding", None)
    if obs is not None:
        obs.dist.barrier()                                       # fmcmc_run is collective when observation-sharded
    flags = (A.RUN_APPEND if append else 0) | (A.RUN_DEVICE_STATE if resident_state else 0) | \
        (A.RUN_KEEP_STATE if keep_state else 0)
    out = model.run(spec, nsteps, nchains, initial=initial, burnin=burnin, thin=thin, stream=stream,
                    istate=istate, dstate=dstate if dstate.size and A.state_len(spec["type"], k, kernel._kf) else None,
                    flags=flags, chain_offset=sharding.offset if sharding else 0,
                    want_draws=want_draws, nchains_total=sharding.total if sharding else 0, into=into)
    if not resident_state and not keep_state:
        kernel.absorb_state(k)
    rep = out["report"]
    info.MCMC_OUTPUT.report = rep
    info.MCMC_OUTPUT.reports.append(rep)
    # every chain is a view of the one [chain][row][param] array the device filled
    return McmcList.from_array(out["ans"], rep.first_iter, rep.last_iter, thin, names), out


def _fetch_state(model, kernel, nchains):
    """The resident kernel state -> the kernel objects (R/mcmc.R:629-631), once per MCMC() call."""
    k = model.k
    istate, dstate = kernel.state_arrays(nchains, k)
    model.fetch_state(istate, dstate if dstate.size and A.state_len(kernel.type, k, kernel._kf) else None)
    kernel.absorb_state(k)


def MCMC(initial, fun, nsteps, *, seed=None, nchains=1, burnin=0, thin=1, kernel=None, multicore=False,
         conv_checker=None, cl=None, progress=False, chain_id=1, device=None, fed=None, path=0,
         shard="chains", **dots):
    """Drop-in for fmcmc::MCMC (R/mcmc.R:325-479).

    initial   vector (recycled), nchains x k matrix, or a previous Mcmc / McmcList (restart)
    fun       a DeviceFamily (ll_gaussian_lm / ll_logistic / ll_hier_normal)
    nsteps    rows per chain, including the initial state
    seed      Philox seed (the reference's set.seed(seed)); results do not depend on #GPUs
    kernel    kernel_normal() by default; state is written back into the object
    multicore / cl   accepted for signature compatibility; chains always run concurrently on the
              GPU(s).  Under torchrun (torch.distributed initialised) chains are sharded over ranks
              and each rank returns its own chains.
    fed       FedStream or a list of FedStream (one per bulk): verification mode
    shard     under torchrun: "chains" (default; each rank owns a range of chains, no data-path collective) or
              "observations" (few chains on a huge n: each rank holds a row slice of X / y, every rank runs
              ALL chains and returns the same result; the per-step exchange of partial sums runs inside the
              CUDA kernels over NVLink peer memory)
    """
    if dots:
        raise TypeError("The following arguments passed via -...- are not present in -fun-:\n - "
                        + ",\n - ".join(dots) + ".\nDevice families carry their data (X, y) themselves.")
    if not isinstance(fun, DeviceFamily):
        raise TypeError("`fun` must be one of the built-in device log-posterior families "
                        "(ll_gaussian_lm, ll_logistic, ll_hier_normal): an arbitrary closure cannot run on "
                        "the GPU and fmcmc_b200 has no CPU fallback.")
    if isinstance(initial, McmcList) and nchains != len(initial):
        raise ValueError("The parameter `nchains` must equal the number of chains passed by `initial`.")
    if multicore and nchains == 1:
        raise ValueError("When `multicore = TRUE`, `nchains` should be greater than 1.")
    if kernel is None:
        kernel = kernel_normal()
    if not isinstance(kernel, FmcmcKernel):
        raise TypeError("`kernel` must be an fmcmc_kernel built by one of the kernel_*() constructors.")
    nsteps, burnin, thin, nchains = int(nsteps), int(burnin), int(thin), int(nchains)
    if nchains < 1:                                             # R/mcmc.R:508-520
        raise ValueError("`nchains` must be an integer greater than 1.")
    if burnin >= nsteps:
        raise ValueError(f"-burnin- ({burnin}) cannot be >= than -nsteps- ({nsteps}).")
    if thin >= nsteps:
        raise ValueError(f"-thin- ({thin}) cannot be > than -nsteps- ({nsteps}).")
    if thin < 1:
        raise ValueError("-thin- should be >= 1.")
    if isinstance(conv_checker, AutoChecker):                  # R/convergence.R:383-386: picked by the number of chains
        conv_checker = conv_checker.gelman if nchains > 1 else conv_checker.geweke
    if conv_checker is not None and isinstance(conv_checker, GelmanChecker) and nchains < 2:
        raise ValueError("Convergence test with the Gelman is only available when `nchains` > 1L.")

    if shard not in ("chains", "observations"):
        raise ValueError('`shard` must be "chains" or "observations".')
    obs_sharding = None
    if shard == "observations":
        from .dist import ObservationSharding, current_sharding as _cs
        if _cs(nchains) is not None:
            obs_sharding = ObservationSharding()
    sharding = current_sharding(nchains) if obs_sharding is None else None
    nlocal = sharding.local if sharding else nchains
    init_all, names = check_initial(initial, nchains)
    if init_all.shape[1] != fun.k:
        raise ValueError(f"Incorrect length of -initial-: the family has {fun.k} parameters, got {init_all.shape[1]}.")
    kernel.to_spec(fun.k) if not kernel.is_list else None       # validates the kernel before touching the GPU
    init_local = init_all[sharding.offset:sharding.offset + nlocal] if sharding else init_all
    if device is None:
        device = sharding.device.index if (sharding and sharding.on_cuda) else (obs_sharding.device if obs_sharding else 0)
    if seed is None:
        seed = int(np.random.SeedSequence().generate_state(2, dtype=np.uint32).astype(np.uint64) @ np.array([1, 1 << 32], dtype=np.uint64))
        if obs_sharding is not None:                            # every rank must draw the same streams
            box = [seed]
            obs_sharding.dist.broadcast_object_list(box, src=0)
            seed = box[0]

    info.MCMC_OUTPUT.clear(nlocal)
    info.MCMC_init(initial=init_all, nsteps=nsteps, seed=seed, nchains=nchains, burnin=burnin, thin=thin,
                   kernel=kernel, conv_checker=conv_checker)
    info.MCMC_OUTPUT.kernel = kernel
    info.MCMC_OUTPUT.obs_sharding = obs_sharding
    if obs_sharding:
        model = DeviceModel(obs_sharding.local_family(fun), device=device)
        obs_sharding.attach(model, fun.n, 2 * nchains)          # kernel_ram evaluates 2 columns per chain
    else:
        # X / y go to HBM (and are packed for the tensor-core path) ONCE per family object and device, like the data an
        # R closure captures: later MCMC() calls on the same family reuse the resident copy (fun.release() frees it)
        model = fun.device_model(device)
    model.set_path(path or 0)
    feds = fed if isinstance(fed, (list, tuple)) else ([fed] if fed is not None else None)
    try:
        if conv_checker is None:
            chains, out = _run_bulk(model, init_local, nsteps, nlocal, burnin, thin, kernel, seed, 0,
                                    feds[0] if feds else None, names, sharding, append=False)
            info.MCMC_OUTPUT.logpost = list(out["logpost"])      # per-chain views of the arrays the device filled
            info.MCMC_OUTPUT.draws = list(out["draws"])
            ans = chains[0] if nchains == 1 else chains
        else:
            ans = _with_conv_checker(model, init_local, nsteps, nlocal, nchains, burnin, thin, kernel, seed, feds,
                                     names, sharding, conv_checker)
    finally:
        if obs_sharding:
            model.close()
    info.MCMC_finalize()
    return ans


def _with_conv_checker(model, initial, nsteps, nlocal, nchains, burnin, thin, kernel, seed, feds, names,
                       sharding, conv_checker):
    """MCMC_with_conv_checker, R/mcmc.R:841-1019."""
    freq = getattr(conv_checker, "freq", None)
    if freq is None:
        freq = nsteps // 2
        warnings.warn(f"The -conv_checker- function has no freq attribute. Default value set to be {freq}")
    if freq * 2 > nsteps:
        freq = 0
    if freq > 0:
        bulks = [freq] * ((nsteps - burnin) // freq)
        if (nsteps - burnin) % freq:
            bulks.append((nsteps - burnin) - sum(bulks))
    else:
        bulks = [nsteps]
    bulks[0] += burnin
    convergence_data_flush()
    device_checker = isinstance(conv_checker, GelmanChecker)
    total_keep = sum(A.rows_kept(b, burnin if i == 0 else 0, thin) for i, b in enumerate(bulks))
    if device_checker:
        model.store_reset(nlocal, total_keep)
    fixed = np.broadcast_to(np.asarray(getattr(kernel, "fixed", False), dtype=bool), (model.k,))
    free_mask = (~fixed).astype(np.uint8)
    converged, i = False, 0
    # ONE set of host arrays for the whole call, filled bulk after bulk by the device (append_chains, R/mcmc.R:947, without
    # re-copying what is already there).  Their pages are touched by a helper thread while the first bulk computes: a fresh
    # anonymous mapping costs ~0.35 ms per MB of page faults when the copy engine is the first to write it.
    k = model.k
    ans_all = np.empty((nlocal, total_keep, k))
    draws_all = np.empty((nlocal, total_keep, k))
    lp_all = np.empty((nlocal, total_keep))
    toucher = None
    if ans_all.nbytes >= (4 << 20):
        toucher = threading.Thread(target=_populate, args=((ans_all, draws_all, lp_all),), daemon=True)
        toucher.start()
    mcpars, done = [], 0

    def so_far():
        """append_chains of every bulk so far as an array-backed mcmc.list (views of the rows filled so far)."""
        start, end, th = append_mcpar(mcpars)
        return McmcList.from_array(ans_all[:, :done], start, end, th, names)

    for i, nst in enumerate(bulks, start=1):
        resident = False
        if i > 1:
            # :909-911 restarts from ans[niter(ans),], the last KEPT row.  When thin divides the previous bulk that is the
            # last row computed, which is still resident on the device (no copy); otherwise it is an earlier row
            prev_rows = bulks[i - 2] - burnin
            initial = None if prev_rows % thin == 0 else np.ascontiguousarray(ans_all[:, done - 1, :])
            burnin = 0
            resident = True                                      # the kernel state never leaves the device between bulks
        chains, out = _run_bulk(model, initial, nst, nlocal, burnin, thin, kernel, seed, i - 1,
                                feds[i - 1] if feds else None, names, sharding, append=device_checker,
                                resident_state=resident, into=(ans_all, draws_all, lp_all, done),
                                keep_state=(i == 1 and len(bulks) > 1))
        done += out["ans"].shape[1]
        mcpars.append(chains.mcpar)
        convergence_msg_set()
        steps = sum(bulks[:i])
        nsamp = done
        if device_checker:
            # the device checker reads the sample store: it only needs the shape of what has been accumulated
            conv_checker._device_ctx = (model, nlocal, free_mask, sharding)
            converged = conv_checker(_Accumulated(append_mcpar(mcpars), nsamp, int(free_mask.sum()),
                                                  sharding.total if sharding else nlocal))
            conv_checker._device_ctx = None
        else:                                                    # arbitrary checker: host path on the samples
            free = np.where(free_mask)[0]
            ans = so_far()
            converged = conv_checker(ans.select(free) if nchains > 1 else ans[0][:, free])
        msg = convergence_msg_get()
        if converged:
            _message("Convergence has been reached with ", steps, " steps. ", "" if msg is None else msg + " ",
                     "(", nsamp, " final count of samples).")
            break
        _message("No convergence yet (steps count: ", steps, "). ", "" if msg is None else msg + " ",
                 "Trying with the next bulk.")
    if toucher is not None:
        toucher.join()
    ans = so_far()
    if i == len(bulks) and not converged:
        _message("No convergence reached after ", sum(bulks[:i]), " steps (", ans.niter(),
                 " final count of samples).")
    if len(bulks) > 1:
        _fetch_state(model, kernel, nlocal)                      # write-back once (R/mcmc.R:629-631)
    info.MCMC_OUTPUT.logpost = list(lp_all[:, :done])
    info.MCMC_OUTPUT.draws = list(draws_all[:, :done])
    return ans[0] if nchains == 1 else ans


def _populate(arrays):
    """Fault the pages of freshly allocated arrays in, writable, WITHOUT touching their contents
    (madvise(MADV_POPULATE_WRITE), Linux >= 5.14): safe while the device is already copying into them.  Best effort."""
    import ctypes
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        libc.madvise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        for a in arrays:
            lo = a.ctypes.data & ~4095
            hi = (a.ctypes.data + a.nbytes + 4095) & ~4095
            libc.madvise(lo, hi - lo, 23)                        # MADV_POPULATE_WRITE; an error (old kernel) is ignored
    except Exception:
        pass


class _Accumulated:
    """What convergence_gelman's device path needs to know about the accumulated mcmc.list: its shape and mcpar (the
    samples themselves are in the device's sample store)."""

    def __init__(self, mcpar, niter, nvar, nchain):
        self.mcpar, self._niter, self._nvar, self._nchain = mcpar, niter, nvar, nchain

    def niter(self):
        return self._niter

    def nvar(self):
        return self._nvar

    def nchain(self):
        return self._nchain
