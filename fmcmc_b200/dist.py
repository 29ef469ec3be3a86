"""Multi-GPU plumbing: one process per GPU (torchrun), chains sharded across ranks with no
data-path collective; NCCL is used only at convergence checks to
  - all_gather the per-chain means / variances  (psrf),
  - all_reduce the k x k within-chain scatter sum (W, needed for mpsrf).
torch is used for device memory and torch.distributed only — no compute."""
from __future__ import annotations

import os


class ChainSharding:
    def __init__(self, nchains_total: int):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.total = nchains_total
        base, rem = divmod(nchains_total, self.world)
        self.counts = [base + (1 if r < rem else 0) for r in range(self.world)]
        self.offset = sum(self.counts[:self.rank])
        self.local = self.counts[self.rank]
        self.on_cuda = dist.get_backend() == "nccl"
        self.device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))) if self.on_cuda \
            else torch.device("cpu")

    def gelman(self, model, row_begin, row_end, free_mask, nlocal, kf, niter, timings=None):
        """partials on every GPU -> all_gather(xbar, s2) + all_reduce(wsum) -> replicated finish.
        timings: a dict that receives the wall-clock milliseconds of every stage (each closed by a device
        synchronise, so the figures are only meaningful as a breakdown; leave None on the production path)."""
        import time
        torch, dist = self.torch, self.dist
        t = [time.perf_counter()]

        def lap():
            if timings is not None:
                if self.on_cuda:
                    torch.cuda.synchronize(self.device)
                t.append(time.perf_counter())

        if self.on_cuda:
            xbar = torch.empty((nlocal, kf), dtype=torch.float64, device=self.device)
            s2 = torch.empty_like(xbar)
            ws = torch.empty((kf * kf,), dtype=torch.float64, device=self.device)
            model.gelman_partials(row_begin, row_end, free_mask, nlocal,
                                  out=(xbar.data_ptr(), s2.data_ptr(), ws.data_ptr()))
        else:   # gloo (CPU tests): statistics computed by the caller-supplied model on host arrays
            xb, s, w = model.gelman_partials(row_begin, row_end, free_mask, nlocal)
            xbar, s2, ws = torch.from_numpy(xb), torch.from_numpy(s), torch.from_numpy(w.reshape(-1, order="F").copy())
        lap()
        gx, gs = self.all_gather_rows(xbar), self.all_gather_rows(s2)
        lap()
        dist.all_reduce(ws, op=dist.ReduceOp.SUM)
        lap()
        if self.on_cuda:
            torch.cuda.synchronize(self.device)
            out = model.gelman_finish(niter, self.total, kf, gx.data_ptr(), gs.data_ptr(), ws.data_ptr(), dev_in=True)
        else:
            out = model.gelman_finish(niter, self.total, kf, gx.numpy(), gs.numpy(),
                                      ws.numpy().reshape(kf, kf, order="F"))
        lap()
        if timings is not None:
            d = [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])]
            timings.update(stats_ms=d[0], all_gather_ms=d[1], all_reduce_ms=d[2], finish_ms=d[3],
                           all_gather_bytes_per_rank=2 * nlocal * kf * 8, all_reduce_bytes=kf * kf * 8)
        return out

    def pooled_variance(self, model, free_mask):
        """rm_invariant's number over ALL chains (R/convergence.R:171-173): every rank's (count, mean, M2) of its part of
        the store, gathered and merged in rank order with Chan's pairwise formula."""
        mine = [float(v) for v in model.store_pooled(free_mask)]
        box = [None] * self.world
        self.dist.all_gather_object(box, mine)
        return combine_pooled(box)

    def all_gather_rows(self, t):
        """all_gather of [n_r][kf] blocks with (possibly) different n_r, in rank order."""
        torch, dist = self.torch, self.dist
        if len(set(self.counts)) == 1:
            out = torch.empty((self.total, t.shape[1]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(out, t.contiguous())
            return out
        # uneven shards: pad every block to the largest one (collectives need equal sizes), then trim
        mx = max(self.counts)
        pad = torch.zeros((mx, t.shape[1]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        out = torch.empty((self.world * mx, t.shape[1]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, pad)
        return torch.cat([out[r * mx:r * mx + n] for r, n in enumerate(self.counts)], dim=0)


def combine_pooled(triples):
    """[(count, mean, M2), ...] -> the variance (divisor n - 1) of the union, merged left to right."""
    n, mean, m2 = 0.0, 0.0, 0.0
    for nb, mb, qb in triples:
        if nb <= 0:
            continue
        tot = n + nb
        delta = mb - mean
        m2 = m2 + qb + delta * delta * n * nb / tot
        mean = mean + delta * nb / tot
        n = tot
    return m2 / (n - 1.0) if n > 1 else 0.0


def current_sharding(nchains_total: int):
    """ChainSharding when torch.distributed is initialised with world_size > 1, else None."""
    try:
        import torch.distributed as dist
    except Exception:
        return None
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    return ChainSharding(nchains_total)


class ObservationSharding:
    """Few chains on a huge n: rows of X / y are split over the ranks, every rank runs the SAME chains, and the
    per-step exchange of partial log-likelihood sums happens inside the CUDA kernels over NVLink peer memory
    (include/fmcmc_b200.h, fmcmc_shard_*): torch.distributed only carries the 144-byte IPC handles once."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = int(os.environ.get("LOCAL_RANK", 0))

    def row_slice(self, n: int) -> slice:
        """Even-sized contiguous row blocks (16-byte aligned columns), the remainder goes to the last rank."""
        per = (n // self.world) & ~1
        lo = self.rank * per
        return slice(lo, n if self.rank == self.world - 1 else lo + per)

    def local_family(self, family):
        """The rank's row slice of a device family (gaussian_lm / logistic)."""
        from . import _abi as A
        from .families import DeviceFamily
        if family.family == A.FAMILY_HIER_NORMAL:
            raise ValueError("observation sharding supports ll_gaussian_lm and ll_logistic")
        sl = self.row_slice(family.n)
        X = family.X[sl]
        import numpy as np
        return DeviceFamily(family.family, X.shape[0], p_x=family.p_x, X=np.asfortranarray(X), y=family.y[sl],
                            flags=family.flags, hyper=family.hyper)

    def attach(self, model, n_total: int, max_cols: int) -> None:
        import ctypes as C
        from . import _abi as A
        mine = model.shard_alloc(self.world, max_cols, n_total)
        blobs = [None] * self.world
        self.dist.all_gather_object(blobs, bytes(mine))
        handles = [A.ShardHandles.from_buffer_copy(b) for b in blobs]
        model.shard_attach(self.rank, self.world, handles)
        self.dist.barrier()
