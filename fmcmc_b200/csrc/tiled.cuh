// tiled.cuh — observation-tiled stepping path (path 2) for large n.
//
// Per MH row two kernels run back to back on one stream:
//   tiled_head    one warp per chain: finishes row i-1 (fixed-order reduction of the
//                 per-CTA partial sums -> f(theta1) -> RAM phase B -> accept/reject ->
//                 ans/draws/logpost rows) and proposes row i (propose_warp).
//   tiled_loglik  the hot kernel.  Grid = (observation slices) x (chain blocks of 512).
//                 Each CTA streams its slice of X / y through a 4-stage TMA pipeline
//                 (cp.async.bulk + mbarrier full/empty ring).  Lane <-> chain: every
//                 thread keeps the parameter vectors of 2 chains in registers, reads
//                 the X tile by shared-memory BROADCAST (one LDS.128 feeds 4 DFMAs per
//                 lane), applies the family's per-observation epilogue and accumulates
//                 the chain's partial log-likelihood privately -> no cross-lane
//                 reduction, no atomics, deterministic.  X is read from HBM once per
//                 row per chain block and shared by all 512 chains of the CTA.
//                 TL_RO observations x 2 chains = 2*TL_RO independent evaluations are in
//                 flight per thread so the FP64 pipe (the binding roof) stays fed.
#pragma once
#include "families.cuh"
#include "propose.cuh"

#define TL_THREADS 256
#define TL_RC 2
#define TL_CHAINS (TL_THREADS * TL_RC)
#define TL_TILE 128
#define TL_STAGES 4
#define TL_HEAD_WARPS 4
#ifndef TL_RO
#define TL_RO 4
#endif

#define FM_MAX_PEERS 8
// Observation sharding across GPUs (few chains, huge n): every rank holds a row slice of X / y and runs the SAME
// chains.  The likelihood kernel of rank r stores its per-slice partial sums straight into EVERY rank's exchange
// buffer over NVLink (peer stores) and, once its last CTA is done, raises flag[r] = step on every rank.  The head
// kernel of each rank waits for all flags and reduces the world * gx partials in one fixed order, so all ranks take
// bit-identical decisions with no host-side collective on the step path.  Double-buffered by step parity: a peer
// can be at most one step ahead (its next head needs this rank's next flag).
struct ShardExchange {
  int world, rank;                              // world == 0 / 1: not sharded
  double* peer_partial[FM_MAX_PEERS];           // rank g's buffer: [2][world * gx][ncols_max]
  unsigned long long* peer_flags[FM_MAX_PEERS]; // rank g's flags:  [world]
  unsigned long long step;                      // value the flags take when this launch's partials are complete
  unsigned int* done;                           // local CTA-completion counter of the likelihood kernel
  long long parity_stride;                      // doubles between the two parity blocks
};

struct TiledBuffers {
  ShardExchange sx;
  double* partial;  // [gx_total][ncols]: this step's (parity) block, all ranks' slices when sharded
  int gx_total;     // slices the head kernel reduces (world * gx when sharded, else gx)
  int gx;           // observation slices (gridDim.x of tiled_loglik)
  int gsl;          // path 4: consecutive slices walked by one CTA (0 / 1: one; grid = gx / gsl * cb CTAs)
  long long l2_keep_tiles;  // path 3, one chain block, X larger than L2: tiles [0, l2_keep_tiles) are fetched with the evict_last policy and
                            // stay L2-resident from MH row to MH row, the rest with evict_first (0: no hints)
  int cb;           // chain blocks (DMMA kernel: 1-D grid of gx * cb CTAs, chain block fastest)
  int ncols;        // C (or 2C for kernel_ram: second half = un-reflected proposals)
  int tune;         // only read when built with -DFMCMC_I8_TUNE_HOOKS (profiling experiments, tiled_i8.cuh)
  int exact_core;   // path 4: keep the degree-4 log-cosh core (kernel_ram: its adaptation consumes f itself and amplifies 1e-14)
};

// barriers | TL_STAGES x (PB columns + y) x TL_TILE doubles | logistic: softplus table (softplus.h)
__host__ __device__ inline size_t tiled_smem_bytes(int PB, int family) {
  return 128 + (size_t)TL_STAGES * ((size_t)PB * TL_TILE + TL_TILE) * sizeof(double) +
         (family == FMCMC_FAMILY_LOGISTIC ? (size_t)FM_SP_ENTRIES * 16 : 0);
}

// One observation, one chain: per-observation term of the family.
//   FAMILY logistic, YBIN: y is known to be exactly 0.0 or 1.0 (checked at model creation)
template <int FAMILY, bool YBIN>
__device__ __forceinline__ double tile_term(double e, double y, const double2* __restrict__ sp_tab) {
  if (FAMILY == FMCMC_FAMILY_LOGISTIC) return logistic_term_tab<YBIN>(e, y, sp_tab);
  const double r = y - e;  // Gaussian LM: e already holds the linear predictor incl. intercept
  return r * r;
}

template <int FAMILY, int PB, bool YBIN>
__global__ void __launch_bounds__(TL_THREADS, 1)
tiled_loglik_kernel(ModelParams mp, const double* __restrict__ prop, const double* __restrict__ prop_u, int C,
                    TiledBuffers tb, const int* __restrict__ err) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + TL_STAGES;
  double* stage0 = reinterpret_cast<double*>(smem_raw + 128);
  constexpr int STAGE_DOUBLES = PB * TL_TILE + TL_TILE;
  const int tid = threadIdx.x, lane = tid & 31;
  pdl_launch_dependents();
  pdl_wait();
  if (err[0] != 0) return;
  double2* sp_tab = reinterpret_cast<double2*>(stage0 + (size_t)TL_STAGES * STAGE_DOUBLES);
  if (FAMILY == FMCMC_FAMILY_LOGISTIC) {  // 32 KB, L2-resident after the first CTA; read by generic loads only
    const double2* g = reinterpret_cast<const double2*>(mp.sp_tab);
    for (int e = tid; e < FM_SP_ENTRIES; e += TL_THREADS) sp_tab[e] = g[e];
  }

  const long long ntiles = (mp.ld + TL_TILE - 1) / TL_TILE;
  const int p_x = mp.p_x;

  if (tid == 0) {
    for (int s = 0; s < TL_STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], TL_THREADS / 32);
    }
    mbar_fence_init();
  }
  // columns >= p_x of every stage are never written by TMA: zero them once
  for (int s = 0; s < TL_STAGES; s++)
    for (int e = p_x * TL_TILE + tid; e < PB * TL_TILE; e += TL_THREADS) stage0[(size_t)s * STAGE_DOUBLES + e] = 0.0;
  // make the generic-proxy zero fill visible before async-proxy traffic is consumed
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  auto issue = [&](long long tile, int s) {  // executed by thread 0 only
    const long long row0 = tile * TL_TILE;
    const long long rows = min((long long)TL_TILE, mp.ld - row0);
    const uint32_t bytes = (uint32_t)(rows * 8);
    double* dst = stage0 + (size_t)s * STAGE_DOUBLES;
    mbar_expect_tx(&full[s], bytes * (uint32_t)(p_x + 1));
    for (int j = 0; j < p_x; j++) bulk_g2s(dst + (size_t)j * TL_TILE, mp.X + (size_t)j * mp.ld + row0, bytes, &full[s]);
    bulk_g2s(dst + (size_t)PB * TL_TILE, mp.y + row0, bytes, &full[s]);
  };

  // ---- this thread's two chains: parameters into registers ----------------------
  const int col0 = blockIdx.y * TL_CHAINS + tid, col1 = col0 + TL_THREADS;
  const int icpt = (FAMILY == FMCMC_FAMILY_GAUSSIAN_LM && (mp.flags & FMCMC_MODEL_INTERCEPT)) ? 1 : 0;
  double th0[PB], th1[PB], b00 = 0.0, b01 = 0.0;
  {
    const double* s0 = (col0 < tb.ncols) ? (col0 < C ? prop + (size_t)col0 * mp.k : prop_u + (size_t)(col0 - C) * mp.k) : nullptr;
    const double* s1 = (col1 < tb.ncols) ? (col1 < C ? prop + (size_t)col1 * mp.k : prop_u + (size_t)(col1 - C) * mp.k) : nullptr;
#pragma unroll
    for (int j = 0; j < PB; j++) {
      th0[j] = (s0 && j < p_x) ? s0[icpt + j] : 0.0;
      th1[j] = (s1 && j < p_x) ? s1[icpt + j] : 0.0;
    }
    if (icpt) { b00 = s0 ? s0[0] : 0.0; b01 = s1 ? s1[0] : 0.0; }
  }

  // ---- pipeline prologue -----------------------------------------------------------
  const long long first = blockIdx.x, step = gridDim.x;
  if (tid == 0) {
    long long t = first;
    for (int s = 0; s < TL_STAGES && t < ntiles; s++, t += step) issue(t, s);
  }

  double acc0 = 0.0, acc1 = 0.0;
  long long it = 0;
  for (long long tile = first; tile < ntiles; tile += step, it++) {
    const int s = (int)(it % TL_STAGES);
    const uint32_t ph = (uint32_t)((it / TL_STAGES) & 1);
    if (tid == 0 && it >= 1) {  // refill the stage consumed in the previous iteration
      const long long pit = it - 1;
      const int ps = (int)(pit % TL_STAGES);
      const long long nt = tile - step + (long long)TL_STAGES * step;
      if (nt < ntiles) {
        mbar_wait(&empty[ps], (uint32_t)((pit / TL_STAGES) & 1));
        issue(nt, ps);
      }
    }
    mbar_wait(&full[s], ph);
    const double* Xs = stage0 + (size_t)s * STAGE_DOUBLES;
    const double* ys = Xs + (size_t)PB * TL_TILE;
    const long long row0 = tile * TL_TILE;
    const int valid = (int)min((long long)TL_TILE, mp.n - row0);  // < TL_TILE only for the last tile
    if (valid == TL_TILE) {
#pragma unroll 1
      for (int o = 0; o < TL_TILE; o += TL_RO) {
        double e0[TL_RO], e1[TL_RO];
#pragma unroll
        for (int q = 0; q < TL_RO; q++) { e0[q] = b00; e1[q] = b01; }
#pragma unroll
        for (int j = 0; j < PB; j++) {
#pragma unroll
          for (int q = 0; q < TL_RO; q += 2) {
            const double2 x = *reinterpret_cast<const double2*>(Xs + j * TL_TILE + o + q);  // warp-wide broadcast
            e0[q] = fma(x.x, th0[j], e0[q]);
            e1[q] = fma(x.x, th1[j], e1[q]);
            e0[q + 1] = fma(x.y, th0[j], e0[q + 1]);
            e1[q + 1] = fma(x.y, th1[j], e1[q + 1]);
          }
        }
#pragma unroll
        for (int q = 0; q < TL_RO; q += 2) {
          const double2 yy = *reinterpret_cast<const double2*>(ys + o + q);
          acc0 += tile_term<FAMILY, YBIN>(e0[q], yy.x, sp_tab);
          acc1 += tile_term<FAMILY, YBIN>(e1[q], yy.x, sp_tab);
          acc0 += tile_term<FAMILY, YBIN>(e0[q + 1], yy.y, sp_tab);
          acc1 += tile_term<FAMILY, YBIN>(e1[q + 1], yy.y, sp_tab);
        }
      }
    } else {
      for (int o = 0; o < valid; o++) {
        double e0 = b00, e1 = b01;
#pragma unroll
        for (int j = 0; j < PB; j++) {
          const double x = Xs[j * TL_TILE + o];
          e0 = fma(x, th0[j], e0);
          e1 = fma(x, th1[j], e1);
        }
        const double yv = ys[o];
        acc0 += tile_term<FAMILY, false>(e0, yv, sp_tab);
        acc1 += tile_term<FAMILY, false>(e1, yv, sp_tab);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  if (col0 < tb.ncols) tb.partial[(size_t)blockIdx.x * tb.ncols + col0] = acc0;
  if (col1 < tb.ncols) tb.partial[(size_t)blockIdx.x * tb.ncols + col1] = acc1;
}

// Fixed-order reduction of one column of the partial sums by one warp.
__device__ __forceinline__ double reduce_partials(const TiledBuffers& tb, long long col, int lane) {
  double s = 0.0;
  for (int g = lane; g < tb.gx_total; g += FM_WARP) s += __ldcg(&tb.partial[(size_t)g * tb.ncols + col]);
  return warp_sum(s);
}

// Finishes row `row - 1` and proposes row `row` (1 <= row <= T + 1).
template <int KC>
__global__ void __launch_bounds__(TL_HEAD_WARPS * 32)
tiled_head_kernel(ModelParams mp, KParams kp, StreamParams sp, RunBuffers rb, TiledBuffers tb,
                  const double* initial, long long row, int mat_doubles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long c = (long long)blockIdx.x * TL_HEAD_WARPS + warp;
  // the likelihood kernel of this row may set itself up beside us (it reads the proposals only after its own pdl_wait);
  // everything below reads what the previous likelihood launch wrote
  pdl_launch_dependents();
  pdl_wait();
  if (c >= rb.nchains || rb.err[0] != 0) return;
  const int k = kp.k;
  double* scr = reinterpret_cast<double*>(smem_raw) + (size_t)warp * (4 * k + mat_doubles);
  double* th0 = rb.cur_theta + (size_t)c * k;
  double* th1 = rb.prop + (size_t)c * k;
  double* th1u = rb.prop_u + (size_t)c * k;

  if (row == 1) {  // R/mcmc.R:737-743: theta0 = theta1 = initial, f evaluated by the next loglik launch
    const double* src = initial ? initial + (size_t)c * k : th0;
    for (int j = lane; j < k; j += FM_WARP) {
      const double v = src[j];
      th0[j] = v; th1[j] = v; th1u[j] = v;
    }
    if (lane == 0) {
      rb.istate[c * FMCMC_ISTATE_LEN + 3] = 0;
      rb.chain_flags[c] = 0;
    }
    return;
  }

  ChainCtx cx;
  cx.c = c; cx.theta0 = th0; cx.theta1 = th1; cx.theta1u = th1u; cx.scr = scr;
  cx.ans = rb.ans; cx.ans_stride = (long long)rb.nchains * k;
  cx.mat = mat_doubles ? scr + 4 * k : nullptr;

  // ---- finish row r = row - 1 -------------------------------------------------------
  const long long r = row - 1;
  if (tb.sx.world > 1 && !(r == 1 && !initial)) {  // observation-sharded: all ranks' partials of this step must have landed
    bool ok = true;
    if (lane < tb.sx.world) {
      const volatile unsigned long long* fl = tb.sx.peer_flags[tb.sx.rank] + lane;
      unsigned long long t0 = 0, now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      while (*fl < tb.sx.step) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > 60000000000ULL || rb.err[0] != 0) { ok = false; break; }  // 60 s: a peer died / never started
        __nanosleep(200);
      }
    }
    ok = __all_sync(FM_FULL, ok);
    __threadfence_system();
    if (!ok) { if (lane == 0) set_error(rb.err, FMCMC_EPEER, c + 1, r); return; }
  }
  double f1;
  if (r == 1 && !initial) {
    f1 = rb.cur_f[c];  // continuing from the resident state: f(theta0) is known, no loglik launch was made
  } else {
    const double s = reduce_partials(tb, c, lane);
    f1 = family_finish(mp, th1, s);
  }
  if (r == 1) {
    for (int j = lane; j < k; j += FM_WARP) { rb.ans[(size_t)c * k + j] = th0[j]; rb.draws[(size_t)c * k + j] = th0[j]; }
    for (int a = lane; a < kp.kf; a += FM_WARP) {
      rb.colsum[((size_t)c * kp.kf + a) * 2] = th0[kp.free_idx[a]];
      rb.colsum[((size_t)c * kp.kf + a) * 2 + 1] = 0.0;
    }
    if (lane == 0) {
      rb.logpost[c] = f1;
      rb.cur_f[c] = f1;
    }
  } else {
    double f0 = rb.cur_f[c];
    cx.i = r; cx.f0 = f0;
    if (KC == KC_RAM && (rb.chain_flags[c] & 1)) {  // phase B, R/kernel_ram.R:129-150
      const double su = reduce_partials(tb, (long long)rb.nchains + c, lane);
      const double f1u = family_finish(mp, th1u, su);
      const int rc = ram_adapt_warp(kp, rb, cx, f1u, lane);
      if (rc) { if (lane == 0) set_error(rb.err, rc, c + 1, r); return; }
    }
    bool failed = false;
    unsigned long long n_acc = 0;
    f0 = accept_row_warp(kp, sp, rb, c, r, th0, th1, f0, f1, lane, n_acc, failed);
    if (failed) return;
    if (lane == 0) {
      rb.cur_f[c] = f0;
      if (n_acc) atomicAdd(rb.n_accept, n_acc);
    }
  }
  __syncwarp();

  // ---- propose row `row` ---------------------------------------------------------------
  if (row <= rb.T) {
    cx.i = row;
    cx.f0 = rb.cur_f[c];
    const int rc = propose_warp<KC>(kp, sp, rb, cx, lane);
    if (rc && lane == 0) set_error(rb.err, rc, c + 1, row);
  }
}

// ---- the block-parallel head of kernel_adapt: TPC threads per chain ---------------------------------------------------------
// With a handful of chains (the reference's typical usage) the warp-per-chain head above is a single warp walking the
// covariance recurrence and a 32 x 32 left-looking Cholesky factorisation: ~30 us of dependent instructions per MH row, as long
// as the HBM-bound likelihood launch next to it.  Here a chain owns TPC threads - a whole CTA of 1 024 with few chains (thread
// (i, j) = (lane, warp) holds ONE entry of the k_f x k_f matrices), 128 threads = 4 warps with many (8 chains per CTA; lane i of
// warp w holds the 8 entries (i, w), (i, w + 4) ... (i, w + 28)): the recurrence is one multiply-add per entry, and the
// factorisation runs RIGHT-looking in registers - after column p is final (its warp: square root and divisions), every entry
// (i, j > p) subtracts L[i][p] L[j][p].  Each entry still sees its subtractions in the order p = 0, 1, 2 ... with the same
// unfused operations, so the factor - hence every draw - is bit-identical to the left-looking loop (chol_lower_warp) and to the
// oracle; the critical path is 32 steps of (sqrt, divide, multiply-subtracts, one named barrier among the chain's threads)
// instead of 496 dependent multiply-subtracts.  Finishing the previous row (partial sums, accept / reject, output rows) is the
// job of the chain's first warp, exactly as above.
// Chosen by fmcmc_run for kernel_adapt with the Cholesky draw, bw = 0, k_f <= 32 and at most one chain per SM (TPC = 1 024).  The
// 128-thread instantiation is correct (bit-identical in tests/test_gpu_session3.py when enabled) but slower than the
// warp-per-chain head at 1 024 chains (58 against 33 us per row): not instantiated by the library.
#define TL_HEADC_THREADS 1024
__host__ __device__ inline size_t tiled_headc_smem_bytes(int tpc) {   // per chain: L columns [32][33], x / m / mp / z [4][32], stop flag
  return (size_t)(TL_HEADC_THREADS / tpc) * ((32 * 33 + 4 * 32) * sizeof(double) + 16);
}
template <int TPC>
__global__ void __launch_bounds__(TL_HEADC_THREADS)
tiled_head_adapt_cta_kernel(ModelParams mp, KParams kp, StreamParams sp, RunBuffers rb, TiledBuffers tb,
                            const double* initial, long long row) {
  constexpr int CPB = TL_HEADC_THREADS / TPC;   // chains per CTA
  constexpr int W = TPC / 32;                   // warps per chain
  constexpr int NE = 32 / W;                    // matrix entries per thread: columns w, w + W, ...
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int grp = threadIdx.x / TPC, t = threadIdx.x % TPC, lane = t & 31, warp = t >> 5;
  double* sbase = reinterpret_cast<double*>(smem_raw) + (size_t)grp * (32 * 33 + 4 * 32 + 2);
  double (*s_L)[33] = reinterpret_cast<double (*)[33]>(sbase);   // s_L[p][i] = L[i][p]
  double* s_x = sbase + 32 * 33;
  double* s_m = s_x + 32;
  double* s_mp = s_m + 32;
  double* s_z = s_mp + 32;
  volatile int* s_stop = reinterpret_cast<volatile int*>(s_z + 32);   // != 0: the chain stops here (error recorded, or nothing left to do)
  const long long c = (long long)blockIdx.x * CPB + grp;
  auto chain_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(TPC) : "memory"); };   // the chain's threads only
  pdl_launch_dependents();
  pdl_wait();
  if (rb.err[0] != 0 || c >= rb.nchains) return;   // (whole chains leave: their barriers are their own)
  const int k = kp.k, kf = kp.kf;
  double* th0 = rb.cur_theta + (size_t)c * k;
  double* th1 = rb.prop + (size_t)c * k;
  double* th1u = rb.prop_u + (size_t)c * k;
  if (row == 1) {  // R/mcmc.R:737-743
    if (warp == 0) {
      const double* src = initial ? initial + (size_t)c * k : th0;
      for (int j = lane; j < k; j += FM_WARP) {
        const double v = src[j];
        th0[j] = v; th1[j] = v; th1u[j] = v;
      }
      if (lane == 0) {
        rb.istate[c * FMCMC_ISTATE_LEN + 3] = 0;
        rb.chain_flags[c] = 0;
      }
    }
    return;
  }
  if (t == 0) *s_stop = 0;
  chain_sync();

  // ---- finish row r = row - 1 (the chain's first warp; same steps as tiled_head_kernel) ----
  if (warp == 0) {
    const long long r = row - 1;
    bool stop = false;
    if (tb.sx.world > 1 && !(r == 1 && !initial)) {
      bool ok = true;
      if (lane < tb.sx.world) {
        const volatile unsigned long long* fl = tb.sx.peer_flags[tb.sx.rank] + lane;
        unsigned long long t0 = 0, now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (*fl < tb.sx.step) {
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (now - t0 > 60000000000ULL || rb.err[0] != 0) { ok = false; break; }
          __nanosleep(200);
        }
      }
      ok = __all_sync(FM_FULL, ok);
      __threadfence_system();
      if (!ok) { if (lane == 0) set_error(rb.err, FMCMC_EPEER, c + 1, r); stop = true; }
    }
    if (!stop) {
      double f1;
      if (r == 1 && !initial) f1 = rb.cur_f[c];
      else f1 = family_finish(mp, th1, reduce_partials(tb, c, lane));
      if (r == 1) {
        for (int j = lane; j < k; j += FM_WARP) { rb.ans[(size_t)c * k + j] = th0[j]; rb.draws[(size_t)c * k + j] = th0[j]; }
        for (int a = lane; a < kf; a += FM_WARP) {
          rb.colsum[((size_t)c * kf + a) * 2] = th0[kp.free_idx[a]];
          rb.colsum[((size_t)c * kf + a) * 2 + 1] = 0.0;
        }
        if (lane == 0) { rb.logpost[c] = f1; rb.cur_f[c] = f1; }
      } else {
        double f0 = rb.cur_f[c];
        bool failed = false;
        unsigned long long n_acc = 0;
        f0 = accept_row_warp(kp, sp, rb, c, r, th0, th1, f0, f1, lane, n_acc, failed);
        if (failed) stop = true;
        else if (lane == 0) {
          rb.cur_f[c] = f0;
          if (n_acc) atomicAdd(rb.n_accept, n_acc);
        }
      }
    }
    if (row > rb.T) stop = true;   // the last launch only finishes row T
    if (stop && lane == 0) *s_stop = 1;
  }
  __threadfence_block();
  chain_sync();   // (also publishes the first warp's global writes - theta0, the ans row, colsum - to the chain's threads)
  if (*s_stop) return;

  // ---- propose row `row`: R/kernel_adapt.R:84-182, the steps of propose_warp<KC_ADAPT> ----
  const long long i = row;
  long long* ist = rb.istate + c * FMCMC_ISTATE_LEN;
  long long abs_iter = ist[0], flags = ist[1];
  double* Sigma = rb.dstate + (size_t)c * kp.dlen;
  double* Mean_prev = Sigma + (size_t)kf * kf;
  double* L = rb.work + (size_t)c * rb.worklen;
  int* cflag = rb.chain_flags + c;
  bool dirty = !(*cflag & 2);
  chain_sync();   // every thread has read the chain's state words before anyone updates them below
  ChainCtx cx;
  cx.c = c; cx.i = i; cx.theta0 = th0; cx.theta1 = th1; cx.theta1u = th1u; cx.scr = nullptr; cx.f0 = 0.0;
  cx.ans = rb.ans; cx.ans_stride = (long long)rb.nchains * k; cx.mat = nullptr;
  if (!(flags & FMCMC_STATE_INIT)) {  // :87-115
    for (int e = t; e < kf * kf; e += TPC) Sigma[e] = ((e % kf) == (e / kf)) ? kp.eps : 0.0;
    flags |= FMCMC_STATE_INIT;
    dirty = true;
  }
  if (kp.until > (double)abs_iter && abs_iter > kp.warmup && i > 2 && (i % kp.freq) == 0) {  // :118
    if (!(flags & FMCMC_STATE_HAS_MEAN)) {  // :130-131 (colMeans from the compensated running sums, as in propose_warp)
      if (warp == 0) {
        const double* cs = rb.colsum + (size_t)c * 2 * kf;
        const double nn = (double)(i - 1);
        for (int a = lane; a < kf; a += FM_WARP) {
          const double hi = cs[2 * a], lo = cs[2 * a + 1];
          const double q = xdiv(hi, nn);
          const double rr = fma(-q, nn, hi);
          Mean_prev[a] = xadd(q, xdiv(xadd(rr, lo), nn));
        }
      }
      flags |= FMCMC_STATE_HAS_MEAN;
    }
    const double tt = (double)(abs_iter - kp.freq);  // :144
    if (i - kp.freq < 1 || tt == 0.0) {              // the reference indexes row <= 0 / divides by zero here
      if (t == 0) set_error(rb.err, FMCMC_EUNSUP, c + 1, row);
      return;
    }
    __threadfence_block();
    chain_sync();
    for (long long jj = 0; jj < kp.freq; jj++) {  // rows (i-freq):(i-1), R/recursive.R:78-110
      const double tj = tt + (double)jj;
      if (warp == 0) {
        const double* xr = ans_row(cx, kp, i - kp.freq + jj);
        for (int a = lane; a < kf; a += FM_WARP) {
          const double xa = xr[kp.free_idx[a]];
          const double mpa = Mean_prev[a];
          s_x[a] = xa;
          s_mp[a] = mpa;
          s_m[a] = xdiv(xadd(xmul(mpa, tj), xa), tj + 1.0);  // mean_recursive :126
        }
      }
      chain_sync();
      const double c1 = xdiv(tj - 1.0, tj), c2 = xdiv(1.0, tj);
      for (int e = t; e < kf * kf; e += TPC) {  // cov_recursive :112-118, Sd = 1, eps = 1e-5
        const int a = e % kf, b = e / kf;
        double inner = xsub(xmul(tj, xmul(s_mp[a], s_mp[b])), xmul(tj + 1.0, xmul(s_m[a], s_m[b])));
        inner = xadd(inner, xmul(s_x[a], s_x[b]));
        inner = xadd(inner, xmul(1e-5, a == b ? kp.eps : 0.0));
        Sigma[e] = xadd(xmul(c1, Sigma[e]), xmul(c2, inner));
      }
      if (warp == 0)
        for (int a = lane; a < kf; a += FM_WARP) Mean_prev[a] = s_m[a];
      __threadfence_block();
      chain_sync();
    }
    dirty = true;
  }
  abs_iter += 1;  // :170
  if (dirty) {
    __threadfence_block();
    chain_sync();   // Sigma complete (its entries were written by other threads of the chain)
    const int ri = lane;
    double v[NE];
#pragma unroll
    for (int q = 0; q < NE; q++) {
      const int cj = warp + W * q;
      v[q] = (ri < kf && cj < kf && ri >= cj) ? Sigma[ri + cj * kf] : 0.0;
    }
    for (int p = 0; p < kf; p++) {
      if (warp == p % W) {  // column p is final: pivot, square root, divisions (chol_lower_warp's second pass)
#pragma unroll
        for (int q = 0; q < NE; q++)
          if (q == p / W) {
            const double s = __shfl_sync(FM_FULL, v[q], p);
            if (!(s > 0.0)) {
              if (lane == 0) *s_stop = p + 1;
            } else {
              const double ljj = sqrt(s);
              v[q] = (lane == p) ? ljj : xdiv(v[q], ljj);
              if (lane >= p && lane < kf) s_L[p][lane] = v[q];
            }
          }
      }
      chain_sync();
      if (*s_stop) break;
      const double lip = s_L[p][ri < kf ? ri : 0];
#pragma unroll
      for (int q = 0; q < NE; q++) {
        const int cj = warp + W * q;
        if (ri < kf && cj < kf && ri >= cj && cj > p) v[q] = xsub(v[q], xmul(lip, s_L[p][cj]));
      }
    }
    if (*s_stop) {  // mvrnorm: "'Sigma' is not positive definite"
      if (t == 0) set_error(rb.err, FMCMC_ENOTPD, c + 1, row);
      return;
    }
#pragma unroll
    for (int q = 0; q < NE; q++) {   // cached for the rows that do not re-adapt
      const int cj = warp + W * q;
      if (ri < kf && cj < kf) L[ri + cj * kf] = (ri >= cj) ? v[q] : 0.0;
    }
    if (t == 0) *cflag |= 2;
  }
  if (warp == 0) {
    for (int a = lane; a < kf; a += FM_WARP) s_z[a] = draw_z(sp, rb, cx, a);
    for (int j = lane; j < k; j += FM_WARP) th1[j] = th0[j];
    __syncwarp();
    for (int a = lane; a < kf; a += FM_WARP) {  // :173-180
      double s = 0.0;
      for (int b = 0; b <= a; b++) s = xadd(s, xmul(dirty ? s_L[b][a] : L[a + b * kf], s_z[b]));
      const int w = kp.free_idx[a];
      th1[w] = reflect1(xadd(th0[w], xadd(kp.mu[w], s)), kp.lb[w], kp.ub[w]);
    }
    if (lane == 0) { ist[0] = abs_iter; ist[1] = flags; }
  }
}
