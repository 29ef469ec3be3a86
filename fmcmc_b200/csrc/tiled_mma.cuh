// tiled_mma.cuh — observation-tiled likelihood kernel, FP64 tensor-core (DMMA) variant.
//
// Same contract, grid and TMA pipeline as tiled_loglik_kernel (tiled.cuh), but the X.Theta
// contraction runs on mma.sync.aligned.m8n8k4.f64 (SASS DMMA; tcgen05 has no FP64 kind):
//   M = 8 observations, N = 8 chains, K = 4 parameters per instruction.
// The FP64 roof is the same as DFMA's (profiles/r01_microbench_fp64_pipes.txt: 37.1 TF/s either
// way, shared datapath) — what changes is everything AROUND the FP64 pipe: one conflict-free
// LDS.64 fetches an 8x4 A fragment that feeds NT MMAs (= 8*NT DFMA-equivalents per lane), against
// one broadcast LDS.128 per 4 DFMAs in the lane<->chain kernel, whose shared-memory wavefronts ran
// at 58 % of peak and cost ~30 % of the FP64 issue slots (profiles/r01_v1_tiled_loglik_ncu.txt,
// profiles/r01_microbench_fp64_issue.txt).  It also scales to p_x = 128 (config 5): the B
// fragments (Theta) of NT chain tiles stay in registers, NT * PB/4 <= 64 doubles per lane.
//
// Fragment layout (PTX ISA, m8n8k4 .f64), g = lane / 4, t = lane % 4:
//   A[row = g][col = t]   = X[obs0 + g][4 s + t]          (shared memory, column stride CS)
//   B[row = t][col = g]   = Theta[chain_tile*8 + g][4 s + t]   (registers, loaded once per launch)
//   C[row = g][col = 2t + {0,1}] = eta[obs0 + g][chain_tile*8 + 2t + {0,1}]
// Every lane therefore finishes 2 chains x 1 observation per (obs tile, chain tile): the family
// epilogue runs on the C registers in place and accumulates per (chain tile, column) privately;
// one shuffle reduction over g at the very end.  Fixed order => deterministic.
#pragma once
#include "tiled.cuh"

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  // volatile: keeps the source order (k-step outer, chain tile inner) so consecutive DMMAs are independent;
  // left to itself ptxas chains the 8 k-steps of one tile back to back and waits out the latency each time
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

template <int PB>
struct MmaGeom {
  static constexpr int TR = PB <= 32 ? 128 : (PB == 64 ? 64 : 32);  // observations per pipeline stage
  // column stride == 4 (mod 16) doubles: an LDS.64 is served per half-warp (g = 0..3 | 4..7, t = 0..3), and
  // bank(t, g) = (8 t + 2 g) mod 32 is then distinct inside each half (TR + 8 gave 2-way conflicts: ncu r01)
  static constexpr int CS = TR + 4;
  static constexpr int KS = PB / 4;        // k-steps
  static constexpr int STAGE_DOUBLES = PB * CS + TR;
};

// ---- tile-major copy of X / y in HBM ---------------------------------------------------------------
// The DMMA kernel's pipeline stage is stored in HBM exactly as it sits in shared memory: for tile i,
//   Xt[i] = { PB columns x CS rows (rows >= TR and columns >= p_x are zero padding), y[TR] }
// so that ONE cp.async.bulk of STAGE_DOUBLES * 8 bytes (35-37 KB) fills a stage.  Built once per model
// (fmcmc_run, first DMMA launch).  Without it a stage takes p_x + 1 column copies of TR * 8 bytes
// (127 x 256 B at p_x = 127), all issued by one thread: that was the bottleneck of config 5 (38 % of the
// FP64 roof, profiles/r01_bench_v2_dmma_cfg5_first.json).
template <int PB>
__global__ void __launch_bounds__(256) pack_tiles_kernel(const double* __restrict__ X, const double* __restrict__ y,
                                                         long long n, long long ld, int p_x, double* __restrict__ Xt) {
  using G = MmaGeom<PB>;
  const long long tile = blockIdx.x, row0 = tile * G::TR;
  double* dst = Xt + (size_t)tile * G::STAGE_DOUBLES;
  for (int e = threadIdx.x; e < G::STAGE_DOUBLES; e += blockDim.x) {
    double v = 0.0;
    if (e < PB * G::CS) {
      const int j = e / G::CS, r = e % G::CS;
      if (j < p_x && r < G::TR && row0 + r < n) v = X[(size_t)j * ld + row0 + r];
    } else {
      const int r = e - PB * G::CS;
      if (row0 + r < n) v = y[row0 + r];
    }
    dst[e] = v;
  }
}

template <int PB>
__host__ __device__ inline size_t tiled_mma_smem_bytes(int family) {
  return 128 + (size_t)TL_STAGES * MmaGeom<PB>::STAGE_DOUBLES * sizeof(double) +
         (family == FMCMC_FAMILY_LOGISTIC ? (size_t)FM_SP_ENTRIES * 16 : 0);
}

// OSPLIT = false: the CTA's warps own different chains (NWARPS * NT * 8 chains per CTA) and every warp walks
//                 all observations of the CTA's slice — the many-chain mapping (configs 3, 5).
// OSPLIT = true : all warps share the same NT * 8 chains and split the observation tiles of a stage between
//                 them — the few-chain mapping (the reference's typical 1-4 chains on a huge n).  With <= 16
//                 chains the FP64 work per observation drops below the time HBM needs to deliver it, and the
//                 kernel becomes HBM-bound: X streams through the TMA pipeline at the memory roof.
// PIPE = true: software pipelining inside the warp — the DMMAs of observation block t+1 are issued before the
//               epilogue of block t (two C-fragment sets, ping-pong), so the FP64 pipe has independent work
//               while the epilogue's dependent polynomial chains wait out their latency.
template <int FAMILY, int PB, bool YBIN, int NWARPS, int NT, int MO, bool OSPLIT, bool PIPE = false>
__global__ void __launch_bounds__(NWARPS * 32, 1)
tiled_loglik_mma_kernel(ModelParams mp, const double* __restrict__ prop, const double* __restrict__ prop_u, int C,
                        TiledBuffers tb, const int* __restrict__ err) {
  using G = MmaGeom<PB>;
  constexpr int TR = G::TR, CS = G::CS, KS = G::KS, STAGE_DOUBLES = G::STAGE_DOUBLES;
  constexpr int NTHREADS = NWARPS * 32;
  constexpr int CPB = OSPLIT ? NT * 8 : NWARPS * NT * 8;  // chains per CTA
  constexpr int OSTEP = OSPLIT ? NWARPS * 8 * MO : 8 * MO;  // observations between two blocks of one warp
  static_assert(NT * KS <= 64, "B fragments must fit in registers");
  static_assert(TR % OSTEP == 0, "tile rows must be a multiple of the observation block");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + TL_STAGES;
  double* stage0 = reinterpret_cast<double*>(smem_raw + 128);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  // 1-D grid of gx * cb CTAs, chain block fastest: CTAs that run at the same time stream the SAME
  // observation slices, so X comes from HBM once per step and from L2 for the other chain blocks
  const int chain_block = (int)(blockIdx.x % (unsigned)tb.cb), slice = (int)(blockIdx.x / (unsigned)tb.cb);
  double2* sp_tab = reinterpret_cast<double2*>(stage0 + (size_t)TL_STAGES * STAGE_DOUBLES);
  if (FAMILY == FMCMC_FAMILY_LOGISTIC) {
    const double2* gt = reinterpret_cast<const double2*>(mp.sp_tab);
    for (int e = tid; e < FM_SP_ENTRIES; e += NTHREADS) sp_tab[e] = gt[e];
  }

  const long long ntiles = (mp.n + TR - 1) / TR;
  const int p_x = mp.p_x;

  if (tid == 0) {
    for (int s = 0; s < TL_STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NWARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  // With a handful of chains this kernel is HBM-bound and X (256 MB at n = 1e6, p = 32) does not fit the 126 MB L2: streamed
  // with the default policy every MH row evicts all of it for the next one.  The first l2_keep_tiles tiles are fetched
  // evict_last, the others evict_first, so that part of X is served from L2 on every row after the first.
  const bool l2_hints = tb.l2_keep_tiles > 0;
  uint64_t pol_keep = 0, pol_stream = 0;
  if (l2_hints && tid == 0) { pol_keep = l2_policy_evict_last(); pol_stream = l2_policy_evict_first(); }
  auto issue = [&](long long tile, int s) {  // executed by thread 0 only: one bulk copy per stage (tile-major Xt)
    constexpr uint32_t BYTES = (uint32_t)(STAGE_DOUBLES * sizeof(double));
    mbar_expect_tx(&full[s], BYTES);
    if (l2_hints)
      bulk_g2s_hint(stage0 + (size_t)s * STAGE_DOUBLES, mp.Xt + (size_t)tile * STAGE_DOUBLES, BYTES, &full[s],
                    tile < tb.l2_keep_tiles ? pol_keep : pol_stream);
    else
      bulk_g2s(stage0 + (size_t)s * STAGE_DOUBLES, mp.Xt + (size_t)tile * STAGE_DOUBLES, BYTES, &full[s]);
  };

  // ---- pipeline prologue: X does not depend on the head kernel, so the first stages are requested BEFORE waiting for it
  // (programmatic dependent launch: this CTA may be resident while the head kernel of the row still runs) ----
  const long long first = slice, step = tb.gx;
  int n_issued = 0;
  if (tid == 0) {
    long long tl = first;
    for (int s = 0; s < TL_STAGES && tl < ntiles; s++, tl += step, n_issued++) issue(tl, s);
  }
  pdl_launch_dependents();
  pdl_wait();
  if (err[0] != 0) {  // a previous row failed: let the copies in flight land, then leave
    if (tid == 0)
      for (int s = 0; s < n_issued; s++) mbar_wait(&full[s], 0u);
    return;
  }

  // ---- B fragments: Theta of this warp's NT chain tiles, in registers for the whole launch ----
  const int icpt = (FAMILY == FMCMC_FAMILY_GAUSSIAN_LM && (mp.flags & FMCMC_MODEL_INTERCEPT)) ? 1 : 0;
  const int chain0 = chain_block * CPB + (OSPLIT ? 0 : warp * (NT * 8));
  const int ofirst = OSPLIT ? warp * 8 * MO : 0;
  auto theta_of = [&](int col) -> const double* {
    if (col >= tb.ncols) return nullptr;
    return col < C ? prop + (size_t)col * mp.k : prop_u + (size_t)(col - C) * mp.k;
  };
  double B[KS][NT];
  double cinit[NT][2];
#pragma unroll
  for (int ct = 0; ct < NT; ct++) {
    const double* th = theta_of(chain0 + ct * 8 + g);
#pragma unroll
    for (int s = 0; s < KS; s++) {
      const int j = 4 * s + t;
      B[s][ct] = (th && j < p_x) ? th[icpt + j] : 0.0;
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const double* thc = theta_of(chain0 + ct * 8 + 2 * t + h);
      cinit[ct][h] = (icpt && thc) ? thc[0] : 0.0;
    }
  }

  double acc[NT][2];
#pragma unroll
  for (int ct = 0; ct < NT; ct++) acc[ct][0] = acc[ct][1] = 0.0;

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
    const double* ys = Xs + (size_t)PB * CS;
    const long long row0 = tile * TR;
    const int valid = (int)min((long long)TR, mp.n - row0);  // < TR only for the last tile
    const double* Ag = Xs + t * CS + g;                        // this lane's A-fragment column / row
    auto mma_block = [&](double (&c)[MO][NT][2], int o) {  // MO observation tiles x NT chain tiles = MO*NT independent DMMA chains
#pragma unroll
      for (int mo = 0; mo < MO; mo++)
#pragma unroll
        for (int ct = 0; ct < NT; ct++) { c[mo][ct][0] = cinit[ct][0]; c[mo][ct][1] = cinit[ct][1]; }
#pragma unroll
      for (int ks = 0; ks < KS; ks++) {
        double a[MO];
#pragma unroll
        for (int mo = 0; mo < MO; mo++) a[mo] = Ag[ks * 4 * CS + o + 8 * mo];
#pragma unroll
        for (int mo = 0; mo < MO; mo++)
#pragma unroll
          for (int ct = 0; ct < NT; ct++) dmma_m8n8k4(c[mo][ct][0], c[mo][ct][1], a[mo], B[ks][ct]);
      }
    };
    auto epilogue_block = [&](double (&c)[MO][NT][2], int o) {
#pragma unroll
      for (int mo = 0; mo < MO; mo++) {
        const double yv = ys[o + 8 * mo + g];
#pragma unroll
        for (int ct = 0; ct < NT; ct++) {
          acc[ct][0] += tile_term<FAMILY, YBIN>(c[mo][ct][0], yv, sp_tab);
          acc[ct][1] += tile_term<FAMILY, YBIN>(c[mo][ct][1], yv, sp_tab);
        }
      }
    };
    if (valid == TR) {
      if (PIPE) {
        static_assert(!PIPE || ((TR / OSTEP) % 2 == 0), "ping-pong needs an even number of blocks per stage");
        double c0[MO][NT][2], c1[MO][NT][2];
        mma_block(c0, ofirst);
#pragma unroll 1
        for (int o = ofirst; o < TR; o += 2 * OSTEP) {
          mma_block(c1, o + OSTEP);
          epilogue_block(c0, o);
          if (o + 2 * OSTEP < TR) mma_block(c0, o + 2 * OSTEP);
          epilogue_block(c1, o + OSTEP);
        }
      } else {
#pragma unroll 1
        for (int o = ofirst; o < TR; o += OSTEP) {
          double c[MO][NT][2];
          mma_block(c, o);
          epilogue_block(c, o);
        }
      }
    } else {
      for (int o = OSPLIT ? warp * 8 : 0; o < valid; o += OSPLIT ? NWARPS * 8 : 8) {  // rows >= valid: zero padding, selected away
        double c[NT][2];
#pragma unroll
        for (int ct = 0; ct < NT; ct++) { c[ct][0] = cinit[ct][0]; c[ct][1] = cinit[ct][1]; }
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
          const double a = Ag[ks * 4 * CS + o];
#pragma unroll
          for (int ct = 0; ct < NT; ct++) dmma_m8n8k4(c[ct][0], c[ct][1], a, B[ks][ct]);
        }
        const bool live = (o + g) < valid;
        const double yv = ys[o + g];
#pragma unroll
        for (int ct = 0; ct < NT; ct++) {
          const double v0 = tile_term<FAMILY, false>(c[ct][0], yv, sp_tab);
          const double v1 = tile_term<FAMILY, false>(c[ct][1], yv, sp_tab);
          acc[ct][0] += live ? v0 : 0.0;
          acc[ct][1] += live ? v1 : 0.0;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // ---- reduce over the 8 observation rows (lanes with equal t), write the CTA's partial sums ----
  // Sharded over observations: the sums go to every rank's exchange buffer (peer stores over NVLink), at
  // slice index rank * gx + slice, then the last CTA of the grid publishes the step flag everywhere.
  const int world = tb.sx.world > 1 ? tb.sx.world : 1;
  auto store_partial = [&](int col, double v) {
    if (world == 1) {
      tb.partial[(size_t)slice * tb.ncols + col] = v;
    } else {
      const size_t off = (size_t)(tb.sx.step & 1ULL) * tb.sx.parity_stride +
                         ((size_t)tb.sx.rank * tb.gx + slice) * tb.ncols + col;
      for (int pg = 0; pg < world; pg++) tb.sx.peer_partial[pg][off] = v;
    }
  };
  double* red = stage0;  // OSPLIT: [NWARPS][NT * 8] cross-warp staging; the pipeline stages are drained by now
  if (OSPLIT) __syncthreads();
#pragma unroll
  for (int ct = 0; ct < NT; ct++)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      double v = acc[ct][h];
      v += __shfl_xor_sync(FM_FULL, v, 4);
      v += __shfl_xor_sync(FM_FULL, v, 8);
      v += __shfl_xor_sync(FM_FULL, v, 16);
      const int lc = ct * 8 + 2 * t + h;
      if (OSPLIT) {
        if (g == 0) red[warp * (NT * 8) + lc] = v;
      } else {
        const int col = chain0 + lc;
        if (g == 0 && col < tb.ncols) store_partial(col, v);
      }
    }
  if (OSPLIT) {
    __syncthreads();
    if (tid < NT * 8) {
      double v = 0.0;
      for (int w = 0; w < NWARPS; w++) v += red[w * (NT * 8) + tid];  // fixed order: deterministic
      const int col = chain0 + tid;
      if (col < tb.ncols) store_partial(col, v);
    }
  }
  if (world > 1) {  // publish: every CTA fences its peer stores, the last one raises the flags
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      const unsigned int total = gridDim.x;
      const unsigned int prev = atomicAdd(tb.sx.done, 1u);
      if (prev == total - 1) {
        *tb.sx.done = 0u;
        __threadfence_system();
        for (int pg = 0; pg < world; pg++) {
          volatile unsigned long long* fl = tb.sx.peer_flags[pg] + tb.sx.rank;
          *fl = tb.sx.step;
        }
        __threadfence_system();
      }
    }
  }
}
