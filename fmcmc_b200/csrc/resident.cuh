// resident.cuh — chain-resident fused MH kernel (path 1).
//
// One launch runs ALL rows of a bulk: the whole loop of R/mcmc.R:726-783 lives on
// the device.  Data (X, y, group) is staged once into shared memory by TMA bulk
// copies (UBLKCP) and reused by every chain of the CTA for every step.  A chain is
// owned either by one warp (WPC = true, many chains: configs 2/4) or by a whole CTA
// (WPC = false, few chains: config 1).  Per step the owning group
//   1. draws the proposal (Philox or fed stream) and reflects it   (propose_warp)
//   2. evaluates the family's log-posterior over the n observations (FP64 FMA,
//      warp-shuffle + shared-memory reduction in a fixed order => deterministic)
//   3. makes the MH accept/reject decision and writes ans/draws/logpost rows.
#pragma once
#include "families.cuh"
#include "propose.cuh"

#define RES_MAX_WARPS 32

template <bool WPC>
__device__ __forceinline__ void group_sync() {
  if (WPC) __syncwarp();
  else __syncthreads();
}

// Sum over the chain's group; every thread receives the same value.
template <bool WPC>
__device__ __forceinline__ double group_reduce(double v, double* red, int& parity, int warp, int lane, int nwarps) {
  v = warp_sum(v);
  if (WPC) return v;
  double* r = red + parity * RES_MAX_WARPS;
  parity ^= 1;
  if (lane == 0) r[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < nwarps; w++) s += r[w];
  return s;
}

template <bool WPC, int KC>
__global__ void __launch_bounds__(256)
mh_resident_kernel(ModelParams mp, KParams kp, StreamParams sp, RunBuffers rb, const double* initial,
                   int data_in_smem, int chain_smem_doubles, int mat_doubles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sm = reinterpret_cast<double*>(smem_raw + 128);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int k = kp.k;

  // ---- stage the data once (TMA bulk copy, completion on an mbarrier) --------
  const double* X = mp.X;
  const double* y = mp.y;
  const int* grp = mp.group;
  double* sm_after = sm;
  if (data_in_smem) {
    double* sX = sm;
    double* sy = sX + (size_t)mp.p_x * mp.ld;
    int* sg = reinterpret_cast<int*>(sy + mp.ld);
    const uint32_t bx = (uint32_t)((size_t)mp.p_x * mp.ld * 8), by = (uint32_t)(mp.ld * 8);
    const uint32_t bg = mp.group ? (uint32_t)(mp.ld * 4) : 0u;  // ld is even => multiples of 8; pad below
    const uint32_t bg16 = bg & ~15u;
    if (tid == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(bar, bx + by + bg16);
      for (uint32_t off = 0; off < bx; off += 32768u)
        bulk_g2s(reinterpret_cast<char*>(sX) + off, reinterpret_cast<const char*>(mp.X) + off,
                 min(32768u, bx - off), bar);
      bulk_g2s(sy, mp.y, by, bar);
      if (bg16) bulk_g2s(sg, mp.group, bg16, bar);
    }
    if (bg != bg16)
      for (uint32_t e = bg16 / 4 + tid; e < bg / 4; e += blockDim.x) sg[e] = mp.group[e];
    mbar_wait(bar, 0);
    __syncthreads();
    X = sX; y = sy;
    if (mp.group) grp = sg;
    sm_after = reinterpret_cast<double*>(reinterpret_cast<char*>(sg) + (((size_t)bg + 15) & ~(size_t)15));
  }
  ModelParams lmp = mp;  // same shapes; data pointers passed explicitly below

  double* red = sm_after;  // 2 * RES_MAX_WARPS doubles
  double* chain_base = red + 2 * RES_MAX_WARPS;
  const int chains_per_block = WPC ? nwarps : 1;
  const long long c = (long long)blockIdx.x * chains_per_block + (WPC ? warp : 0);
  if (c >= rb.nchains) return;  // WPC: no block-level sync below this point
  double* th0 = chain_base + (size_t)(WPC ? warp : 0) * chain_smem_doubles;
  double* th1 = th0 + k;
  double* th1u = th1 + k;
  double* scr = th1u + k;  // 4k doubles
  __shared__ int s_flag[RES_MAX_WARPS];
  __shared__ int s_flag_b[RES_MAX_WARPS];  // status of kernel_ram's phase B: a separate word, so that the accept
                                           // step's write to gflag cannot race with slower warps still reading it
  int* gflag = &s_flag[WPC ? warp : 0];
  int* gflag_b = &s_flag_b[WPC ? warp : 0];

  const int gtid = WPC ? lane : tid;
  const int gsize = WPC ? FM_WARP : (int)blockDim.x;
  const bool leader_warp = WPC ? true : (warp == 0);
  int parity = 0;

  auto loglik = [&](const double* th) -> double {
    double part = family_partial(lmp, X, y, grp, th, gtid, gsize);
    double tot = group_reduce<WPC>(part, red, parity, warp, lane, nwarps);
    return family_finish(lmp, th, tot);
  };

  // ---- row 1: R/mcmc.R:737-743 ------------------------------------------------
  if (leader_warp) {
    const double* src = initial ? initial + (size_t)c * k : rb.cur_theta + (size_t)c * k;
    for (int j = lane; j < k; j += FM_WARP) { th0[j] = src[j]; th1[j] = src[j]; th1u[j] = src[j]; }
  }
  group_sync<WPC>();
  // continuing from the device-resident state: f(theta0) is already known (same value the
  // reference recomputes at R/mcmc.R:742), so the first likelihood pass is skipped
  double f0 = initial ? loglik(th0) : rb.cur_f[c];
  unsigned long long n_acc = 0;
  if (leader_warp) {
    const size_t off = (size_t)c;
    for (int j = lane; j < k; j += FM_WARP) { rb.ans[off * k + j] = th0[j]; rb.draws[off * k + j] = th0[j]; }
    for (int a = lane; a < kp.kf; a += FM_WARP) { rb.colsum[((size_t)c * kp.kf + a) * 2] = th0[kp.free_idx[a]]; rb.colsum[((size_t)c * kp.kf + a) * 2 + 1] = 0.0; }
    if (lane == 0) {
      rb.logpost[off] = f0;
      rb.istate[c * FMCMC_ISTATE_LEN + 3] = 0;
      rb.chain_flags[c] = 0;
    }
  }
  group_sync<WPC>();

  ChainCtx cx;
  cx.c = c; cx.theta0 = th0; cx.theta1 = th1; cx.theta1u = th1u; cx.scr = scr;
  cx.ans = rb.ans; cx.ans_stride = (long long)rb.nchains * k;
  cx.mat = mat_doubles ? scr + 4 * k : nullptr;  // chain_smem_doubles = 7k + mat_doubles

  // ---- rows 2..T: R/mcmc.R:749-783 ----------------------------------------------
  for (long long i = 2; i <= rb.T; i++) {
    cx.i = i; cx.f0 = f0;
    if (leader_warp) {
      int rc = propose_warp<KC>(kp, sp, rb, cx, lane);
      if (lane == 0) {
        int fl = 0;
        if (rc) { set_error(rb.err, rc, c + 1, i); fl = 4; }
        else if (KC == KC_RAM && (rb.chain_flags[c] & 1)) {
          bool same = true;
          for (int j = 0; j < k; j++) same &= (th1[j] == th1u[j]);
          fl = same ? 1 : 2;  // 1: adapt with f1u = f1, 2: adapt, evaluate f(theta1u)
        }
        *gflag = fl;
      }
    }
    group_sync<WPC>();
    const int fl = *gflag;
    if (fl == 4) break;
    const double f1 = loglik(th1);
    if (KC == KC_RAM && fl) {  // kernel_ram phase B, R/kernel_ram.R:129-150
      const double f1u = (fl == 1) ? f1 : loglik(th1u);
      int rc = 0;
      if (leader_warp) rc = ram_adapt_warp(kp, rb, cx, f1u, lane);
      if (leader_warp && lane == 0 && rc) set_error(rb.err, rc, c + 1, i);
      if (rc) { /* leader warp only; others learn via failed flag below */ }
      if (leader_warp && lane == 0) *gflag_b = rc ? 4 : 0;
      group_sync<WPC>();
      if (*gflag_b == 4) break;
    }
    if (leader_warp) {
      bool failed = false;
      f0 = accept_row_warp(kp, sp, rb, c, i, th0, th1, f0, f1, lane, n_acc, failed);
      if (lane == 0) {
        scr[0] = f0;
        *gflag = failed ? 4 : 0;
      }
    }
    group_sync<WPC>();
    f0 = scr[0];
    if (*gflag == 4) break;
    group_sync<WPC>();  // scr[0] / gflag are rewritten by the next row's proposal
  }

  // ---- carry the state to the next bulk (R/mcmc.R:909-911) ----------------------
  if (leader_warp) {
    __syncwarp();
    for (int j = lane; j < k; j += FM_WARP) {
      rb.cur_theta[(size_t)c * k + j] = th0[j];
      rb.prop[(size_t)c * k + j] = th1[j];
    }
    if (lane == 0) {
      rb.cur_f[c] = f0;
      atomicAdd(rb.n_accept, n_acc);
    }
  }
}

// f(theta) for `count` parameter vectors: one CTA each (R/mcmc.R:742, exported for tests).
__global__ void logpost_kernel(ModelParams mp, const double* theta, double* out, int count) {
  __shared__ double red[2 * RES_MAX_WARPS];
  const int c = blockIdx.x;
  if (c >= count) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  int parity = 0;
  const double* th = theta + (size_t)c * mp.k;
  double part = family_partial(mp, mp.X, mp.y, mp.group, th, tid, blockDim.x);
  double tot = group_reduce<false>(part, red, parity, warp, lane, nwarps);
  if (tid == 0) out[c] = family_finish(mp, th, tot);
}
