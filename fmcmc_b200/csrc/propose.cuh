// propose.cuh — warp-cooperative transition kernels (one warp per chain).
//
// Device restatement of the reference's proposal closures:
//   kernel_normal / _reflective   R/kernel_normal.R:65-72, 149-164
//   kernel_unif / _reflective     R/kernel_unif.R:53-57, 124-135
//   kernel_adapt                  R/kernel_adapt.R:84-182  (+ R/recursive.R:63-139)
//   kernel_ram                    R/kernel_ram.R:90-160
//   kernel_nmirror / _umirror     R/kernel_mirror.R:74-157, 198-284
//   plan_update_sequence          R/kernel.R:66-133
//   reflect_on_boundaries         R/kernel.R:450-493
// All arithmetic that produces sample values is unfused (fadd/fmul/...), in the
// order of the R expressions, so that fed-stream runs reproduce the CPU oracle.
#pragma once
#include "common.cuh"

// Kernel classes: the stepping kernels are instantiated once per class so that each instance only carries
// the code of its own transition kernel (the all-in-one build stalled on instruction fetch: ncu
// `no_instruction`, profiles/r01_v2_cfg4_*).
enum { KC_PLAIN = 0 /* normal / unif (+ reflective) */, KC_ADAPT = 1, KC_RAM = 2, KC_MIRROR = 3 };
__host__ __device__ inline int kernel_class(int type) {
  switch (type) {
    case FMCMC_KERNEL_ADAPT: return KC_ADAPT;
    case FMCMC_KERNEL_RAM: return KC_RAM;
    case FMCMC_KERNEL_NMIRROR:
    case FMCMC_KERNEL_UMIRROR: return KC_MIRROR;
  }
  return KC_PLAIN;
}

// Everything a chain's warp needs while proposing row `i` (1-based, R's i).
struct ChainCtx {
  long long c;            // local chain
  long long i;            // row being proposed
  const double* theta0;   // [k]
  double* theta1;         // [k] in: previous proposal, out: proposal (reflected)
  double* theta1u;        // [k] out: RAM un-reflected proposal
  double f0;
  double* scr;            // shared-memory scratch, >= 4*k doubles, private to the warp
  const double* ans;      // this run's ans buffer base ([T][C][k]); row r (1-based) of chain c at
  long long ans_stride;   //   ans[(r-1)*ans_stride + c*k + j], ans_stride = C*k
  double* mat;            // shared-memory matrix scratch private to the warp (4 kf^2 doubles) or null: the
                          //   factorisations then run in global memory (rb.work), ~20x the latency per access
};

__device__ __forceinline__ const double* ans_row(const ChainCtx& cx, const KParams& kp, long long r) {
  return cx.ans + (r - 1) * cx.ans_stride + cx.c * kp.k;
}

// ---- draws ------------------------------------------------------------------
__device__ __forceinline__ double draw_u01(const StreamParams& sp, const RunBuffers& rb, const ChainCtx& cx, int j) {
  if (sp.mode == FMCMC_STREAM_FED)
    return sp.z[((size_t)cx.c * rb.T + (size_t)(cx.i - 1)) * sp.kdraw + j];
  double u0, u1;
  philox_u2(sp.seed, (uint32_t)(rb.chain_offset + cx.c), sp.run, (uint32_t)cx.i, 1u + (uint32_t)(j >> 1), u0, u1);
  return (j & 1) ? u1 : u0;
}
__device__ __forceinline__ double draw_z(const StreamParams& sp, const RunBuffers& rb, const ChainCtx& cx, int j) {
  if (sp.mode == FMCMC_STREAM_FED)
    return sp.z[((size_t)cx.c * rb.T + (size_t)(cx.i - 1)) * sp.kdraw + j];
  return qnorm_dev(draw_u01(sp, rb, cx, j));
}
// Student t, df = kf: normal / sqrt(chisq/df), chisq by Marsaglia-Tsang on Philox slots.
__device__ double draw_t(const StreamParams& sp, const RunBuffers& rb, const ChainCtx& cx, int j, double df) {
  if (sp.mode == FMCMC_STREAM_FED)
    return sp.z[((size_t)cx.c * rb.T + (size_t)(cx.i - 1)) * sp.kdraw + j];
  const uint32_t chain = (uint32_t)(rb.chain_offset + cx.c), row = (uint32_t)cx.i;
  const uint32_t base = 1u + (uint32_t)j * 32u;
  double u0, u1, a = 0.5 * df, boost = 1.0;
  philox_u2(sp.seed, chain, sp.run, row, base + 31u, u0, u1);
  const double znum = qnorm_dev(u0);
  if (a < 1.0) { boost = pow(u1, 1.0 / a); a += 1.0; }
  const double d = a - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * d);
  double g = d;
  for (uint32_t at = 0; at < 31u; at++) {
    philox_u2(sp.seed, chain, sp.run, row, base + at, u0, u1);
    double x = qnorm_dev(u0);
    double v = 1.0 + cc * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    g = d * v;
    if (log(u1) < 0.5 * x * x + d - d * v + d * log(v)) break;
  }
  return znum / sqrt(2.0 * g * boost / df);
}

// ---- update scheme: which coordinate(s) move at row i -------------------------
// joint -> returns -1 (all free coordinates); otherwise the single active coordinate.
__device__ __forceinline__ int active_single(const KParams& kp, const StreamParams& sp, const RunBuffers& rb,
                                             const ChainCtx& cx) {
  switch (kp.scheme) {
    case FMCMC_SCHEME_ORDERED: return kp.free_idx[(cx.i - 1) % kp.kf];
    case FMCMC_SCHEME_EXPLICIT: return kp.order[(cx.i - 1) % kp.order_len] - 1;
    case FMCMC_SCHEME_RANDOM: {
      if (kp.seq) return kp.seq[(size_t)cx.c * kp.seq_len + (size_t)(cx.i - 1)] - 1;
      double u0, u1;
      philox_u2(sp.seed, (uint32_t)(rb.chain_offset + cx.c), FM_PLAN_RUN, (uint32_t)cx.i, 0u, u0, u1);
      int pick = (int)(u1 * kp.kf);
      if (pick >= kp.kf) pick = kp.kf - 1;
      return kp.free_idx[pick];
    }
  }
  return -1;
}

// ---- warp-cooperative lower Cholesky, column-major, left-looking; same operation
// order as the oracle's chol_lower().  Returns 0 or (failing pivot + 1), warp-uniform.
__device__ int chol_lower_warp(int k, const double* A, double* L, int lane) {
  for (int e = lane; e < k * k; e += FM_WARP) L[e] = 0.0;
  __syncwarp();
  for (int j = 0; j < k; j++) {
    // every row i >= j (lane-strided) forms v_i = A[i][j] - sum_{p<j} L[i][p] L[j][p] in the oracle's order;
    // row j's value is the pivot s, so the diagonal costs no extra serial pass
    const int owner = j & (FM_WARP - 1);  // lane that holds row j: rows are dealt out as i = j + lane - owner ...
    double vj = 0.0;
    for (int i = j + ((lane - owner) & (FM_WARP - 1)); i < k; i += FM_WARP) {
      double v = A[i + j * k];
      for (int p = 0; p < j; p++) v = xsub(v, xmul(L[i + p * k], L[j + p * k]));
      if (i == j) vj = v;
      else L[i + j * k] = v;  // un-normalised; divided by the pivot below
    }
    const double s = __shfl_sync(FM_FULL, vj, owner);
    if (!(s > 0.0)) return j + 1;  // warp-uniform
    const double ljj = sqrt(s);
    for (int i = j + ((lane - owner) & (FM_WARP - 1)); i < k; i += FM_WARP) {
      if (i == j) L[j + j * k] = ljj;
      else L[i + j * k] = xdiv(L[i + j * k], ljj);
    }
    __syncwarp();
  }
  return 0;
}

// ---- MASS::mvrnorm's factor: F = V diag(sqrt(max(ev, 0))), eigen(Sigma, symmetric = TRUE) with R's ordering and the
// sign convention of oracle/fmcmc_oracle.c jacobi_eigen() (R/kernel_adapt.R:173-178; FMCMC_MVN_EIGEN).  Cyclic-by-row
// Jacobi on the warp: every lane derives the same rotation from the same three elements with the oracle's unfused
// operations, the three length-k updates of a rotation are lane-strided (their elements are independent), so A, V and
// therefore every draw are the oracle's bit for bit.  A, V: kf^2 doubles of scratch each (shared or global), evs:
// 2 kf doubles.  Returns 0 or FMCMC_ENOTPD (mvrnorm's tol = 1e-6), warp-uniform.
__device__ int eigen_factor_warp(int k, const double* Sigma, double* A, double* V, double* F, double* evs, int lane) {
  for (int e = lane; e < k * k; e += FM_WARP) {
    A[e] = Sigma[e];
    V[e] = ((e % k) == (e / k)) ? 1.0 : 0.0;
  }
  __syncwarp();
  for (int sweep = 0; sweep < 64; sweep++) {
    bool rotated = false;
    for (int p = 0; p < k - 1; p++)
      for (int q = p + 1; q < k; q++) {
        const double apq = A[p + q * k], app = A[p + p * k], aqq = A[q + q * k];
        if (apq == 0.0 || fabs(apq) <= xmul(1e-17, sqrt(fabs(xmul(app, aqq))))) continue;  // warp-uniform
        rotated = true;
        const double theta = xdiv(xsub(aqq, app), xmul(2.0, apq));
        const double t = xdiv(theta >= 0 ? 1.0 : -1.0, xadd(fabs(theta), sqrt(xadd(xmul(theta, theta), 1.0))));
        const double c = xdiv(1.0, sqrt(xadd(xmul(t, t), 1.0))), s = xmul(t, c);
        __syncwarp();
        for (int r = lane; r < k; r += FM_WARP) {  // columns p, q
          const double arp = A[r + p * k], arq = A[r + q * k];
          A[r + p * k] = xsub(xmul(c, arp), xmul(s, arq));
          A[r + q * k] = xadd(xmul(s, arp), xmul(c, arq));
        }
        __syncwarp();
        for (int r = lane; r < k; r += FM_WARP) {  // rows p, q; eigenvector columns p, q
          const double apr = A[p + r * k], aqr = A[q + r * k];
          A[p + r * k] = xsub(xmul(c, apr), xmul(s, aqr));
          A[q + r * k] = xadd(xmul(s, apr), xmul(c, aqr));
          const double vrp = V[r + p * k], vrq = V[r + q * k];
          V[r + p * k] = xsub(xmul(c, vrp), xmul(s, vrq));
          V[r + q * k] = xadd(xmul(s, vrp), xmul(c, vrq));
        }
        __syncwarp();
      }
    if (!rotated) break;
  }
  int* idx = reinterpret_cast<int*>(evs + k);
  if (lane == 0) {  // decreasing eigenvalues; among equal ones the higher original index first (R reverses LAPACK's order)
    for (int i = 0; i < k; i++) { evs[i] = A[i + i * k]; idx[i] = i; }
    for (int i = 0; i < k - 1; i++) {
      int m = i;
      for (int j = i + 1; j < k; j++)
        if (evs[j] > evs[m] || (evs[j] == evs[m] && idx[j] > idx[m])) m = j;
      if (m != i) {
        const double tv = evs[i]; evs[i] = evs[m]; evs[m] = tv;
        const int ti = idx[i]; idx[i] = idx[m]; idx[m] = ti;
      }
    }
  }
  __syncwarp();
  const double ev0 = evs[0];
  bool bad = false;
  for (int j = lane; j < k; j += FM_WARP) {  // one eigenvector per lane: sign convention, scaling
    const double ev = evs[j];
    bad |= ev < xmul(-1e-6, fabs(ev0));
    const double* v = V + (size_t)idx[j] * k;
    int m = 0;
    for (int r = 1; r < k; r++)
      if (fabs(v[r]) > fabs(v[m])) m = r;
    const bool neg = v[m] < 0.0;
    const double sd = sqrt(ev > 0.0 ? ev : 0.0);
    for (int r = 0; r < k; r++) F[r + j * k] = xmul(neg ? -v[r] : v[r], sd);
  }
  bad = __any_sync(FM_FULL, bad);
  __syncwarp();
  return bad ? FMCMC_ENOTPD : 0;
}

// RAM phase B (after the likelihood of the un-reflected proposal is known):
// Sigma <- t(chol(Sigma (I + eta (a_n - arate) UU'/|U|^2) Sigma'))   R/kernel_ram.R:132-150
__device__ int ram_adapt_warp(const KParams& kp, const RunBuffers& rb, const ChainCtx& cx, double f1u, int lane) {
  const int kf = kp.kf;
  double* S = rb.dstate + (size_t)cx.c * kp.dlen;
  const double* U = rb.ubuf + (size_t)cx.c * kf;
  double* Mid = cx.mat ? cx.mat : rb.work + (size_t)cx.c * rb.worklen;  // 4 * kf*kf doubles of scratch
  double* T1 = Mid + kf * kf;
  double* Mx = T1 + kf * kf;
  double* L = Mx + kf * kf;
  double a_n = exp(f1u - cx.f0);
  if (a_n > 1.0) a_n = 1.0;
  if (!isfinite(a_n)) a_n = 0.0;
  double eta = pow((double)cx.i, -2.0 / 3.0) * kf;
  if (eta > 1.0) eta = 1.0;
  double nrm2 = 0.0;
  for (int a = 0; a < kf; a++) nrm2 = xadd(nrm2, xmul(U[a], U[a]));
  double nn = sqrt(nrm2);
  nn = xmul(nn, nn);
  const double cfac = xmul(eta, xsub(a_n, kp.arate));
  for (int e = lane; e < kf * kf; e += FM_WARP) {
    int a = e % kf, b = e / kf;
    Mid[e] = xadd(a == b ? 1.0 : 0.0, xdiv(xmul(cfac, xmul(U[a], U[b])), nn));
  }
  __syncwarp();
  for (int e = lane; e < kf * kf; e += FM_WARP) {
    int a = e % kf, b = e / kf;
    double s = 0.0;
    for (int c = 0; c < kf; c++) s = xadd(s, xmul(S[a + c * kf], Mid[c + b * kf]));
    T1[e] = s;
  }
  __syncwarp();
  for (int e = lane; e < kf * kf; e += FM_WARP) {
    int a = e % kf, b = e / kf;
    double s = 0.0;
    for (int c = 0; c < kf; c++) s = xadd(s, xmul(T1[a + c * kf], S[b + c * kf]));
    Mx[e] = s;
  }
  __syncwarp();
  if (chol_lower_warp(kf, Mx, L, lane)) {
    // Matrix::nearPD is third-party: repair by symmetrising + diagonal jitter (same as the oracle).
    if (lane == 0) rb.istate[cx.c * FMCMC_ISTATE_LEN + 2] += 1;
    double md = 0.0;
    for (int a = 0; a < kf; a++) md += fabs(Mx[a + a * kf]);
    md = md / kf;
    if (!(md > 0.0) || !isfinite(md)) md = 1.0;
    __syncwarp();
    for (int e = lane; e < kf * kf; e += FM_WARP) {
      int a = e % kf, b = e / kf;
      if (a > b) {
        double v = 0.5 * (Mx[a + b * kf] + Mx[b + a * kf]);
        Mx[a + b * kf] = v;
        Mx[b + a * kf] = v;
      }
    }
    __syncwarp();
    double jit = 1e-8 * md;
    int ok = 0;
    for (int tr = 0; tr < 20 && !ok; tr++, jit *= 10.0) {
      for (int e = lane; e < kf * kf; e += FM_WARP) T1[e] = Mx[e] + ((e % kf) == (e / kf) ? jit : 0.0);
      __syncwarp();
      ok = !chol_lower_warp(kf, T1, L, lane);
    }
    if (!ok) return FMCMC_ENOTPD;
  }
  __syncwarp();
  for (int e = lane; e < kf * kf; e += FM_WARP) {
    double v = L[e];
    if (kp.constr) {
      int a = e % kf, b = e / kf;
      v = xmul(v, kp.constr[kp.free_idx[a] + (size_t)kp.free_idx[b] * kp.k]);
    }
    S[e] = v;
  }
  __syncwarp();
  return 0;
}

// Proposal for row cx.i.  Executed by one full warp.  Returns 0 or an FMCMC_E* code
// (warp-uniform).  For kernel_ram this is phase A only: theta1u is the un-reflected
// proposal, theta1 its reflection, and chain_flags bit0 says whether phase B
// (ram_adapt_warp, needs f(theta1u)) must run before the accept step.
template <int KC>
__device__ int propose_warp(const KParams& kp, const StreamParams& sp, const RunBuffers& rb, ChainCtx& cx,
                            int lane) {
  const int k = kp.k, kf = kp.kf;
  long long* ist = rb.istate + cx.c * FMCMC_ISTATE_LEN;
  long long abs_iter = ist[0];
  long long flags = ist[1];
  const long long i = cx.i;
  double* th1 = cx.theta1;
  const double* th0 = cx.theta0;

  {
    if (KC == KC_PLAIN) {
      const bool unif = kp.type == FMCMC_KERNEL_UNIF || kp.type == FMCMC_KERNEL_UNIF_REFLECTIVE;
      const bool refl = kp.type == FMCMC_KERNEL_NORMAL_REFLECTIVE || kp.type == FMCMC_KERNEL_UNIF_REFLECTIVE;
      for (int j = lane; j < k; j += FM_WARP) th1[j] = th0[j];
      __syncwarp();
      const int single = active_single(kp, sp, rb, cx);
      const int na = single < 0 ? kf : 1;
      for (int a = lane; a < na; a += FM_WARP) {
        const int w = single < 0 ? kp.free_idx[a] : single;
        double inc;
        if (unif) {
          const double lo = kp.min_[w], hi = kp.max_[w];
          inc = (lo == hi) ? lo : xadd(lo, xmul(xsub(hi, lo), draw_u01(sp, rb, cx, a)));
        } else {
          const double m = kp.mu[w], s = kp.scale[w];
          inc = (s == 0.0) ? m : xadd(m, xmul(s, draw_z(sp, rb, cx, a)));
        }
        double v = xadd(th0[w], inc);
        if (refl) v = reflect1(v, kp.lb[w], kp.ub[w]);
        th1[w] = v;
      }
      __syncwarp();
      return 0;
    }

    if (KC == KC_ADAPT) {
      double* Sigma = rb.dstate + (size_t)cx.c * kp.dlen;
      double* Mean_prev = Sigma + (size_t)kf * kf;
      double* L = rb.work + (size_t)cx.c * rb.worklen;  // cached Cholesky factor
      int* cflag = rb.chain_flags + cx.c;
      bool dirty = !(*cflag & 2);
      if (!(flags & FMCMC_STATE_INIT)) {  // R/kernel_adapt.R:87-115
        for (int e = lane; e < kf * kf; e += FM_WARP) Sigma[e] = ((e % kf) == (e / kf)) ? kp.eps : 0.0;
        flags |= FMCMC_STATE_INIT;
        dirty = true;
        __syncwarp();
      }
      if (kp.until > (double)abs_iter && abs_iter > kp.warmup && i > 2 && (i % kp.freq) == 0) {  // :118
        double* x = cx.scr;
        double* m = cx.scr + kf;
        double* mp = cx.scr + 2 * kf;
        if (kp.bw > 0) {  // :119-125 windowed: Sigma <- Sd * (cov(ans[(i-bw+1):(i-1), which.]) + Ik)
          const long long r0 = i - kp.bw + 1, r1 = i - 1;
          if (r0 < 1) return FMCMC_EUNSUP;  // the reference indexes row <= 0 here
          const double nr = (double)(r1 - r0 + 1);
          const double Sd = kp.Sd > 0.0 ? kp.Sd : 5.76 / kf;
          for (int a = lane; a < kf; a += FM_WARP) {  // column means of the window (Neumaier-compensated)
            double hi = 0.0, lo = 0.0;
            for (long long r = r0; r <= r1; r++) {
              const double v = ans_row(cx, kp, r)[kp.free_idx[a]];
              const double sum = xadd(hi, v), bp = xsub(sum, hi);
              lo = xadd(lo, xadd(xsub(hi, xsub(sum, bp)), xsub(v, bp)));
              hi = sum;
            }
            m[a] = xdiv(xadd(hi, lo), nr);
          }
          __syncwarp();
          for (int e = lane; e < kf * kf; e += FM_WARP) {
            const int a = e % kf, b = e / kf;
            const int ja = kp.free_idx[a], jb = kp.free_idx[b];
            const double ma = m[a], mb = m[b];
            double sacc = 0.0;
            for (long long r = r0; r <= r1; r++) {
              const double* xr = ans_row(cx, kp, r);
              sacc = fma(xr[ja] - ma, xr[jb] - mb, sacc);
            }
            Sigma[e] = xmul(Sd, xadd(xdiv(sacc, nr - 1.0), a == b ? kp.eps : 0.0));
          }
          __syncwarp();
        } else {
        if (!(flags & FMCMC_STATE_HAS_MEAN)) {  // :130-131
          // colMeans(): R accumulates in long double; the running column sum is kept as a
          // compensated (hi, lo) pair so the mean is the correctly rounded one as well
          const double* cs = rb.colsum + (size_t)cx.c * 2 * kf;
          const double nn = (double)(i - 1);
          for (int a = lane; a < kf; a += FM_WARP) {
            const double hi = cs[2 * a], lo = cs[2 * a + 1];
            const double q = xdiv(hi, nn);
            const double r = fma(-q, nn, hi);
            Mean_prev[a] = xadd(q, xdiv(xadd(r, lo), nn));
          }
          flags |= FMCMC_STATE_HAS_MEAN;
          __syncwarp();
        }
        if (i - kp.freq < 1) return FMCMC_EUNSUP;
        const double t = (double)(abs_iter - kp.freq);  // :144
        if (t == 0.0) return FMCMC_EUNSUP;
        for (long long jj = 0; jj < kp.freq; jj++) {  // rows (i-freq):(i-1), R/recursive.R:78-110
          const double* xr = ans_row(cx, kp, i - kp.freq + jj);
          const double tj = t + (double)jj;
          for (int a = lane; a < kf; a += FM_WARP) {
            const double xa = xr[kp.free_idx[a]];
            const double mpa = Mean_prev[a];
            x[a] = xa;
            mp[a] = mpa;
            m[a] = xdiv(xadd(xmul(mpa, tj), xa), tj + 1.0);  // mean_recursive :126
          }
          __syncwarp();
          const double c1 = xdiv(tj - 1.0, tj), c2 = xdiv(1.0, tj);
          for (int e = lane; e < kf * kf; e += FM_WARP) {  // cov_recursive :112-118, Sd = 1, eps = 1e-5
            const int a = e % kf, b = e / kf;
            double inner = xsub(xmul(tj, xmul(mp[a], mp[b])), xmul(tj + 1.0, xmul(m[a], m[b])));
            inner = xadd(inner, xmul(x[a], x[b]));
            inner = xadd(inner, xmul(1e-5, a == b ? kp.eps : 0.0));
            Sigma[e] = xadd(xmul(c1, Sigma[e]), xmul(c2, inner));
          }
          __syncwarp();
          for (int a = lane; a < kf; a += FM_WARP) Mean_prev[a] = m[a];
          __syncwarp();
        }
        }  // bw <= 0
        dirty = true;
      }
      abs_iter += 1;  // :170
      const double* Lr = L;  // factor the matvec below reads
      const bool eigen = kp.mvn_method == FMCMC_MVN_EIGEN;
      if (dirty) {
        if (eigen) {  // MASS::mvrnorm's own factor (verification mode): A, V in shared memory or behind L in rb.work
          double* As = cx.mat ? cx.mat : L + (size_t)kf * kf;
          double* Vs = As + (size_t)kf * kf;
          if (eigen_factor_warp(kf, Sigma, As, Vs, L, cx.scr, lane)) return FMCMC_ENOTPD;
        } else if (cx.mat) {  // factorise in shared memory, keep a copy in HBM for the rows that do not re-adapt
          double* As = cx.mat;
          double* Ls = cx.mat + (size_t)kf * kf;
          for (int e = lane; e < kf * kf; e += FM_WARP) As[e] = Sigma[e];
          __syncwarp();
          if (chol_lower_warp(kf, As, Ls, lane)) return FMCMC_ENOTPD;
          for (int e = lane; e < kf * kf; e += FM_WARP) L[e] = Ls[e];
          Lr = Ls;
        } else if (chol_lower_warp(kf, Sigma, L, lane)) {
          return FMCMC_ENOTPD;  // mvrnorm: "'Sigma' is not positive definite"
        }
        if (lane == 0) *cflag |= 2;
        __syncwarp();
      }
      double* z = cx.scr + 3 * kf;
      for (int a = lane; a < kf; a += FM_WARP) z[a] = draw_z(sp, rb, cx, a);
      for (int j = lane; j < k; j += FM_WARP) th1[j] = th0[j];
      __syncwarp();
      for (int a = lane; a < kf; a += FM_WARP) {  // :173-180
        double s = 0.0;
        const int nb = eigen ? kf : a + 1;  // full factor / lower triangle
        for (int b = 0; b < nb; b++) s = xadd(s, xmul(Lr[a + b * kf], z[b]));
        const int w = kp.free_idx[a];
        th1[w] = reflect1(xadd(th0[w], xadd(kp.mu[w], s)), kp.lb[w], kp.ub[w]);
      }
      __syncwarp();
      if (lane == 0) { ist[0] = abs_iter; ist[1] = flags; }
      return 0;
    }

    if (KC == KC_RAM) {
      double* S = rb.dstate + (size_t)cx.c * kp.dlen;
      double* U = rb.ubuf + (size_t)cx.c * kf;
      if (!(flags & FMCMC_STATE_INIT)) {  // R/kernel_ram.R:114-116
        for (int e = lane; e < kf * kf; e += FM_WARP) S[e] = ((e % kf) == (e / kf)) ? kp.eps : 0.0;
        flags |= FMCMC_STATE_INIT;
        __syncwarp();
      }
      for (int a = lane; a < kf; a += FM_WARP) U[a] = draw_t(sp, rb, cx, a, (double)kf);  // :124
      // :125 theta1 <- env$theta1 (previous proposal): th1 already holds it; copy to theta1u
      for (int j = lane; j < k; j += FM_WARP) cx.theta1u[j] = th1[j];
      __syncwarp();
      for (int a = lane; a < kf; a += FM_WARP) {  // :126
        double s = 0.0;
        for (int b = 0; b < kf; b++) s = xadd(s, xmul(S[a + b * kf], U[b]));
        const int w = kp.free_idx[a];
        const double v = xadd(th0[w], s);
        cx.theta1u[w] = v;
        th1[w] = reflect1(v, kp.lb[w], kp.ub[w]);  // :158 (applied after phase B in R; same value)
      }
      const bool adapting = kp.until > (double)abs_iter && abs_iter > kp.warmup && (i % kp.freq) == 0;  // :129
      __syncwarp();
      if (lane == 0) {
        rb.chain_flags[cx.c] = (rb.chain_flags[cx.c] & ~1) | (adapting ? 1 : 0);
        ist[0] = abs_iter + 1;  // :155
        ist[1] = flags;
      }
      return 0;
    }

    if (KC == KC_MIRROR) {
      double* mu = rb.dstate + (size_t)cx.c * kp.dlen;
      double* scale = mu + k;
      double* obs = mu + 2 * k;
      if (!(flags & FMCMC_STATE_INIT)) {
        for (int j = lane; j < k; j += FM_WARP) { mu[j] = kp.mu[j]; scale[j] = kp.scale[j]; obs[j] = 0.0; }
        flags |= FMCMC_STATE_INIT;
        __syncwarp();
      }
      const long long ai = abs_iter;
      const long long nad0 = kp.nadapt_len > 0 ? kp.nadapt[0] : -1;
      if (ai >= 1 && ai <= kp.warmup) {  // R/kernel_mirror.R:104-112
        const double* x = ans_row(cx, kp, i - 1);
        for (int j = lane; j < k; j += FM_WARP)
          mu[j] = xdiv(xadd(xmul(mu[j], (double)ai), x[j]), (double)ai + 1.0);
      }
      if (ai == nad0) {  // :115-119
        if (i - 1 < 2) return FMCMC_EUNSUP;  // quirk D8
        const long long same = (i - 2) - ist[3];
        const double v = xsub(1.0, xdiv((double)same, (double)(i - 2)));
        for (int j = lane; j < k; j += FM_WARP) obs[j] = v;
        flags = (flags & ~(3LL << FMCMC_STATE_OBS_SHIFT)) | (1LL << FMCMC_STATE_OBS_SHIFT);
      } else if (nad0 >= 0 && ai > nad0 && ai <= kp.warmup) {  // :120-128
        if (i - 2 < 1) return FMCMC_EUNSUP;  // quirk D8
        const double* x1 = ans_row(cx, kp, i - 1);
        const double* x2 = ans_row(cx, kp, i - 2);
        for (int j = lane; j < k; j += FM_WARP) {
          const double ind = (x1[j] != x2[j]) ? 1.0 : 0.0;
          obs[j] = xdiv(xadd(xmul(obs[j], (double)ai), ind), (double)ai + 1.0);
        }
        flags = (flags & ~(3LL << FMCMC_STATE_OBS_SHIFT)) | (2LL << FMCMC_STATE_OBS_SHIFT);
      }
      bool in_nadapt = false;
      for (int q = 0; q < kp.nadapt_len; q++) in_nadapt |= (kp.nadapt[q] == ai);
      __syncwarp();
      if (in_nadapt) {  // :131-137
        const double den = tan(xmul(xdiv(M_PI, 2.0), kp.arate));
        for (int j = lane; j < k; j += FM_WARP)
          scale[j] = xdiv(xmul(scale[j], tan(xmul(xdiv(M_PI, 2.0), obs[j]))), den);
      }
      for (int j = lane; j < k; j += FM_WARP) th1[j] = th0[j];
      __syncwarp();
      const int single = active_single(kp, sp, rb, cx);
      const int na = single < 0 ? kf : 1;
      const double sqrt3 = sqrt(3.0);
      for (int a = lane; a < na; a += FM_WARP) {
        const int w = single < 0 ? kp.free_idx[a] : single;
        const double centre = xsub(xmul(2.0, mu[w]), th0[w]);
        double v;
        if (kp.type == FMCMC_KERNEL_NMIRROR) {  // :146-150
          v = (scale[w] == 0.0) ? centre : xadd(centre, xmul(scale[w], draw_z(sp, rb, cx, a)));
        } else {  // :273-277
          const double lo = xsub(centre, xmul(sqrt3, scale[w])), hi = xadd(centre, xmul(sqrt3, scale[w]));
          v = (lo == hi) ? lo : xadd(lo, xmul(xsub(hi, lo), draw_u01(sp, rb, cx, a)));
        }
        th1[w] = reflect1(v, kp.lb[w], kp.ub[w]);
      }
      __syncwarp();
      if (lane == 0) { ist[0] = abs_iter + 1; ist[1] = flags; }
      return 0;
    }
  }
  return FMCMC_EINVAL;
}

// Accept / reject + bookkeeping for row i, executed by ALL 32 lanes of the chain's warp (coordinates are
// lane-strided: the chain state may live in HBM, where a serial loop costs one L2 round trip per
// coordinate).  R/mcmc.R:752-778.  th0/th1 are the chain's state vectors (any address space); returns the
// (possibly updated) f0, warp-uniform.  n_acc is only meaningful in lane 0.
__device__ __forceinline__ double accept_row_warp(const KParams& kp, const StreamParams& sp, const RunBuffers& rb,
                                                  long long c, long long i, double* th0, const double* th1,
                                                  double f0, double f1, int lane, unsigned long long& n_acc,
                                                  bool& failed) {
  const int k = kp.k;
  const size_t row_off = ((size_t)(i - 1) * rb.nchains + (size_t)c);
  double* draws = rb.draws + row_off * k;
  double* ans = rb.ans + row_off * k;
  for (int j = lane; j < k; j += FM_WARP) draws[j] = th1[j];
  if (lane == 0) rb.logpost[row_off] = f1;
  if (isnan(f1)) {  // :758-765
    if (lane == 0) set_error(rb.err, FMCMC_ENAN, c + 1, i);
    failed = true;
    return f0;
  }
  const double ratio = f1 - f0;  // R/kernel.R:302-303
  if (isnan(ratio)) {            // quirk D10
    if (lane == 0) set_error(rb.err, FMCMC_ENANRATIO, c + 1, i);
    failed = true;
    return f0;
  }
  double logu;
  if (sp.mode == FMCMC_STREAM_FED) {
    logu = sp.logu[(size_t)c * rb.T + (size_t)(i - 1)];
  } else {
    double u0, u1;
    philox_u2(sp.seed, (uint32_t)(rb.chain_offset + c), sp.run, (uint32_t)i, 0u, u0, u1);
    logu = log(u0);
  }
  if (logu < ratio) {  // :770 (warp-uniform)
    bool changed = false;
    for (int j = lane; j < k; j += FM_WARP) {
      const double v = th1[j];
      changed |= (th0[j] != v);
      th0[j] = v;
    }
    changed = __any_sync(FM_FULL, changed);
    if (lane == 0) {
      if (changed) rb.istate[c * FMCMC_ISTATE_LEN + 3] += 1;
      n_acc += 1;
    }
    f0 = f1;
  }
  __syncwarp();
  for (int j = lane; j < k; j += FM_WARP) ans[j] = th0[j];
  double* cs = rb.colsum + (size_t)c * 2 * kp.kf;
  for (int a = lane; a < kp.kf; a += FM_WARP) {  // Neumaier two-sum: (hi, lo) += theta0
    const double x = th0[kp.free_idx[a]], hi = cs[2 * a];
    const double sum = xadd(hi, x);
    const double bp = xsub(sum, hi);
    const double e = xadd(xsub(hi, xsub(sum, bp)), xsub(x, bp));
    cs[2 * a] = sum;
    cs[2 * a + 1] = xadd(cs[2 * a + 1], e);
  }
  __syncwarp();
  return f0;
}
