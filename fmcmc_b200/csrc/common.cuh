// common.cuh — device-side PODs and small helpers shared by the stepping paths.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/fmcmc_b200.h"

#define FM_LN_SQRT_2PI 0.918938533204672741780329736406
#define FM_WARP 32
#define FM_FULL 0xffffffffu

// Device copy of fmcmc_kernel_spec: every pointer is a DEVICE pointer.
struct KParams {
  int type, k, kf, scheme, order_len, nadapt_len, mvn_method;
  const int* order;        // explicit scheme (1-based)
  const int* seq;          // random scheme, fed: [C][seq_len] (1-based)
  long long seq_len;
  const double *mu, *scale, *min_, *max_, *lb, *ub;
  const unsigned char* fixed;
  const int* free_idx;     // [kf] which(!fixed), 0-based
  long long warmup, freq, bw;
  double until, eps, Sd, arate;
  const long long* nadapt;
  const double* constr;    // k x k col-major or null
  long long dlen;          // doubles of state per chain
};

struct StreamParams {
  int mode, kdraw;
  unsigned long long seed;
  unsigned int run;
  const double* logu;  // FED [C][T]
  const double* z;     // FED [C][T][kdraw]
};

// Device copy of the model.
struct ModelParams {
  int family;
  unsigned flags;
  long long n;     // observations held by this GPU
  long long n_total;  // observations of the whole model (== n unless sharded over observations across GPUs)
  long long ld;    // leading dimension of X (>= n, even)
  int p_x, n_groups, k;
  int y_binary;    // logistic: every y is exactly +0.0 or 1.0 (fast epilogue)
  const double* X;  // [p_x][ld]
  const double* y;  // [ld]
  const int* group; // [ld]
  double h0, h1;
  const double* Xt;      // tile-major copy of X / y for the DMMA kernel (tiled_mma.cuh), or null
  const unsigned char* Xq;  // tile-major int8 slices of X (+ y) for the tcgen05 kernel (tiled_i8.cuh), or null
  const int* i8_cexp;       // [p_x] column exponents of the slicing
  const double* i8_sxy;     // [p_x] sum_i (y_i - 1/2) x_ij (binary logistic)
  const double* i8_cmax;    // [p_x] max_i |x_ij|, then [1] max_i sum_j x_ij^2 (bounds on |eta| for the un-clamped epilogue)
  const double* sp_tab;  // logistic: (S_k, G_k) softplus table in global memory (softplus.h)
  const double* sp_tab4; // logistic: the 128-per-unit table of the split-integer kernel (softplus.h, FM_SP4_*)
  const double* sp_tab8; // logistic: its 256-per-unit table: (tau, T) of log(2 cosh(a / 2)) (softplus.h, fm_lcosh_table8_fill)
  const double* sp_tab8m; // the same grid with T mean-corrected for the degree-3 core (softplus.h, fm_lcosh_table8m_fill)
  const double* sp_tab6r; // 64-per-unit least-squares cubics, 256 B per point: (c1, c0) x 8 then (c2, c3) x 8 - one copy per shared-memory bank group (fm_lcosh_table6r_fill)
};

// Per-run device buffers shared by both paths.
struct RunBuffers {
  int nchains;
  long long chain_offset;
  long long T;             // rows
  double* ans;             // [T][C][k]
  double* draws;           // [T][C][k]
  double* logpost;         // [T][C]
  double* cur_theta;       // [C][k]  theta0 (state carried between runs)
  double* cur_f;           // [C]     f0
  double* prop;            // [C][k]  theta1 (last proposal; RAM reads fixed coords from it)
  double* prop_u;          // [C][k]  RAM un-reflected proposal
  long long* istate;       // [C][4]
  double* dstate;          // [C][dlen]
  double* colsum;          // [C][kf][2] compensated (hi, lo) running column sums of this run's ans rows (kernel_adapt)
  double* ubuf;            // [C][kf] RAM's U of the current row
  double* work;            // [C][worklen] scratch matrices (L cache, RAM temporaries)
  long long worklen;
  int* chain_flags;        // [C] bit0: RAM adapting this row, bit1: cached Cholesky valid
  int* err;                // [4] code, chain(1-based), row(1-based), spare
  unsigned long long* n_accept;
};

// ---- unfused FP64 arithmetic: the proposal / adaptation code mirrors the R
// expressions operation by operation so fed-stream runs reproduce the oracle's
// samples.  The likelihood kernels use FMA freely.
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }

// ---- Philox4x32-10 (same constants / slot map as oracle/fmcmc_oracle.c) ----
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&o)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}
#define FM_PLAN_RUN 0xFFFFFFFFu
__device__ __forceinline__ void philox_u2(unsigned long long seed, uint32_t chain, uint32_t run,
                                          uint32_t row, uint32_t slot, double& u0, double& u1) {
  uint32_t o[4];
  philox4x32_10(chain, run, row, slot, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  u0 = u53(o[0], o[1]);
  u1 = u53(o[2], o[3]);
}

// Wichura AS 241 (PPND16), the inversion R itself uses for rnorm().
__device__ __forceinline__ double qnorm_dev(double p) {
  double q = p - 0.5, r, val;
  if (fabs(q) <= 0.425) {
    r = .180625 - q * q;
    return q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r +
                    45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r +
                 133.14166789178437745) * r + 3.387132872796366608) /
           (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r +
                21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r +
             42.313330701600911252) * r + 1.);
  }
  r = (q < 0) ? p : 1.0 - p;
  r = sqrt(-log(r));
  if (r <= 5.) {
    r += -1.6;
    val = (((((((r * 7.7454501427834140764e-4 + .0227238449892691845833) * r + .24178072517745061177) * r +
               1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r +
            4.6303378461565452959) * r + 1.42343711074968357734) /
          (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + .0151986665636164571966) * r +
               .14810397642748007459) * r + .68976733498510000455) * r + 1.6763848301838038494) * r +
            2.05319162663775882187) * r + 1.);
  } else {
    r += -5.;
    val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + .0012426609473880784386) * r +
               .026532189526576123093) * r + .29656057182850489123) * r + 1.7848265399172913358) * r +
            5.4637849111641143699) * r + 6.6579046435011037772) /
          (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r +
               7.868691311456132591e-4) * r + .0148753612908506148525) * r + .13692988092273580531) * r +
            .59983220655588793769) * r + 1.);
  }
  return q < 0.0 ? -val : val;
}

// ---- R/kernel.R:450-493 reflect_on_boundaries, one coordinate ----
// R's %% and %/% on positive doubles == exact floored remainder / quotient; fmod()
// is exact, the quotient is recovered by rounding (x - r)/d.
__device__ __forceinline__ double reflect1(double x, double lb, double ub) {
  double d = xsub(ub, lb);
  if (x > ub) {
    double da = xsub(x, ub);
    double r, q;
    if (da <= d && fabs(d) * 2.220446049250313e-16 > 1.0) { r = (da == d) ? 0.0 : da; q = (da == d) ? 1.0 : 0.0; }
    else { r = fmod(da, d); q = rint(xdiv(xsub(da, r), d)); }
    bool odd = fmod(q, 2.0) != 0.0;
    return odd ? xadd(lb, r) : xsub(ub, r);
  } else if (x < lb) {
    double db = xsub(lb, x);
    double r, q;
    if (db <= d && fabs(d) * 2.220446049250313e-16 > 1.0) { r = (db == d) ? 0.0 : db; q = (db == d) ? 1.0 : 0.0; }
    else { r = fmod(db, d); q = rint(xdiv(xsub(db, r), d)); }
    bool odd = fmod(q, 2.0) != 0.0;
    return odd ? xsub(ub, r) : xadd(lb, r);
  }
  return x;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FM_FULL, v, o);
  return v;
}

__device__ __forceinline__ void set_error(int* err, int code, long long chain, long long row) {
  if (atomicCAS(&err[0], 0, code) == 0) {
    err[1] = (int)chain;
    err[2] = (int)row;
  }
}

// ---- programmatic dependent launch (griddepcontrol): the two kernels of an MH row are chained on one stream; with the launch
// attribute cudaLaunchAttributeProgrammaticStreamSerialization the next kernel's CTAs may become resident - and run everything
// that does not read the previous kernel's results: barrier / tensor-memory set-up, the first X stages - while the previous
// kernel is still running.  pdl_wait() returns when the previous kernel has completed and its writes are visible; without
// the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// L2 eviction policies for the bulk copies (createpolicy): lines fetched with evict_last stay in L2 while evict_first lines
// are around to be replaced - a data set larger than L2 that is streamed once per kernel launch keeps a chosen part resident
// from launch to launch instead of none of it (plain LRU evicts everything every pass).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16 B aligned.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
