// tiled_i8.cuh — observation-tiled likelihood kernel, split-integer tensor-core variant (path 4).
//
// Same contract and grid as tiled_loglik_mma_kernel (tiled_mma.cuh): a CTA owns one observation slice and
// one block of 128 chains and writes partial[slice][chain].  What changes is where the X.Theta contraction runs.
// The FP64 pipe is the binding roof of paths 2 / 3, and 32 of its 51 instruction slots per
// evaluation are the dot product.  tcgen05 has no FP64 kind, but B200's 5th-generation tensor cores multiply
// int8 with exact int32 accumulation (tcgen05.mma kind::i8, SASS UTCIMMA), so the product is computed
// EXACTLY on integer slices of the operands (Ozaki splitting) and only reassembled in FP64:
//
//   column j of X:  x_ij   = 2^cexp_j * sum_s sx_s[i][j] 2^(-6 - 7 s),   sx_s in [-64, 64]          (once per model)
//   chain c:        th_cj 2^cexp_j = 2^eth_c * sum_s sth_s[c][j] 2^(-6 - 7 s)                         (once per launch)
//   eta_ic = 2^(eth_c - 12) * sum_d 2^(-7 d) a_d,     a_d = sum_{s + s' = d} sum_j sx_s[i][j] sth_s'[c][j]
//
// a_d (d = 0 .. NS-1) are int32 accumulators in tensor memory.  Dropping the pairs with s + s' >= NS leaves an
// absolute error of ~2^(-7 NS - 5) max_j|theta_j 2^cexp_j| in eta — relative to the largest single term of the dot
// product, thanks to the column exponents (NS = 6: ~1e-12; the FP64 rounding error of a 32-term dot product is
// ~1e-15).  Summed over n observations with unbiased signs the log-posterior stays ~1e-13 .. 1e-15 relative;
// kernel_ram, whose adaptation consumes f itself, runs on 7 slices (fmcmc_run).  Everything is integer until the
// reassembly, so results are bit-reproducible and independent of how chains are grouped into CTAs / calls / GPUs
// (per-column and per-chain exponents only, no block-wide scaling).
//
// Mapping: TMEM lane = chain (A operand = Theta slices, written once per launch into tensor memory with
// tcgen05.st when they fit, else shared memory), TMEM column = observation (B operand = the X slice tile,
// streamed from HBM by one cp.async.bulk per stage in exactly the core-matrix layout the MMA reads).
// One MMA per (Theta slice j, K block) multiplies the slice with X slices 0 .. NS-1-j AT ONCE (they are consecutive
// row groups of the B operand, N = 32 (NS - j)); the product with X slice s lands 32 s columns to the right, i.e.
// on diagonal s + j: NS instructions per 32-observation block instead of NS (NS + 1) / 2 (a UTCIMMA costs ~80 clk
// to issue whatever its N - profiles/r01_i8_findings.md).
// An epilogue thread owns one chain: it reads the NS accumulators of an observation with tcgen05.ld, merges them
// exactly (int32 pairs -> int64, ONE I2F.F64.S64), scales (1 FP64 instruction), runs the family epilogue and adds
// to a private sum — no cross-lane reduction.  Warp roles: 16 epilogue warps (four per TMEM lane quarter, 8 of the
// 32 columns each), 1 TMA producer, 1 MMA issuer (whole warp in the loop, elect.sync around the issue);
// accumulators are double-buffered in TMEM.  On B200 the FP64 pipe and the tensor pipe share their datapath
// (DFMA drops to 3 % of its rate while UTCIMMAs run), so the kernel's time is tensor time PLUS FP64 time; the
// double buffer hides latencies only.
//
// FP64-pipe slots per evaluation: 12 (binary logistic with the 256-per-unit (tau, T) table of log(2 cosh(eta / 2)): 3 range
// reduction with the scale folded in + 9 Taylor core and accumulation; 16 with the 128-per-unit softplus table used when the
// Theta slices need the shared memory; the linear term theta . X'(y - 1/2) is a per-chain dot product with a vector computed
// once per model) against 51 on path 3; Gaussian: 3 against 2 p_x + 2.
// Role branches test a warp index that went through redux.sync, so ptxas knows they are warp-uniform and keeps TMEM / barrier
// addresses in uniform registers (profiles/r01_i8_findings.md, section 7).
#pragma once
#include "tiled.cuh"

#define I8_CHAINS 128
#define I8_MAX_EPI_WARPS 16
// Epilogue warps in two groups, one per accumulator set: the four warps of a TMEM lane quarter are split 2 + 2, a group
// only visits the blocks of ITS set and each of its warps takes 16 of the block's 32 columns (two chunks of 8) instead of
// 8 columns of every block.  The per-block overhead - barrier wait, fences, release, loop control: ~25 warp instructions
// - is then paid once per 16 evaluations instead of once per 8, in a kernel whose time is its issued instructions.
// (B200, cfg3, per launch: un-grouped 1.663 ms; grouped 1.680; grouped with the two-chunk loop unrolled 1.651 - and at
// K = 128, where the tensor pipe binds, 181.2 against 186.1 ms.)
#ifndef I8_GROUPED
#define I8_GROUPED 1
#endif
#if I8_GROUPED && !defined(I8_CC_UNROLL)
#define I8_CC_UNROLL 1
#endif

// Digit width of the integer slices.  8 (default): balanced base-256 digits in [-128, 127] - the whole int8 range - with the
// leading digit scaled by 2^7: NS = 5 slices carry 39 bits (dropped pairs s + s' >= 5: <= 6 2^-40 = 5.5e-12 of the largest
// term), 15 slice pairs instead of 21, five diagonals to merge instead of six, and at K = 128 every Theta slice fits tensor
// memory next to the two accumulator sets.  NS = 6 (kernel_ram, short data): 7 2^-48 = 2.5e-14.  7 (round 1's scheme, NS = 6 /
// 7, digits in [-64, 64], 1.6e-12 / 1.4e-14): -DI8_DIGIT_BITS=7.
#ifndef I8_DIGIT_BITS
#define I8_DIGIT_BITS 8
#endif
#ifndef I8_TRIPLE
#define I8_TRIPLE 0   // 1: three accumulator sets where they fit (I8Geom::NACC).  Measured on B200 (cfg3): 1.537 ms against 1.517
#endif                // with two - the waiting it removes comes back as math-pipe stalls (profiles/r02_findings.md, section 9)

#define I8_NS_LO (I8_DIGIT_BITS == 8 ? 5 : 6)   // slices of the default accuracy tier; kernel_ram and n < 65 536 use one more
// Observations per MMA block.  32 by default.  At K = 128 (four K blocks) the kernel is bound by the tensor pipe, which needs
// >= 64 clk per M = 128 int8 MMA whatever its N <= 96 (profiles/r01_i8_findings.md, section 5): the triangle N = 160 .. 32 of a
// 32-observation block costs 347 clk per K block against 255 at peak.  With 64-observation blocks (Gaussian family, K = 128,
// five slices) the triangle is N = 160 + 160, 256, 192, 128, 64: 542 clk per 64 observations, 22 % less tensor time.  The five
// diagonals of 64 observations take 320 accumulator columns, so there is ONE accumulator set (320 + 160 columns of Theta
// slices = 480): MMAs and epilogue alternate - which costs nothing here, the two were time-additive already (the tensor pipe
// and the FP64 pipe share their datapath).
#ifndef I8_WIDE_BLOCKS
#define I8_WIDE_BLOCKS 1   // measured on B200 (cfg5): 135.2 ms per launch against 138.7 with 32-observation blocks and two sets
#endif
// Gaussian, K = 128, two accumulator sets only (-DI8_WIDE_BLOCKS=0): the integer half of the epilogue for all of a warp's columns
// first, the FP64 half after, so that the integer half could overlap the other set's MMAs.  Built, parity-green, measured:
// 138.65 ms against 138.7 - no gain (the MMA sequence alone takes 111.5 ms with its issue floors; what the epilogue adds on
// top is not the FP64 / integer interleaving).  Off.
#ifndef I8_PHASE_SPLIT
#define I8_PHASE_SPLIT 0
#endif
// 48-observation blocks (N = 240, 192, 144, 96, 48: 435 clk per K block and 48 observations, two sets of 240 columns, the 32
// columns left hold Theta slices 0 .. 3 of the first K block, the others are read from shared memory): build-time experiments,
// -DI8_BLK_LOGIT_K32=48 / -DI8_BLK_GAUSS_K128=48 (profiles/r02_findings.md, section 19)
#ifndef I8_BLK_LOGIT_K32
#define I8_BLK_LOGIT_K32 32
#endif
#ifndef I8_BLK_GAUSS_K128
#define I8_BLK_GAUSS_K128 (I8_WIDE_BLOCKS != 0 ? 64 : 32)
#endif
template <int NS, int KB>
__host__ __device__ constexpr int i8_blk(int family) {
  if (I8_DIGIT_BITS == 8 && NS == 5) {
    if (family == FMCMC_FAMILY_GAUSSIAN_LM && KB == 4) return I8_BLK_GAUSS_K128;
    if (family == FMCMC_FAMILY_LOGISTIC && KB == 1) return I8_BLK_LOGIT_K32;
  }
  return 32;
}
template <int NS, int KB, int BLK_ = 32>
struct I8Geom {
  static constexpr int DB = I8_DIGIT_BITS;
  static constexpr int BLK = BLK_;                                  // observations per MMA block
  static constexpr int TO = KB == 1 ? (BLK == 48 ? 96 : 128) : (KB == 2 ? 64 : BLK);   // observations per pipeline stage
  static constexpr int NBLK = TO / BLK;
  static constexpr int SLAB_BYTES = BLK * 32;                        // one slice x one K block of a block: BLK / 8 row groups x 2 chunks x 128 B
  static constexpr int BLOCK_BYTES = KB * NS * SLAB_BYTES;           // [kb][slice][group][chunk][8 rows][16 B]
  static constexpr int SLICE_BYTES = NBLK * BLOCK_BYTES;
  static constexpr int STAGE_BYTES = SLICE_BYTES;                   // a stage is pure MMA operand: only the tensor pipe holds it
  static constexpr int ACC_COLS = NS * BLK;                         // TMEM columns of one accumulator set
  // Accumulator sets: two - or, with -DI8_TRIPLE=1, THREE when they fit tensor memory with all but the last Theta slice beside
  // them (K = 32, 5 slices: 3 x 160 + 4 x 8 = 512 columns exactly).  With two sets a group of epilogue warps waits for the MMAs of
  // its next block (12 % of the warps' time in round 2's profile); with three the MMA warp runs a block ahead - correct
  // (all parity tests), but not faster: while the tensor pipe works the FP64 pipe does not, wherever the warps happen to wait.
  // Wide blocks (BLK = 64): one set.
  static constexpr int NACC = BLK > 48 ? 1 : ((I8_TRIPLE != 0 && KB == 1 && 3 * ACC_COLS + (NS - 1) * 8 <= 512) ? 3 : 2);
  // A operand (Theta slices): 8 columns per (slice, K block) unit.  NACC == 2: as many K blocks as fit beside the two sets
  // live in tensor memory (all slices of those K blocks), the rest in shared memory; NACC == 3: slices 0 .. NS-2 in tensor
  // memory, the last slice (one MMA of N = 32 per block) in shared memory.  An MMA whose A comes from shared memory re-reads 4 KB.
  // When not even one K block fits with all its slices (48-observation blocks: 32 columns left), slices 0 .. A_TS-1 of the first K
  // block do.
  static constexpr int A_COL0 = NACC * ACC_COLS;
  static constexpr int A_FULL = (512 - A_COL0) / (NS * 8);          // K blocks that fit with all their slices
  static constexpr int A_TS = NACC == 3 ? NS - 1 : (A_FULL >= 1 ? NS : ((512 - A_COL0) / 8 < NS ? (512 - A_COL0) / 8 : NS));
  static constexpr int A_TKB = NACC == 3 ? 1 : (A_FULL >= 1 ? (A_FULL < KB ? A_FULL : KB) : ((512 - A_COL0) >= 8 ? 1 : 0));
  static constexpr int A_SKB = KB - A_TKB;
  static constexpr int A_SMEM_BYTES = (NS * A_SKB + (NS - A_TS) * A_TKB) * 4096;
  __host__ __device__ static constexpr bool a_in_tmem(int i, int kb) { return kb < A_TKB && i < A_TS; }
  __host__ __device__ static constexpr int a_tmem_col(int i, int kb) { return A_COL0 + (i * A_TKB + kb) * 8; }
  __host__ __device__ static constexpr int a_smem_unit(int i, int kb) {
    return kb < A_TKB ? NS * A_SKB + (i - A_TS) * A_TKB + kb : i * A_SKB + (kb - A_TKB);
  }
  static constexpr int SHIFT = 2 * (DB - 1) + DB * (NS - 1);        // eta = t * 2^(eth - SHIFT)
};

// table of the logistic epilogue (softplus.h): level 2 = 256 entries per unit (160 KB, (tau, T) of log(2 cosh(a / 2)), one degree-4
// Taylor core: 9 FP64 instructions) when it leaves room for two pipeline stages, else level 1 = 128 per unit (80 KB, (S, G) of
// log1p(exp(-a)), two degree-4 polynomials: 13)
template <int NS, int KB, int BLK = 32>
__host__ __device__ constexpr int i8_smem_fixed(int table_bytes) {
  return 256 + I8Geom<NS, KB, BLK>::A_SMEM_BYTES + table_bytes + (I8_MAX_EPI_WARPS / 4) * I8_CHAINS * 8 + 1024;
}
// prologue scratch: per (part, chain) partial maxima / sums of the chain's parameters, 5 doubles each.  The logistic family
// passes them through the table region before the table is loaded; the Gaussian family has no table and gets a region of its own
__host__ __device__ constexpr int i8_scratch_bytes(int family) {
  return family == FMCMC_FAMILY_LOGISTIC ? 0 : (I8_MAX_EPI_WARPS / 4) * I8_CHAINS * 5 * 8;
}
template <int NS, int KB>
__host__ __device__ constexpr int i8_table_level() {
  constexpr int BL = i8_blk<NS, KB>(FMCMC_FAMILY_LOGISTIC);
  return (232448 - i8_smem_fixed<NS, KB, BL>(FM_SP8_ENTRIES * 16)) / I8Geom<NS, KB, BL>::STAGE_BYTES >= 2 ? 2 : 1;
}
template <int NS, int KB>
__host__ __device__ constexpr int i8_table_entries() { return i8_table_level<NS, KB>() == 2 ? FM_SP8_ENTRIES : FM_SP4_ENTRIES; }
// level 3 (binary logistic with the level-2 table only): the bank-group-replicated cubic table (softplus.h,
// fm_lcosh_table6r_fill; 256 B per point) - as many points as fit beside TWO pipeline stages, i.e. |eta| up to ~11; chosen per
// CTA when every chain of the CTA bounds |eta| below the table's end, else the CTA loads a level-2 table into the same region.
#ifndef I8_REP_TABLE
#define I8_REP_TABLE 1
#endif
template <int NS, int KB>
__host__ __device__ constexpr int i8_rep_entries() {
  constexpr int BL = i8_blk<NS, KB>(FMCMC_FAMILY_LOGISTIC);
  const int fit = (232448 - i8_smem_fixed<NS, KB, BL>(0) - 2 * I8Geom<NS, KB, BL>::STAGE_BYTES) / FM_LC6_POINT_BYTES;
  const int e = fit > FM_LC6_ENTRIES_MAX ? FM_LC6_ENTRIES_MAX : fit;
  return (I8_REP_TABLE != 0 && i8_table_level<NS, KB>() == 2 && e >= 5 * FM_LC6_H + 1) ? e : 0;  // worth it from |eta| <= 5 on
}
template <int NS, int KB>
__host__ __device__ constexpr int i8_table_bytes(int family, bool ybin) {
  const int plain = i8_table_entries<NS, KB>() * 16, rep = ybin ? i8_rep_entries<NS, KB>() * FM_LC6_POINT_BYTES : 0;
  return family == FMCMC_FAMILY_LOGISTIC ? (plain > rep ? plain : rep) : 0;
}
// pipeline depth: as many stages as fit beside the Theta slices and the softplus table, at most 6.
template <int NS, int KB>
__host__ __device__ constexpr int i8_stages(int family, bool ybin) {
  constexpr int B = i8_blk<NS, KB>(FMCMC_FAMILY_GAUSSIAN_LM), BL = i8_blk<NS, KB>(FMCMC_FAMILY_LOGISTIC);
  const int stage = family == FMCMC_FAMILY_GAUSSIAN_LM ? I8Geom<NS, KB, B>::STAGE_BYTES : I8Geom<NS, KB, BL>::STAGE_BYTES;
  const int fixed = family == FMCMC_FAMILY_GAUSSIAN_LM ? i8_smem_fixed<NS, KB, B>(0) : i8_smem_fixed<NS, KB, BL>(i8_table_bytes<NS, KB>(family, ybin));
  const int fit = (232448 - fixed - i8_scratch_bytes(family)) / stage;
  return fit > 6 ? 6 : fit;
}
template <int NS, int KB>
__host__ __device__ inline size_t tiled_i8_smem_bytes(int family, bool ybin) {
  constexpr int B = i8_blk<NS, KB>(FMCMC_FAMILY_GAUSSIAN_LM), BL = i8_blk<NS, KB>(FMCMC_FAMILY_LOGISTIC);
  const size_t stage = family == FMCMC_FAMILY_GAUSSIAN_LM ? I8Geom<NS, KB, B>::STAGE_BYTES : I8Geom<NS, KB, BL>::STAGE_BYTES;
  const size_t asmem = family == FMCMC_FAMILY_GAUSSIAN_LM ? I8Geom<NS, KB, B>::A_SMEM_BYTES : I8Geom<NS, KB, BL>::A_SMEM_BYTES;
  size_t b = 256 + (size_t)i8_stages<NS, KB>(family, ybin) * stage + asmem +
             (size_t)i8_table_bytes<NS, KB>(family, ybin) + (I8_MAX_EPI_WARPS / 4) * I8_CHAINS * sizeof(double) + i8_scratch_bytes(family);
  return b < 120 * 1024 ? 120 * 1024 : b;  // one CTA per SM: a CTA allocates all 512 TMEM columns
}

// ---- slicing -----------------------------------------------------------------------------------------
#define I8_EMAX 480
// smallest exponent e (clamped) with m < 2^e - 8-bit digits: with m 2^-e < 127 / 128, so that the leading digit, carry
// included, stays <= 127
__device__ __forceinline__ int i8_exponent(double m) {
  if (!(m > 0.0)) return 0;
  int e = ilogb(m) + 1;
#if I8_DIGIT_BITS == 8
  if (!(scalbn(m, -e) < 127.0 / 128.0)) e += 1;
#endif
  return max(-I8_EMAX, min(I8_EMAX, e));
}
// |u| <= 1 (8-bit digits: |u| < 127 / 128)  ->  u = sum_s s[s] 2^(-(DB - 1) - DB s) + O(2^(-DB NS + 1)), every step exact in FP64
template <int NS>
__device__ __forceinline__ void i8_slices(double u, int (&s)[NS]) {
#if I8_DIGIT_BITS == 8
  double r = u * 128.0;
#pragma unroll
  for (int i = 0; i < NS; i++) {
    const double q = rint(r);          // in [-128, 128]: a remainder of +-1/2 becomes +-128
    s[i] = (int)q;
    r = (r - q) * 256.0;
  }
#pragma unroll
  for (int i = NS - 1; i >= 1; i--)    // balanced digits in [-128, 127]: +128 (+129 with a carry) is -128 (-127) and one more above
    if (s[i] >= 128) { s[i] -= 256; s[i - 1] += 1; }
#else
  double r = u * 64.0;
#pragma unroll
  for (int i = 0; i < NS; i++) {
    const double q = rint(r);
    s[i] = (int)q;
    r = (r - q) * 128.0;
  }
#endif
}

// Tile-major int8 copy of X (+ y and the row exponents), built once per model: tile t of TO observations is
// stored exactly as a pipeline stage sits in shared memory:
//   [block b = obs / 32][k block kb][slice s][group g = (obs % 32) / 8][chunk c = (k % 32) / 16][row = obs % 8][16 bytes k % 16]
//   then meta[TO]
// i.e. 128-byte core matrices (8 rows x 16 B) in the canonical K-major no-swizzle operand layout of tcgen05.mma
// (leading byte offset 128 between the two K chunks, stride byte offset 256 between row groups).  Inside a block and
// a K block the slices are CONSECUTIVE row groups, so slices 0 .. NS-1-j of a block form ONE B operand with
// N = 32 (NS - j) rows: a single MMA multiplies Theta slice j with all of them, and because slice s lands
// 32 s columns further right, its product falls on the accumulator columns of diagonal s + j.
// column maxima (as the bit patterns of non-negative doubles, which order like integers) + non-finite check
__global__ void __launch_bounds__(256) i8_colmax_kernel(const double* __restrict__ X, long long n, long long ld,
                                                       unsigned long long* __restrict__ colmax_bits, int* __restrict__ bad) {
  const int j = blockIdx.x;
  double m = 0.0;
  bool nf = false;
  for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.y * blockDim.x) {
    const double a = fabs(X[(size_t)j * ld + i]);
    if (!(a < 0x1p480)) nf = true;  // NaN, Inf or beyond the exponent window
    else m = fmax(m, a);
  }
  if (nf) atomicOr(bad, 1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(FM_FULL, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(&colmax_bits[j], (unsigned long long)__double_as_longlong(m));
}
// column exponents: x'_ij = x_ij 2^-cexp[j] in (-1, 1); theta'_j = theta_j 2^cexp[j] keeps every product unchanged
__global__ void i8_colexp_kernel(const unsigned long long* __restrict__ colmax_bits, int p_x, int* __restrict__ cexp) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < p_x) cexp[j] = i8_exponent(__longlong_as_double((long long)colmax_bits[j]));
}
// largest squared row norm max_i sum_j x_ij^2 (bit pattern of a non-negative double, like colmax): with Cauchy-Schwarz the
// second bound on |eta| of the un-clamped logistic epilogue
__global__ void __launch_bounds__(256) i8_rownorm_kernel(const double* __restrict__ X, long long n, long long ld, int p_x,
                                                        unsigned long long* __restrict__ rmax2_bits) {
  double m = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < p_x; j++) {
      const double x = X[(size_t)j * ld + i];
      s = fma(x, x, s);
    }
    m = fmax(m, s);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(FM_FULL, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(rmax2_bits, (unsigned long long)__double_as_longlong(m));
}

// sxy[j] = sum_i (y_i - 1/2) x_ij, fixed-order reduction (one CTA per column): the linear part of the binary
// logistic log-likelihood, sum_i [y_i eta_i - eta_i / 2] = theta . sxy, leaves the per-evaluation epilogue
__global__ void __launch_bounds__(1024) i8_sxy_kernel(const double* __restrict__ X, const double* __restrict__ y, long long n,
                                                     long long ld, double* __restrict__ sxy) {
  __shared__ double red[1024];
  const int j = blockIdx.x, t = threadIdx.x;
  double hi = 0.0, lo = 0.0;  // compensated (two-sum) partial
  for (long long i = t; i < n; i += 1024) {
    const double v = (y[i] - 0.5) * X[(size_t)j * ld + i];
    const double s = hi + v, bb = s - hi;
    lo += (hi - (s - bb)) + (v - bb);
    hi = s;
  }
  red[t] = hi + lo;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (t < o) red[t] += red[t + o];
    __syncthreads();
  }
  if (t == 0) sxy[j] = red[0];
}

template <int NS, int KB, int BLK>
__global__ void pack_i8_kernel(const double* __restrict__ X, long long n, long long ld, int p_x,
                               const int* __restrict__ cexp, unsigned char* __restrict__ Xq) {
  using G = I8Geom<NS, KB, BLK>;
  const long long tile = blockIdx.x;
  const int t = threadIdx.x;  // blockDim.x == TO
  const long long row = tile * G::TO + t;
  unsigned char* dst = Xq + (size_t)tile * G::STAGE_BYTES;
  const bool valid = row < n;
  unsigned char* grp = dst + (size_t)(t / G::BLK) * G::BLOCK_BYTES + ((t % G::BLK) / 8) * 256 + (t % 8) * 16;
  for (int kb = 0; kb < KB; kb++)
    for (int c = 0; c < 2; c++) {
      uint32_t w[NS][4];
#pragma unroll
      for (int i = 0; i < NS; i++) w[i][0] = w[i][1] = w[i][2] = w[i][3] = 0u;
#pragma unroll
      for (int q = 0; q < 16; q++) {
        const int j = kb * 32 + c * 16 + q;
        const double u = (valid && j < p_x) ? scalbn(X[(size_t)j * ld + row], -cexp[j]) : 0.0;
        int s[NS];
        i8_slices<NS>(u, s);
#pragma unroll
        for (int i = 0; i < NS; i++) w[i][q / 4] |= (uint32_t)(uint8_t)(int8_t)s[i] << (8 * (q % 4));
      }
#pragma unroll
      for (int i = 0; i < NS; i++)
        *reinterpret_cast<uint4*>(grp + (size_t)(kb * NS + i) * G::SLAB_BYTES + c * 128) = make_uint4(w[i][0], w[i][1], w[i][2], w[i][3]);
    }
}

// ---- tcgen05 wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor: K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}
// D[tmem] (+)= A * B, int8 x int8 -> int32; A from tensor memory or shared memory, B from shared memory
__device__ __forceinline__ void tc_mma_i8_ts(uint32_t d, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_i8_ss(uint32_t d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st_x8(uint32_t taddr, const uint32_t (&w)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(w[0]), "r"(w[1]),
               "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
template <int NW>
__device__ __forceinline__ void tc_st(uint32_t taddr, const uint32_t (&w)[NW]) {
  static_assert(NW == 2 || NW == 4 || NW == 8, "2, 4 or 8 columns per store");
  if constexpr (NW == 8) tc_st_x8(taddr, w);
  else if constexpr (NW == 4)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
  else
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(w[0]), "r"(w[1]) : "memory");
}
template <int CH>
__device__ __forceinline__ void tc_ld(uint32_t taddr, uint32_t (&v)[CH]) {
  static_assert(CH == 4 || CH == 8, "4 or 8 columns per load");
  if (CH == 8)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4 % CH]), "=r"(v[5 % CH]), "=r"(v[6 % CH]), "=r"(v[7 % CH])
                 : "r"(taddr)
                 : "memory");
  else
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
}
// all NS accumulators of 8 columns (diagonal d sits 32 d columns further right) in ONE asm statement: the address reaches
// the uniform datapath once (one R2UR) and the NS loads carry immediate offsets, instead of one R2UR per load
template <int NS, int CH, int BLK = 32>
__device__ __forceinline__ void tc_ld_diagonals(uint32_t taddr, uint32_t (&a)[NS][CH]) {
  if constexpr (CH == 8 && NS == 6 && BLK == 32) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%48];\n"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%48+32];\n"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%48+64];\n"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%24,%25,%26,%27,%28,%29,%30,%31}, [%48+96];\n"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%32,%33,%34,%35,%36,%37,%38,%39}, [%48+128];\n"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%40,%41,%42,%43,%44,%45,%46,%47}, [%48+160];\n"
        : "=r"(a[0][0]), "=r"(a[0][1]), "=r"(a[0][2]), "=r"(a[0][3]), "=r"(a[0][4]), "=r"(a[0][5]), "=r"(a[0][6]), "=r"(a[0][7]), "=r"(a[1][0]), "=r"(a[1][1]), "=r"(a[1][2]), "=r"(a[1][3]), "=r"(a[1][4]), "=r"(a[1][5]), "=r"(a[1][6]), "=r"(a[1][7]), "=r"(a[2][0]), "=r"(a[2][1]), "=r"(a[2][2]), "=r"(a[2][3]), "=r"(a[2][4]), "=r"(a[2][5]), "=r"(a[2][6]), "=r"(a[2][7]), "=r"(a[3][0]), "=r"(a[3][1]), "=r"(a[3][2]), "=r"(a[3][3]), "=r"(a[3][4]), "=r"(a[3][5]), "=r"(a[3][6]), "=r"(a[3][7]), "=r"(a[4][0]), "=r"(a[4][1]), "=r"(a[4][2]), "=r"(a[4][3]), "=r"(a[4][4]), "=r"(a[4][5]), "=r"(a[4][6]), "=r"(a[4][7]), "=r"(a[5][0]), "=r"(a[5][1]), "=r"(a[5][2]), "=r"(a[5][3]), "=r"(a[5][4]), "=r"(a[5][5]), "=r"(a[5][6]), "=r"(a[5][7])
        : "r"(taddr)
        : "memory");
  } else if constexpr (CH == 8 && NS == 7 && BLK == 32) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%56];\n"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%56+32];\n"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%56+64];\n"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%24,%25,%26,%27,%28,%29,%30,%31}, [%56+96];\n"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%32,%33,%34,%35,%36,%37,%38,%39}, [%56+128];\n"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%40,%41,%42,%43,%44,%45,%46,%47}, [%56+160];\n"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%48,%49,%50,%51,%52,%53,%54,%55}, [%56+192];\n"
        : "=r"(a[0][0]), "=r"(a[0][1]), "=r"(a[0][2]), "=r"(a[0][3]), "=r"(a[0][4]), "=r"(a[0][5]), "=r"(a[0][6]), "=r"(a[0][7]), "=r"(a[1][0]), "=r"(a[1][1]), "=r"(a[1][2]), "=r"(a[1][3]), "=r"(a[1][4]), "=r"(a[1][5]), "=r"(a[1][6]), "=r"(a[1][7]), "=r"(a[2][0]), "=r"(a[2][1]), "=r"(a[2][2]), "=r"(a[2][3]), "=r"(a[2][4]), "=r"(a[2][5]), "=r"(a[2][6]), "=r"(a[2][7]), "=r"(a[3][0]), "=r"(a[3][1]), "=r"(a[3][2]), "=r"(a[3][3]), "=r"(a[3][4]), "=r"(a[3][5]), "=r"(a[3][6]), "=r"(a[3][7]), "=r"(a[4][0]), "=r"(a[4][1]), "=r"(a[4][2]), "=r"(a[4][3]), "=r"(a[4][4]), "=r"(a[4][5]), "=r"(a[4][6]), "=r"(a[4][7]), "=r"(a[5][0]), "=r"(a[5][1]), "=r"(a[5][2]), "=r"(a[5][3]), "=r"(a[5][4]), "=r"(a[5][5]), "=r"(a[5][6]), "=r"(a[5][7]), "=r"(a[6][0]), "=r"(a[6][1]), "=r"(a[6][2]), "=r"(a[6][3]), "=r"(a[6][4]), "=r"(a[6][5]), "=r"(a[6][6]), "=r"(a[6][7])
        : "r"(taddr)
        : "memory");
  } else {
#pragma unroll
    for (int d = 0; d < NS; d++) tc_ld<CH>(taddr + d * BLK, a[d]);
  }
}
// wait for this thread's tcgen05.ld's; the registers are threaded through so no use can be scheduled above it
template <int NS, int CH>
__device__ __forceinline__ void tc_wait_ld(uint32_t (&a)[NS][CH]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int d = 0; d < NS; d++)
#pragma unroll
    for (int e = 0; e < CH; e++) asm volatile("" : "+r"(a[d][e]));
}
// one lane of a CONVERGED warp: ptxas then keeps descriptors / addresses in uniform registers and issues the
// tcgen05 / TMA instruction directly (inside an `if (lane == 0)` region it wraps every UTCIMMA in an
// ELECT + R2UR loop, ~80 clk per instruction - profiles/r01_i8_findings.md)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}
// mbarrier wait for the single-thread roles: back off instead of spinning in the issue slots the epilogue needs
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned ns) {
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    __nanosleep(ns);
  }
}

// scheduling fence for 16 doubles: every later use of tv[] depends on this statement, and it on all of tv[]
__device__ __forceinline__ void i8_pin16(double (&t)[16]) {
  asm volatile("" : "+d"(t[0]), "+d"(t[1]), "+d"(t[2]), "+d"(t[3]), "+d"(t[4]), "+d"(t[5]), "+d"(t[6]), "+d"(t[7]), "+d"(t[8]),
                    "+d"(t[9]), "+d"(t[10]), "+d"(t[11]), "+d"(t[12]), "+d"(t[13]), "+d"(t[14]), "+d"(t[15]));
}
// t = sum_d a_d 2^(7 (NS - 1 - d)), exact: adjacent diagonals are merged in int32 (|a_d| <= (d+1) * 32 KB * 4096 < 2^24),
// up to three merged pairs in int64 (32 x 32 -> 64-bit multiply-adds), then ONE conversion per group.
// (Measured alternatives, clk per warp-evaluation of the merge alone - profiles/microbench/i8_epilogue_rate.cu:
//  I2F.F64.S64 27.9 | magic-number int64 30.1 | int32 pairs + 3 I2F.S32 + 2 DFMA 28.8 | 3 magic32 + 2 DFMA 30.9 |
//  6 magic32 + 5 DFMA 42.3.)
template <int NS, int CH, int KB = 1>
__device__ __forceinline__ double i8_assemble(const uint32_t (&a)[NS][CH], int e, int tune = 0) {
  constexpr int DB = I8_DIGIT_BITS;
  // |a_d| <= (d + 1) 32 KB 2^(2 DB - 2); a merged pair a_{2p} 2^DB + a_{2p+1} stays in int32 while that is below 2^31
  constexpr long long UNIT = 32LL * KB * (1LL << (2 * DB - 2));
  if constexpr (NS == 5 && DB == 8 && 4 * UNIT * 256 + 5 * UNIT < (1LL << 31)) {
    // five diagonals, K <= 64: pairs (a1, a2) and (a3, a4) in int32, the unpaired diagonal on TOP - a0 2^32 is an add into the
    // high word, where an unpaired a4 at the bottom needs a sign extension and a register move to become a 64-bit addend:
    // 2 IMAD + SHF + IMAD.WIDE + IADD + I2F
    const int v1 = (int)a[1][e] * 256 + (int)a[2][e];
    const int v2 = (int)a[3][e] * 256 + (int)a[4][e];
    // the 64-bit addend of the multiply-add is (v2, sign(v2) + a0): a0 2^32 rides on the sign extension of v2
    const int hi0 = (v2 >> 31) + (int)a[0][e];
    long long addend;
    asm("mov.b64 %0, {%1, %2};" : "=l"(addend) : "r"(v2), "r"(hi0));
    const long long w = (long long)v1 * 65536LL + addend;
    return __ll2double_rn(w);
  } else if (NS <= 6) {
    constexpr int NV = (NS + 1) / 2;
    long long w;
    if (NS & 1) {
      w = (int)a[NS - 1][e];
    } else if ((NS - 1) * UNIT * (1LL << DB) + NS * UNIT < (1LL << 31)) {
      w = (int)a[NS - 2][e] * (1 << DB) + (int)a[NS - 1][e];
    } else {  // 8-bit digits, NS = 6, K = 128: the last pair needs 64 bits
      w = (long long)(int)a[NS - 2][e] * (1 << DB) + (int)a[NS - 1][e];
    }
#pragma unroll
    for (int p = NV - 2; p >= 0; p--) {
      static_assert((NS - 2) * UNIT * (1LL << DB) + (NS - 1) * UNIT < (1LL << 31) || NS == 6, "inner pairs fit int32");
      const int v = (int)a[2 * p][e] * (1 << DB) + (int)a[2 * p + 1][e];
      w += (long long)v * (1LL << (DB * (NS - 2 - 2 * p)));
    }
#ifdef FMCMC_I8_TUNE_HOOKS
    if (tune & 8) return __hiloint2double(0x43300000 | (int)((w >> 32) & 0xfffff), (int)w) - 4503599627370496.0;  // no XU op (value is garbage)
#endif
    return __ll2double_rn(w);
  } else {  // 7, 8 slices of 7 bits: two groups
    int v[(NS + 1) / 2];
#pragma unroll
    for (int p = 0; p < NS / 2; p++) v[p] = (int)a[2 * p][e] * 128 + (int)a[2 * p + 1][e];
    if (NS & 1) v[NS / 2] = (int)a[NS - 1][e];
    const long long hi = (long long)v[0] * 16384LL + v[1];
    const long long lo = (long long)v[2] * ((NS & 1) ? 128LL : 16384LL) + v[3];
    return fma(__ll2double_rn(hi), (NS & 1) ? 2097152.0 : 268435456.0, __ll2double_rn(lo));
  }
}

// Polynomial cores of the softplus tables (softplus.h: fm_softplus_tab4_core / fm_softplus_tab8_core) with the coefficients
// in the constant bank (FP64 instructions take c[bank][offset] operands; literals cost a UMOV / IMAD.MOV pair per use).
// Given the remainder d and the table entry S: v = S expm1(-d) and L = log1p(v) / v, so that log1p(exp(-a)) = G + v L.
__constant__ double I8_K1[8] = {1.0 / 120.0, -1.0 / 24.0, 1.0 / 6.0, -0.5, 1.0, 0.2, -0.25, 1.0 / 3.0};   // level 1: Taylor, degree 4
__constant__ double I8_K2[8] = {-0x1.5555582d82db0p-5, 0x1.55555999999f5p-3, -0x1.fffffffffffd2p-2, 0x1.fffffffffff77p-1,
                                -0x1.00000b2f503bap-2, 0x1.555562c14f2f7p-2, -0x1.ffffffffffe89p-2, 0x1.fffffffffff1fp-1};  // level 2: cubic
template <int TL>
__device__ __forceinline__ void i8_softplus_core(double d, double S, double& v, double& L) {
  if (TL == 2) {
    double q = I8_K2[0];
    q = fma(q, d, I8_K2[1]);
    q = fma(q, d, I8_K2[2]);
    q = fma(q, d, I8_K2[3]);
    v = (S * d) * -q;
    L = I8_K2[4];
    L = fma(L, v, I8_K2[5]);
    L = fma(L, v, I8_K2[6]);
    L = fma(L, v, I8_K2[7]);
  } else {
    double q = I8_K1[0];
    q = fma(q, d, I8_K1[1]);
    q = fma(q, d, I8_K1[2]);
    q = fma(q, d, I8_K1[3]);
    q = fma(q, d, I8_K1[4]);
    v = (S * d) * -q;
    L = I8_K1[5];
    L = fma(L, v, I8_K1[6]);
    L = fma(L, v, I8_K1[7]);
    L = fma(L, v, I8_K1[3]);
    L = fma(L, v, I8_K1[4]);
  }
}
// 1.5 * 2^(52 - log2 H): adding a >= 0 to it rounds a to a multiple of 1 / H and leaves round(H a) in the low word
template <int TL>
__device__ __forceinline__ constexpr double i8_magic_h() { return TL == 2 ? 26388279066624.0 : 52776558133248.0; }

// Binary logistic regression without the per-observation response: with z = +-eta,
//   sum_i [min(z_i, 0) - log1p(exp(-|z_i|))] = theta . sxy - sum_i [ |eta_i| / 2 + log1p(exp(-|eta_i|)) ]
// (y_i eta_i - max(eta_i, 0) = (y_i - 1/2) eta_i - |eta_i| / 2), so the epilogue only accumulates the even function of
// eta; no select, no load of y.  The scaling is folded in: eta = t csc with t the integer-valued double from the
// reassembly and csc a power of two, so |eta| is never formed - adding |t| csc to the magic constant rounds it to a multiple
// of 1 / H (k in the low word), the remainder d = |t| csc - k / H is one more fma, and sum |eta| = csc sum |t| (scaled once
// per chain at the end).  14 FP64 instructions with the level-2 table, 16 with level 1.
// SAFE = false (the warp's chains all have |theta'|max < 2^20 / KB, hence |eta| < 2^25): no clamp of the argument at all -
// round(H |eta|) then fits the low word, the table INDEX is clamped (one unsigned min; beyond AMAX the entry is
// G(AMAX) = 4e-18 with a remainder |d| <= 1 / 2H, i.e. the right answer to 4e-18), and |t| enters the two fmas through the
// free source modifier.  SAFE = true clamps |t| itself on its high word (AMAX / csc = hi_clamp) for arbitrary magnitudes.
template <int TL, bool SAFE>
__device__ __forceinline__ void i8_logistic_even_t(double t, double csc, int hi_clamp, double& acc_abs_t, double& acc_g,
                                                   const double2* __restrict__ tab, int tune = 0) {
  constexpr double MAGICH = i8_magic_h<TL>();
  constexpr unsigned KMAX = (TL == 2 ? FM_SP8_ENTRIES : FM_SP4_ENTRIES) - 1u;
  const double tc = SAFE ? __hiloint2double(min(__double2hiint(t) & 0x7fffffff, hi_clamp), __double2loint(t)) : fabs(t);
  const double t2 = fma(tc, csc, MAGICH);
  const int k = SAFE ? __double2loint(t2) : (int)min((unsigned)__double2loint(t2), KMAX);
  const double d = fma(tc, csc, MAGICH - t2);
#ifdef FMCMC_I8_TUNE_HOOKS
  const double2 sg = (tune & 4) ? make_double2(d * 0.25, d) : tab[k];
#else
  const double2 sg = tab[k];
#endif
  double v, L;
  i8_softplus_core<TL>(d, sg.x, v, L);
  acc_abs_t += fabs(t);
  acc_g = fma(v, L, acc_g + sg.y);
}
// Level-2 table (softplus.h: fm_lcosh_table8_fill): entry k = (tau_k, T_k) of h(a) = log(2 cosh(a / 2)) = a / 2 + log1p(exp(-a)),
// the even function the binary-logistic epilogue sums.  h(k/256 + d) = T + d (tau + d u (1/2 + d (-tau/3 + d (1/24 - u/4)))),
// u = 1/4 - tau^2 (degree-4 Taylor; every derivative of h is a polynomial in tau): 9 FP64 instructions including the
// accumulation, 12 with the range reduction, and no separate sum of |eta|.
__device__ __forceinline__ double i8_lcosh_core(double d, double tau, double Tacc) {
  const double u = fma(-tau, tau, 0.25);
  const double i1 = fma(u, -0.25, 1.0 / 24.0);
  const double i2 = fma(d, i1, tau * (-1.0 / 3.0));
  const double i3 = fma(d, i2, 0.5);
  const double P = u * i3;
  const double Q = fma(d, P, tau);
  return fma(d, Q, Tacc);
}
// degree-3 core on the mean-corrected table (softplus.h, fm_lcosh_table8m_fill): 7 FP64 instructions including the accumulation
__device__ __forceinline__ double i8_lcosh3_core(double d, double tau, double Tacc) {
  const double u = fma(-tau, tau, 0.25);
  const double i3 = fma(d, tau * (-1.0 / 3.0), 0.5);
  const double P = u * i3;
  const double Q = fma(d, P, tau);
  return fma(d, Q, Tacc);
}
// tab_s: the table's 32-bit shared-window address (computed once per kernel: through a generic pointer ptxas re-derives the
// window base - S2UR SR_CgaCtaId + 3 uniform instructions - in every chunk)
__device__ __forceinline__ void i8_logistic_lcosh3_fast(double t, double csc, double& acc_h, uint32_t tab_s) {
  constexpr double MAGICH = i8_magic_h<2>();
  const double t2 = fma(fabs(t), csc, MAGICH);
  const double d = fma(fabs(t), csc, MAGICH - t2);
  double tx, ty;
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tx), "=d"(ty) : "r"(tab_s + ((uint32_t)__double2loint(t2) << 4)));
  acc_h = i8_lcosh3_core(d, tx, acc_h + ty);
}
// level 3, the bank-group-replicated cubic table: tab_lane = table address + 16 (lane % 8), points 256 B apart, (c1, c0) in
// the first 128 bytes and (c2, c3) in the second - lane l always reads bank group l % 8, so each of the two loads costs
// 4 wavefronts whatever the indices are.  7 FP64 instructions per evaluation (3 range reduction + 4) instead of 10.
__device__ __forceinline__ void i8_logistic_cubic_rep(double t, double csc, double& acc_h, uint32_t tab_lane) {
  constexpr double MAGICH = 105553116266496.0;   // 1.5 * 2^46: ulp = 1 / 64
  const double t2 = fma(fabs(t), csc, MAGICH);
  const double d = fma(fabs(t), csc, MAGICH - t2);
  const uint32_t addr = tab_lane + ((uint32_t)__double2loint(t2) << 8);
  double c1, c0, c2, c3;
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c1), "=d"(c0) : "r"(addr));
  asm("ld.shared.v2.f64 {%0, %1}, [%2+128];" : "=d"(c2), "=d"(c3) : "r"(addr));
  const double q1 = fma(d, c3, c2);
  const double q2 = fma(d, q1, c1);
  acc_h = fma(d, q2, acc_h + c0);
}
// FAST: the warp's chains bound |eta| <= 39.9 over ALL observations (i8 prologue: min(sum_j |theta_j| max_i |x_ij|,
// |theta|_2 max_i |x_i|_2)), so round(256 |eta|) indexes the table as it is - no clamp of any kind: 12 FP64 + 7 (merge) +
// address + LDS.128 per evaluation.
__device__ __forceinline__ void i8_logistic_lcosh_fast(double t, double csc, double& acc_h, const double2* __restrict__ tab) {
  constexpr double MAGICH = i8_magic_h<2>();
  const double t2 = fma(fabs(t), csc, MAGICH);
  const double d = fma(fabs(t), csc, MAGICH - t2);
  const double2 tt = tab[__double2loint(t2)];
  acc_h = i8_lcosh_core(d, tt.x, acc_h + tt.y);
}
// SAFE: arbitrary magnitudes.  |t| is clamped on its high word (AMAX / csc = hi_clamp), sum |t| is accumulated apart and
// g = h(a_c) - a_c / 2 = log1p(exp(-a_c)) enters the second accumulator: sum h = (csc / 2) sum |t| + sum g.
__device__ __forceinline__ void i8_logistic_lcosh_safe(double t, double csc, int hi_clamp, double& acc_abs_t, double& acc_g,
                                                       const double2* __restrict__ tab) {
  constexpr double MAGICH = i8_magic_h<2>();
  const double tc = __hiloint2double(min(__double2hiint(t) & 0x7fffffff, hi_clamp), __double2loint(t));
  const double t2 = fma(tc, csc, MAGICH);
  const double d = fma(tc, csc, MAGICH - t2);
  const double2 tt = tab[__double2loint(t2)];
  const double h = i8_lcosh_core(d, tt.x, tt.y);
  acc_abs_t += fabs(t);
  acc_g += fma(tc, -0.5 * csc, h);
}
// general response (sum(logp[y == 1]) + sum(logq[y == 0]), anything else contributes nothing), NaN propagated like
// logistic_term_tab (families.cuh)
template <int TL>
__device__ __forceinline__ double i8_logistic_term(double eta, double y, const double2* __restrict__ tab) {
  constexpr double MAGICH = i8_magic_h<TL>();
  // |eta| clamped on the high word alone: >= 40 -> [40, 40 + 2^-15], the last table entry with a tiny remainder
  const double a = __hiloint2double(min(__double2hiint(eta) & 0x7fffffff, 0x40440000), __double2loint(eta));
  const double t2 = a + MAGICH;
  const double d = a + (MAGICH - t2);
  const double2 sg = tab[__double2loint(t2)];
  double g;
  if (TL == 2) {
    g = fma(-0.5, a, i8_lcosh_core(d, sg.x, sg.y));  // log1p(exp(-a)) = h(a) - a / 2
  } else {
    double v, L;
    i8_softplus_core<TL>(d, sg.x, v, L);
    g = fma(v, L, sg.y);
  }
  const double z = (y == 1.0) ? eta : -eta;
  const double r = fm_min0(z) - g;
  return (y == 1.0 || y == 0.0) ? r : 0.0;
}

template <int TL>
__device__ __forceinline__ void i8_logistic_safe(double t, double csc, int hi_clamp, double& acc_abs_t, double& acc_g,
                                                 const double2* __restrict__ tab) {
  if (TL == 2) i8_logistic_lcosh_safe(t, csc, hi_clamp, acc_abs_t, acc_g, tab);
  else i8_logistic_even_t<1, true>(t, csc, hi_clamp, acc_abs_t, acc_g, tab);
}

template <int FAMILY, bool YBIN, int NS, int KB, int EW, int CH>
__global__ void __launch_bounds__((EW + 2) * 32, 1)
tiled_loglik_i8_kernel(ModelParams mp, const double* __restrict__ prop, const double* __restrict__ prop_u, int C,
                       TiledBuffers tb, const int* __restrict__ err) {
  using G = I8Geom<NS, KB, i8_blk<NS, KB>(FAMILY)>;
  static_assert(NS >= 2 && NS <= 8, "2..8 slices");
  static_assert(I8_DIGIT_BITS == 7 || (I8_DIGIT_BITS == 8 && NS <= 6), "8-bit digits: at most 6 slices (int64 merge)");
  static_assert(EW == 8 || EW == 16, "2 or 4 epilogue warps per TMEM lane quarter");
  constexpr int CW = G::BLK / (EW / 4);  // columns (observations) of a block owned by one epilogue warp (grouped epilogue: twice that)
  static_assert((((I8_GROUPED != 0 && EW == 16 && G::NACC >= 2) ? 2 : 1) * CW) % CH == 0, "whole chunks");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  constexpr int STAGES = i8_stages<NS, KB>(FAMILY, YBIN);
  constexpr int NACC = G::NACC;
  static_assert(STAGES >= 2 && 2 * STAGES + 2 * NACC + 2 <= 32, "barriers live in the first 256 bytes");
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint64_t* acc_empty = acc_full + NACC;
  uint64_t* tab_bar = acc_empty + NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tab_bar + 1);
  unsigned char* stage0 = smem_raw + 256;
  unsigned char* sA = stage0 + (size_t)STAGES * G::STAGE_BYTES;
  double2* sp_tab = reinterpret_cast<double2*>(sA + G::A_SMEM_BYTES);
  double* red = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(sp_tab) +
                                          (size_t)i8_table_bytes<NS, KB>(FAMILY, YBIN));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int W_TMA = EW, W_MMA = EW + 1;
  // the warp index through redux.sync (CREDUX writes a uniform register): ptxas then KNOWS that the role branches below are
  // warp-uniform and keeps warp-uniform values - the TMEM addresses of the epilogue's tcgen05.ld's above all - in uniform
  // registers (otherwise: one R2UR per load; a shuffle broadcast does not convince it)
  const int warp_u = (int)__reduce_min_sync(FM_FULL, (unsigned)warp);
  // A CTA owns one block of 128 chains and tb.gsl >= 1 consecutive observation slices, which it walks one after the other and
  // flushes one by one: partial[slice][chain] - hence every bit of the log-posterior - is the same whatever gsl is, but the
  // per-CTA costs (launch, tensor-memory allocation, Theta slicing, the table's bulk copy, pipeline fill and drain: ~6 % of a
  // cfg3 launch with one slice per CTA) are paid once per gsl slices.
  const int gsl = tb.gsl > 0 ? tb.gsl : 1;
  const int chain_block = (int)(blockIdx.x % (unsigned)tb.cb), slice0 = (int)(blockIdx.x / (unsigned)tb.cb) * gsl;
  const int slice = slice0;
  const long long ntiles = (mp.n + G::TO - 1) / G::TO;
  const long long first = slice, step = tb.gx;
  const int p_x = mp.p_x;

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);  // released by the MMA warp's tcgen05.commit alone
    }
    for (int b = 0; b < NACC; b++) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], (I8_GROUPED != 0 && EW == 16 && NACC >= 2) ? EW / 2 : EW);
    }
    mbar_init(tab_bar, 1);
    mbar_fence_init();
  }
  if (warp == W_MMA) {  // the allocating warp also frees
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // X does not depend on the head kernel of this row: the producer requests the first stages BEFORE waiting for it
  // (programmatic dependent launch: this CTA may be resident while the head kernel still runs)
  int n_early = 0;
  if (warp_u == W_TMA) {
    for (long long tile = first; tile < ntiles && n_early < STAGES; tile += step, n_early++) {
      if (elect_one()) {
        mbar_expect_tx(&full[n_early], (uint32_t)G::STAGE_BYTES);
        bulk_g2s(stage0 + (size_t)n_early * G::STAGE_BYTES, mp.Xq + (size_t)tile * G::STAGE_BYTES, (uint32_t)G::STAGE_BYTES, &full[n_early]);
      }
      __syncwarp();
    }
  }
  pdl_launch_dependents();
  pdl_wait();
  if (err[0] != 0) {  // a previous row failed: let the copies in flight land, free tensor memory, leave
    if (warp_u == W_TMA)
      for (int s = 0; s < n_early; s++) mbar_wait(&full[s], 0u);
    if (warp == W_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    return;
  }

  // ---- this thread's chain: exponent, bounds, intercept.  The EW / 4 warps of a TMEM lane quarter hold the same 32 chains:
  // each takes 32 / NPART columns of every K block (loads batched by the unrolled loops), the partial maxima / sums meet in
  // shared memory (the table region, which is not loaded yet; a scratch region of its own for the Gaussian family) ----
  const int icpt = (FAMILY == FMCMC_FAMILY_GAUSSIAN_LM && (mp.flags & FMCMC_MODEL_INTERCEPT)) ? 1 : 0;
  const int lchain = tid & (I8_CHAINS - 1);
  const int col = chain_block * I8_CHAINS + lchain;
  constexpr int NPART = EW / 4, CQ = 32 / NPART;   // parts a chain's columns are split into; columns of a K block per part
  const int part = warp_u >> 2;                   // (producer / MMA warps: no chain)
  const double* th = nullptr;
  if (warp < EW && col < tb.ncols) th = col < C ? prop + (size_t)col * mp.k : prop_u + (size_t)(col - C) * mp.k;
  double thmax = 0.0, b0 = 0.0, lin = 0.0, eb1 = 0.0, eb2 = 0.0;
  bool th_nan = false, th_big = false;
  double* pscr = reinterpret_cast<double*>(FAMILY == FMCMC_FAMILY_LOGISTIC ? reinterpret_cast<unsigned char*>(sp_tab)
                                                                           : reinterpret_cast<unsigned char*>(red) + (I8_MAX_EPI_WARPS / 4) * I8_CHAINS * sizeof(double));
  if (warp < EW) {
    double p_max = 0.0, p_e1 = 0.0, p_e2 = 0.0, p_lin = 0.0;
    int p_flags = 0;
    if (th) {
#pragma unroll
      for (int kb = 0; kb < KB; kb++) {
        double v[CQ];
        int ce[CQ];
#pragma unroll
        for (int q = 0; q < CQ; q++) {   // CQ independent loads in flight
          const int j = kb * 32 + part * CQ + q;
          v[q] = j < p_x ? th[icpt + j] : 0.0;
          ce[q] = j < p_x ? mp.i8_cexp[j] : 0;
        }
#pragma unroll
        for (int q = 0; q < CQ; q++) {
          const int j = kb * 32 + part * CQ + q;
          if (j < p_x) {
            const double a = fabs(v[q] * __hiloint2double((1023 + ce[q]) << 20, 0));  // theta'_j = theta_j 2^cexp[j], |cexp| <= 480
            if (a != a) p_flags |= 1;
            else if (!(a < 0x1p480)) p_flags |= 2;
            p_max = fmax(p_max, a);
            p_e1 = fma(fabs(v[q]), mp.i8_cmax[j], p_e1);  // |eta_i| <= sum_j |theta_j| max_i |x_ij|
            p_e2 = fma(v[q], v[q], p_e2);                 // |eta_i| <= |theta|_2 max_i |x_i|_2
            if (FAMILY == FMCMC_FAMILY_LOGISTIC && YBIN) p_lin = fma(v[q], mp.i8_sxy[j], p_lin);
          }
        }
      }
      if (icpt) b0 = th[0];
    }
    double* ps = pscr + (size_t)(part * I8_CHAINS + lchain) * 5;
    ps[0] = p_max; ps[1] = p_e1; ps[2] = p_e2; ps[3] = p_lin; ps[4] = __hiloint2double(0, p_flags);
  }
  __syncthreads();
  if (warp < EW) {
    int fl = 0;
#pragma unroll
    for (int pp = 0; pp < NPART; pp++) {   // fixed order: every warp of the quarter forms the same values
      const double* ps = pscr + (size_t)(pp * I8_CHAINS + lchain) * 5;
      thmax = fmax(thmax, ps[0]);
      eb1 += ps[1];
      eb2 += ps[2];
      lin += ps[3];
      fl |= __double2loint(ps[4]);
    }
    th_nan = (fl & 1) != 0;
    th_big = (fl & 2) != 0;
  }
  // the region the partial sums went through is about to be written by the table's bulk copy (async proxy)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const bool th_bad = th_nan || th_big;
  const int eth = th_bad ? 0 : i8_exponent(thmax);
  const double csc = __hiloint2double((1023 - G::SHIFT + eth) << 20, 0);  // eta = t * 2^(eth - SHIFT)
  // high word of AMAX / csc (AMAX = 40 = 1.25 * 2^5)
  const int hi_clamp = (0x40440000) + ((G::SHIFT - eth) << 20);
  // |eta| <= 32 KB max|theta'| < 2^(5 + log2 KB + eth): below 2^25 for every chain of the warp, the unclamped epilogue applies
  const bool eta_small = __all_sync(FM_FULL, eth + 5 + (KB == 1 ? 0 : (KB == 2 ? 1 : 2)) <= 25);
  // level-2 table: |eta| <= 39.9 for every observation and every chain of the warp (idle lanes: 0; NaN fails the test) -
  // the un-clamped lcosh epilogue applies (the slicing error of eta, <= 2^-35 thmax, is far inside the 0.1 margin)
  const bool in_table = !th_bad && fmin(eb1, sqrt(eb2 * mp.i8_cmax[p_x])) <= 39.9;
  const bool eta_in_table = __all_sync(FM_FULL, in_table);
  // ... and for every chain of the CTA: the hot loop runs the degree-3 core on the mean-corrected copy of the table
  // (softplus.h).  The table is chosen per CTA because it is shared: one bulk copy (160 KB), overlapped with the Theta slicing.
  const bool cta_in_table = (FAMILY == FMCMC_FAMILY_LOGISTIC && YBIN && i8_table_level<NS, KB>() == 2) ? __all_sync(FM_FULL, __syncthreads_and(in_table && !tb.exact_core) != 0) : false;  // (the vote tells ptxas it is warp-uniform)
  // ... and below the end of the replicated table (level 3) for every chain of the CTA: the conflict-free gather
  constexpr int REP_ENTRIES = (FAMILY == FMCMC_FAMILY_LOGISTIC && YBIN) ? i8_rep_entries<NS, KB>() : 0;
  constexpr double REP_BOUND = REP_ENTRIES > 0 ? (double)(REP_ENTRIES - 1) / FM_LC6_H - 0.1 : -1.0;
  const bool in_rep = !th_bad && fmin(eb1, sqrt(eb2 * mp.i8_cmax[p_x])) <= REP_BOUND;
  const bool cta_rep = REP_ENTRIES > 0 ? __all_sync(FM_FULL, __syncthreads_and(in_rep && !tb.exact_core) != 0) : false;
  // (every thread's reads of the partial sums in the table region precede a block-wide barrier: the votes above, or this one)
  if (FAMILY == FMCMC_FAMILY_LOGISTIC && !(YBIN && i8_table_level<NS, KB>() == 2)) __syncthreads();
  if (FAMILY == FMCMC_FAMILY_LOGISTIC && tid == 0) {
    constexpr uint32_t TAB_BYTES = (uint32_t)i8_table_entries<NS, KB>() * 16u, REP_BYTES = (uint32_t)REP_ENTRIES * FM_LC6_POINT_BYTES;
    mbar_expect_tx(tab_bar, cta_rep ? REP_BYTES : TAB_BYTES);
    if (cta_rep) bulk_g2s(sp_tab, mp.sp_tab6r, REP_BYTES, tab_bar);
    else bulk_g2s(sp_tab, i8_table_level<NS, KB>() == 2 ? (cta_in_table ? mp.sp_tab8m : mp.sp_tab8) : mp.sp_tab4, TAB_BYTES, tab_bar);
  }
  if (warp < EW) {   // Theta slices into the A operand: every epilogue warp writes its CQ columns of its lane quarter's chains
    const uint32_t lane_quarter = (uint32_t)((warp_u & 3) * 32) << 16;
#pragma unroll
    for (int kb = 0; kb < KB; kb++) {
      uint32_t w[NS][CQ / 4];
#pragma unroll
      for (int i = 0; i < NS; i++)
#pragma unroll
        for (int q = 0; q < CQ / 4; q++) w[i][q] = 0u;
      double v[CQ];
      int ce[CQ];
#pragma unroll
      for (int q = 0; q < CQ; q++) {
        const int j = kb * 32 + part * CQ + q;
        const bool live = th && !th_bad && j < p_x;
        v[q] = live ? th[icpt + j] : 0.0;
        ce[q] = live ? mp.i8_cexp[j] : 0;
      }
#pragma unroll
      for (int q = 0; q < CQ; q++) {
        const double u = v[q] * __hiloint2double((1023 + ce[q] - eth) << 20, 0);  // exact: |cexp - eth| <= 960, |u| < 1
        int s[NS];
        i8_slices<NS>(u, s);
#pragma unroll
        for (int i = 0; i < NS; i++) w[i][q / 4] |= (uint32_t)(uint8_t)(int8_t)s[i] << (8 * (q % 4));
      }
#pragma unroll
      for (int i = 0; i < NS; i++) {
        if (G::a_in_tmem(i, kb)) {
          tc_st<CQ / 4>(tmem + G::a_tmem_col(i, kb) + part * (CQ / 4) + lane_quarter, w[i]);
        } else {   // row = chain: 16-byte chunks of 16 columns, the second half of the K block 128 B further
          unsigned char* p = sA + (size_t)G::a_smem_unit(i, kb) * 4096 + (lchain / 8) * 256 + (lchain % 8) * 16 +
                             ((part * CQ) / 16) * 128 + (part * CQ) % 16;
#pragma unroll
          for (int q = 0; q < CQ / 4; q++) reinterpret_cast<uint32_t*>(p)[q] = w[i][q];
        }
      }
    }
    if (G::A_TKB > 0) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (G::A_SMEM_BYTES > 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp_u == W_TMA) {
    // ===== producer: one bulk copy per stage (whole warp in the loop, one elected lane issues) =====
    long long it = n_early;  // the first stages (of the first slice) are already on their way
    for (int sl = 0; sl < gsl; sl++)
    for (long long tile = first + sl + (sl == 0 ? (long long)n_early * step : 0); tile < ntiles; tile += step, it++) {
      const int s = (int)(it % STAGES);
      const uint32_t ph = (uint32_t)((it / STAGES) & 1);
      mbar_wait_sleep(&empty[s], ph ^ 1u, 256);
      if (elect_one()) {
        mbar_expect_tx(&full[s], (uint32_t)G::STAGE_BYTES);
        bulk_g2s(stage0 + (size_t)s * G::STAGE_BYTES, mp.Xq + (size_t)tile * G::STAGE_BYTES, (uint32_t)G::STAGE_BYTES, &full[s]);
      }
      __syncwarp();
    }
  } else if (warp_u == W_MMA) {
    // ===== MMA issuer: NS * KB instructions per block of 32 observations (N = 32 (NS - j)); whole warp in the
    // loop, one elected lane issues =====
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = S32, A = B = signed int8, K-major, N >> 3, M >> 4
    constexpr uint32_t IDESC0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(I8_CHAINS >> 4) << 24);
    const uint32_t sA_addr = smem_u32(sA);
    uint32_t buf = 0, par = 0;  // accumulator set of the next block and the phase of its barriers (block n: n % NACC, (n / NACC) & 1)
    long long it = 0;
    for (int sl = 0; sl < gsl; sl++)
    for (long long tile = first + sl; tile < ntiles; tile += step, it++) {
      const int s = (int)(it % STAGES);
      mbar_wait_sleep(&full[s], (uint32_t)((it / STAGES) & 1), 32);
      tc_fence_after();
      const uint32_t sbase = smem_u32(stage0 + (size_t)s * G::STAGE_BYTES);
#pragma unroll 1
      for (int b = 0; b < G::NBLK; b++) {
        mbar_wait_sleep(&acc_empty[buf], par ^ 1u, 32);
        tc_fence_after();
        const uint32_t bblk = sbase + (uint32_t)b * G::BLOCK_BYTES;
        const uint32_t dbase = tmem + buf * G::ACC_COLS;
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < KB; kb++) {
#pragma unroll
            for (int j = 0; j < NS; j++) {  // Theta slice j times X slices 0 .. NS-1-j at once: diagonals j .. NS-1
              constexpr int NMAX = 256;      // an instruction's N; a wider operand (wide blocks, j = 0: 320 rows) goes in two halves
              const int nrows = G::BLK * (NS - j), nparts = nrows > NMAX ? 2 : 1, npart = nrows / nparts;
              const uint32_t accum = (j > 0 || kb > 0) ? 1u : 0u;
#pragma unroll
              for (int hh = 0; hh < nparts; hh++) {  // rows hh npart ..: 32 bytes per row in 8-row groups of 256 B; their products land npart columns further right
                const uint32_t idesc = IDESC0 | ((uint32_t)(npart >> 3) << 17);
                const uint64_t bdesc = tc_smem_desc(bblk + (uint32_t)(kb * NS) * G::SLAB_BYTES + (uint32_t)(hh * npart) * 32u, 128u, 256u);
                const uint32_t dcol = dbase + j * G::BLK + hh * npart;
                if (G::a_in_tmem(j, kb)) tc_mma_i8_ts(dcol, tmem + G::a_tmem_col(j, kb), bdesc, idesc, accum);
                else tc_mma_i8_ss(dcol, tc_smem_desc(sA_addr + (uint32_t)G::a_smem_unit(j, kb) * 4096u, 128u, 256u), bdesc, idesc, accum);
              }
            }
          }
          tc_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++buf == (uint32_t)NACC) { buf = 0; par ^= 1u; }
      }
      if (elect_one()) tc_commit(&empty[s]);  // the stage's slices are free once every MMA that reads them has retired
      __syncwarp();
    }
  } else {
    // ===== epilogue: thread = chain (TMEM lane); the EW / 4 warps of a lane quarter split the 32 columns =====
    const int q = warp_u & 3, h = warp_u >> 2;
    const uint32_t lane_base = ((uint32_t)(q * 32) << 16) + __reduce_min_sync(FM_FULL, tmem);
    if (FAMILY == FMCMC_FAMILY_LOGISTIC) mbar_wait(tab_bar, 0u);  // the softplus table has landed
    double acc = 0.0, acc2 = 0.0;
    long long it = 0;
    const uint32_t sp_tab_s = smem_u32(sp_tab);
    const uint32_t sp_tab_lane = sp_tab_s + (uint32_t)(lane & (FM_LC6_REP - 1)) * 16u;  // level 3: this lane's copy / bank group
    constexpr bool GRP = I8_GROUPED != 0 && EW == 16 && NACC >= 2;   // (one accumulator set: every warp visits every block)
    constexpr int CWG = GRP ? 2 * CW : CW;          // columns of a block owned by this warp
    const int grp = h & 1;                          // GRP: this warp's group; it visits every other block
    const int hcol = GRP ? (h >> 1) * CWG : h * CW;
    constexpr int BSTEP = GRP ? 2 : 1;              // blocks between two visits
    constexpr bool PHASES = I8_PHASE_SPLIT != 0 && FAMILY == FMCMC_FAMILY_GAUSSIAN_LM && KB == 4 && CWG == 16;
    // accumulator set and barrier phase of the next block this warp visits (block n: n % NACC, (n / NACC) & 1)
    uint32_t buf = GRP ? (uint32_t)grp : 0u, par = 0u;
    for (int sl = 0; sl < gsl; sl++) {
    for (long long tile = first + sl; tile < ntiles; tile += step, it++) {
      const double* ymeta = mp.y + tile * G::TO;  // L1-resident broadcast loads (Gaussian / non-binary logistic only)
      const int valid = (int)min((long long)G::TO, mp.n - tile * G::TO);  // < TO only for the last tile
      // the first block of this stage that belongs to the warp's group (blocks are numbered across stages)
      const int bfirst = GRP ? ((G::NBLK & 1) ? (int)((grp - (int)(it & 1)) & 1) : grp) : 0;
#pragma unroll 1
      for (int b = bfirst; b < G::NBLK; b += BSTEP) {
      {
        mbar_wait(&acc_full[buf], par);
        tc_fence_after();
        if constexpr (PHASES) {
          // Gaussian family at K = 128, where the tensor pipe is busy most of the time: the integer half of the epilogue -
          // tcgen05.ld, merge of the diagonals, int64 -> double - runs for ALL of the warp's columns first (it shares no
          // datapath with the tensor pipe, so it overlaps the MMAs of the other accumulator set), then the FP64 half, which
          // only gets the pipe while no MMA executes.  Interleaved, every evaluation stalls at its first FP64 instruction.
          double tv[CWG];
#pragma unroll
          for (int cc = 0; cc < CWG / CH; cc++) {
            uint32_t a[NS][CH];
            tc_ld_diagonals<NS, CH, G::BLK>(buf * G::ACC_COLS + hcol + cc * CH + lane_base, a);
            tc_wait_ld<NS, CH>(a);
            if (cc == CWG / CH - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&acc_empty[buf]);
            }
#pragma unroll
            for (int e = 0; e < CH; e++) tv[cc * CH + e] = i8_assemble<NS, CH, KB>(a, e, tb.tune);
          }
          i8_pin16(tv);   // nothing of the FP64 half is scheduled above this point
#pragma unroll
          for (int cc = 0; cc < CWG / CH; cc++) {
            const int obs0 = b * G::BLK + hcol + cc * CH;
            if (obs0 + CH <= valid) {
              double2 yy[CH / 2];
#pragma unroll
              for (int e = 0; e < CH / 2; e++) yy[e] = __ldg(reinterpret_cast<const double2*>(ymeta + obs0) + e);
#pragma unroll
              for (int e = 0; e < CH; e++) {
                const double r = ((e & 1) ? yy[e / 2].y : yy[e / 2].x) - fma(tv[cc * CH + e], csc, b0);
                acc = fma(r, r, acc);
              }
            } else {
#pragma unroll
              for (int e = 0; e < CH; e++)
                if (obs0 + e < valid) {
                  const double r = __ldg(ymeta + obs0 + e) - fma(tv[cc * CH + e], csc, b0);
                  acc = fma(r, r, acc);
                }
            }
          }
        } else
#ifdef I8_CC_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
        for (int cc = 0; cc < CWG / CH; cc++) {
          const int col0 = hcol + cc * CH;
          uint32_t a[NS][CH];
#ifdef FMCMC_I8_TUNE_HOOKS  // profiling experiments only (profiles/r01_i8_findings.md): results are garbage
          if (tb.tune & 1) {
#pragma unroll
            for (int d = 0; d < NS; d++)
#pragma unroll
              for (int e = 0; e < CH; e++) a[d][e] = (uint32_t)(lane * 37 + d * 11 + e + (int)buf);
          } else
#endif
          {
            tc_ld_diagonals<NS, CH, G::BLK>(buf * G::ACC_COLS + col0 + lane_base, a);   // diagonal d sits d BLK columns further right
          }
          tc_wait_ld<NS, CH>(a);
          if (cc == CWG / CH - 1) {  // this warp's share of the accumulator set is in registers: hand the buffer back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
          }
          const int obs0 = b * G::BLK + col0;
#ifdef FMCMC_I8_TUNE_HOOKS
          if (tb.tune & 2) {
#pragma unroll
            for (int e = 0; e < CH; e++) acc += (double)(int)(a[0][e] ^ a[NS - 1][e]);
          } else
#endif
          if (REP_ENTRIES > 0 && cta_rep && obs0 + CH <= valid) {
#pragma unroll
            for (int e = 0; e < CH; e++)  // the hot loop of cfg3
              i8_logistic_cubic_rep(i8_assemble<NS, CH, KB>(a, e, tb.tune), csc, acc2, sp_tab_lane);
          } else if (REP_ENTRIES > 0 && cta_rep) {
            for (int e = 0; e < CH; e++)  // last, partial tile of a CTA
              if (obs0 + e < valid) i8_logistic_cubic_rep(i8_assemble<NS, CH, KB>(a, e, tb.tune), csc, acc2, sp_tab_lane);
          } else if (FAMILY == FMCMC_FAMILY_LOGISTIC && YBIN && i8_table_level<NS, KB>() == 2 && cta_in_table && obs0 + CH <= valid) {
#pragma unroll
            for (int e = 0; e < CH; e++)  // |eta| beyond the replicated table: 256-per-unit, mean-corrected
              i8_logistic_lcosh3_fast(i8_assemble<NS, CH, KB>(a, e, tb.tune), csc, acc2, sp_tab_s);
          } else if (FAMILY == FMCMC_FAMILY_LOGISTIC && YBIN && i8_table_level<NS, KB>() == 2 && cta_in_table) {
            for (int e = 0; e < CH; e++)  // last, partial tile of a CTA on the mean-corrected table
              if (obs0 + e < valid) i8_logistic_lcosh3_fast(i8_assemble<NS, CH, KB>(a, e, tb.tune), csc, acc2, sp_tab_s);
          } else if (FAMILY == FMCMC_FAMILY_LOGISTIC && YBIN && i8_table_level<NS, KB>() == 2 && eta_in_table && obs0 + CH <= valid) {
#pragma unroll
            for (int e = 0; e < CH; e++)
              i8_logistic_lcosh_fast(i8_assemble<NS, CH, KB>(a, e, tb.tune), csc, acc2, sp_tab);
          } else if (FAMILY == FMCMC_FAMILY_LOGISTIC && YBIN && i8_table_level<NS, KB>() == 1 && eta_small && obs0 + CH <= valid) {
#pragma unroll
            for (int e = 0; e < CH; e++)
              i8_logistic_even_t<1, false>(i8_assemble<NS, CH, KB>(a, e, tb.tune), csc, hi_clamp, acc, acc2, sp_tab, tb.tune);
          } else if (FAMILY == FMCMC_FAMILY_GAUSSIAN_LM && obs0 + CH <= valid) {
            double2 yy[CH / 2];   // warp-uniform addresses: broadcast loads, two responses each (y is 16-byte aligned, obs0 a multiple of 8)
#pragma unroll
            for (int e = 0; e < CH / 2; e++) yy[e] = __ldg(reinterpret_cast<const double2*>(ymeta + obs0) + e);
#pragma unroll
            for (int e = 0; e < CH; e++) {
              const double r = ((e & 1) ? yy[e / 2].y : yy[e / 2].x) - fma(i8_assemble<NS, CH, KB>(a, e, tb.tune), csc, b0);
              acc = fma(r, r, acc);
            }
          } else if (obs0 + CH <= valid) {
#pragma unroll
            for (int e = 0; e < CH; e++) {
              const double t = i8_assemble<NS, CH, KB>(a, e, tb.tune);
              if (FAMILY == FMCMC_FAMILY_GAUSSIAN_LM) {
                const double r = __ldg(ymeta + obs0 + e) - fma(t, csc, b0);  // warp-uniform address: broadcast
                acc = fma(r, r, acc);
              } else if (YBIN) {
                i8_logistic_safe<i8_table_level<NS, KB>()>(t, csc, hi_clamp, acc, acc2, sp_tab);
              } else {
                acc += i8_logistic_term<i8_table_level<NS, KB>()>(t * csc, __ldg(ymeta + obs0 + e), sp_tab);
              }
            }
          } else {
#pragma unroll
            for (int e = 0; e < CH; e++) {
              if (obs0 + e < valid) {
                const double t = i8_assemble<NS, CH, KB>(a, e, tb.tune);
                if (FAMILY == FMCMC_FAMILY_GAUSSIAN_LM) {
                  const double r = __ldg(ymeta + obs0 + e) - fma(t, csc, b0);
                  acc = fma(r, r, acc);
                } else if (YBIN) {
                  i8_logistic_safe<i8_table_level<NS, KB>()>(t, csc, hi_clamp, acc, acc2, sp_tab);
                } else {
                  acc += i8_logistic_term<i8_table_level<NS, KB>()>(t * csc, __ldg(ymeta + obs0 + e), sp_tab);
                }
              }
            }
          }
        }
        buf += BSTEP;
        if (buf >= (uint32_t)NACC) { buf -= (uint32_t)NACC; par ^= 1u; }
      }
      }
    }
    // ---- this slice is complete: the EW / 4 warps of a lane quarter meet in shared memory (epilogue warps only: the producer
    // and the MMA warp are already working on the next slice), one fixed-order sum per chain, one partial sum per (slice, chain) ----
    // binary logistic: acc holds sum |t|; sum(|eta| / 2 + g) = (csc / 2) sum |t| + sum g
    red[h * I8_CHAINS + q * 32 + lane] = (FAMILY == FMCMC_FAMILY_LOGISTIC && YBIN) ? fma(0.5 * csc, acc, acc2) : acc + acc2;
    asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
    if (warp_u < 4 && col < tb.ncols) {
      double v = red[tid];
#pragma unroll
      for (int hh = 1; hh < EW / 4; hh++) v += red[hh * I8_CHAINS + tid];  // fixed order
      if (FAMILY == FMCMC_FAMILY_LOGISTIC && YBIN) v = (slice0 + sl == 0 ? lin : 0.0) - v;  // theta . sxy enters once per chain
      // non-finite parameters never reach the integer path: NaN propagates (the reference's `undefined` abort),
      // +-Inf / beyond 2^480 gives the rejected-proposal value
      if (th_nan) v = NAN;
      else if (th_big) v = (FAMILY == FMCMC_FAMILY_LOGISTIC) ? -INFINITY : INFINITY;
      tb.partial[(size_t)(slice0 + sl) * tb.ncols + col] = v;
    }
    if (sl + 1 < gsl) asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");   // red is written again at the end of the next slice
    acc = 0.0;
    acc2 = 0.0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
