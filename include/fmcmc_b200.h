/*
 * fmcmc_b200.h — C ABI of libfmcmcb200.so, the B200-native drop-in for the
 * multi-chain Metropolis-Hastings hot path of the R package fmcmc (v0.6-0).
 *
 * Everything here is plain C: pointers, sizes, PODs.  No torch / CUDA types
 * cross this boundary.  The reference-side binding (R `.Call` shim, or the
 * Python ctypes mirror shipped in fmcmc_b200/) is shown in INTEGRATION.md.
 *
 * Reference interfaces replaced (file:line in USCbiostats/fmcmc):
 *   fmcmc_run            R/mcmc.R:485-838 (MCMC_without_conv_checker: chain
 *                        fan-out 643-673, MH loop 726-783, burnin/thin 786-813)
 *   fmcmc_kernel_spec    R/kernel.R:283-317 (kernel_new env), hyper-parameters of
 *                        R/kernel_normal.R:26-31,96-103  R/kernel_unif.R:15-20,74-81
 *                        R/kernel_adapt.R:54-66  R/kernel_ram.R:65-79
 *                        R/kernel_mirror.R:54-64,177-187
 *   per-chain state      the mutable fields of those environments (abs_iter,
 *                        Sigma, Mean_t_prev, mu, scale, obs_arate, nerrors) that
 *                        R/mcmc.R:629-631 ships back from PSOCK workers
 *   fmcmc_model_*        the user closure `fun` (README.md:128-139,356-360;
 *                        vignettes/workflow-with-fmcmc.Rmd:35-41;
 *                        vignettes/advanced-features.Rmd:46-53;
 *                        playground/hierarchical-bayes.Rmd:45-51)
 *   fmcmc_store_* / fmcmc_gelman*   R/mcmc.R:947-968 (append_chains + checker
 *                        call) and R/convergence.R:191-246 -> coda::gelman.diag
 *   fmcmc_cov_recursive / fmcmc_mean_recursive / fmcmc_reflect
 *                        R/recursive.R:63-139, R/kernel.R:450-493 (exported R fns)
 *
 * Conventions
 *   - every entry point returns an int status (0 = FMCMC_OK) and, where it takes
 *     `err`/`errlen`, writes a NUL-terminated message on failure;
 *   - the caller owns every host buffer it passes; the library never frees it;
 *   - device memory is owned by the opaque handles;
 *   - all calls for one handle must come from one host thread (the R main thread).
 */
#ifndef FMCMC_B200_H
#define FMCMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FMCMC_ABI_VERSION 3

/* ---- status codes -------------------------------------------------------- */
enum {
  FMCMC_OK        = 0,
  FMCMC_EINVAL    = 1,  /* bad argument (message says which; same substrings as the R errors) */
  FMCMC_ENAN      = 2,  /* fun(par) undefined at some step (R/mcmc.R:758-765)  */
  FMCMC_ECUDA     = 3,  /* CUDA runtime error                                  */
  FMCMC_ENOMEM    = 4,
  FMCMC_ENOTPD    = 5,  /* kernel_adapt: Sigma not positive definite (MASS::mvrnorm stop) */
  FMCMC_EUNSUP    = 6,  /* configuration the reference itself mishandles or not built */
  FMCMC_ENANRATIO = 7,  /* f1 - f0 is NaN, R: "missing value where TRUE/FALSE needed" (D10) */
  FMCMC_EPEER     = 8   /* observation sharding: a peer GPU did not publish its partial sums in time */
};

/* ---- log-posterior families (the device-side replacement of `fun`) ------- */
enum {
  FMCMC_FAMILY_GAUSSIAN_LM = 1, /* sum dnorm(y - (b0 + X b), sd = theta[k-1], log=TRUE)   */
  FMCMC_FAMILY_LOGISTIC    = 2, /* Bernoulli-logit log-lik  - sum(beta^2)/(2 prior_sd^2)  */
  FMCMC_FAMILY_HIER_NORMAL = 3  /* y_i~N(th_g(i),s), th_g~N(gamma,tau), gamma~U(lo,hi)    */
};

/* model flags */
#define FMCMC_MODEL_INTERCEPT 1u /* GAUSSIAN_LM: theta[0] is an intercept that has no column in X */
#define FMCMC_MODEL_GUARD     2u /* GAUSSIAN_LM: non-finite sum -> -Inf (README.md:135-136)        */
#define FMCMC_MODEL_SCALES    4u /* HIER_NORMAL: theta also carries (sigma, tau) after gamma       */

typedef struct fmcmc_model_desc {
  int32_t  family;
  uint32_t flags;
  int64_t  n;         /* observations                                   */
  int32_t  p_x;       /* columns of X (0 for HIER_NORMAL)               */
  int32_t  n_groups;  /* HIER_NORMAL only                               */
  const double*  X;   /* n x p_x, column-major (R layout); host pointer */
  const double*  y;   /* n                                              */
  const int32_t* group; /* n, 0-based group of each observation (HIER_NORMAL) */
  double   hyper[4];  /* LOGISTIC: [0]=prior sd (2 in the vignette);
                         HIER_NORMAL: [0]=gamma lower, [1]=gamma upper       */
} fmcmc_model_desc;

/* Number of parameters k the family expects (theta length).  <0 on bad desc. */
int32_t fmcmc_model_nparams(const fmcmc_model_desc* desc);

/* ---- transition kernels -------------------------------------------------- */
enum {
  FMCMC_KERNEL_NORMAL            = 1, /* R/kernel_normal.R:26-82   */
  FMCMC_KERNEL_NORMAL_REFLECTIVE = 2, /* R/kernel_normal.R:96-177  */
  FMCMC_KERNEL_UNIF              = 3, /* R/kernel_unif.R:15-66     */
  FMCMC_KERNEL_UNIF_REFLECTIVE   = 4, /* R/kernel_unif.R:74-147    */
  FMCMC_KERNEL_ADAPT             = 5, /* R/kernel_adapt.R:54-208   */
  FMCMC_KERNEL_RAM               = 6, /* R/kernel_ram.R:65-181     */
  FMCMC_KERNEL_NMIRROR           = 7, /* R/kernel_mirror.R:54-173  */
  FMCMC_KERNEL_UMIRROR           = 8  /* R/kernel_mirror.R:177-301 */
};

/* update schemes of plan_update_sequence (R/kernel.R:66-133) */
enum {
  FMCMC_SCHEME_JOINT    = 0,
  FMCMC_SCHEME_ORDERED  = 1,
  FMCMC_SCHEME_RANDOM   = 2,
  FMCMC_SCHEME_EXPLICIT = 3
};

/* how kernel_adapt turns Sigma into a draw */
enum {
  FMCMC_MVN_CHOLESKY = 0, /* mu + L z  (production default: same N(mu, Sigma), O(k^3 / 3) per re-adapted row)   */
  FMCMC_MVN_EIGEN    = 1  /* mu + V sqrt(max(ev,0)) z, MASS::mvrnorm's own form (R/kernel_adapt.R:173-178) with
                             R's eigenvalue ordering; the eigenvector SIGNS are a LAPACK-build artefact, the
                             convention is "largest |component| positive" (DESIGN.md section 5).  Device and
                             oracle agree bit for bit; costs a Jacobi sweep set per re-adapted row. */
};

typedef struct fmcmc_kernel_spec {
  int32_t type;
  int32_t k;            /* length of theta                                        */
  int32_t scheme;       /* NORMAL/UNIF/MIRROR families only                       */
  int32_t order_len;    /* EXPLICIT scheme: length of `order` (== #free params)   */
  const int32_t* order; /* EXPLICIT scheme: 1-based parameter indices             */
  const int32_t* seq;   /* RANDOM scheme, fed mode: [nchains][seq_len] 1-based active
                           coordinate of each row (R/kernel.R:109-112); NULL => Philox */
  int64_t seq_len;      /* rows planned (the first run's nsteps, quirk D7)        */
  /* All vectors below have length k and are already recycled
     (check_dimensions, R/kernel.R:1-15) and NA-processed (process_bounds 25-41). */
  const double*  mu;    /* NORMAL*: proposal mean; ADAPT: mu (indexed by free params
                           inside); MIRROR: initial centre (used when state is fresh)    */
  const double*  scale; /* NORMAL*: sd; MIRROR: initial scale                      */
  const double*  min_;  /* UNIF*                                                   */
  const double*  max_;  /* UNIF*                                                   */
  const double*  lb;    /* reflective / adaptive / mirror kernels; +-DBL_MAX = none */
  const double*  ub;
  const uint8_t* fixed; /* 1 = parameter never moves                               */
  /* adaptive hyper-parameters */
  int64_t warmup;       /* ADAPT 500, RAM 0, MIRROR 500                            */
  int64_t freq;         /* ADAPT/RAM 1                                             */
  int64_t bw;           /* ADAPT windowed covariance (0 = recursive)               */
  double  until;        /* ADAPT/RAM: stop adapting at this abs_iter (Inf = never) */
  double  eps;          /* ADAPT/RAM 1e-4                                          */
  double  Sd;           /* ADAPT: 5.76/k_free when <= 0 (only used when bw > 0)    */
  double  arate;        /* RAM .234, MIRROR .4                                     */
  const int64_t* nadapt;/* MIRROR: checkpoint schedule (R/kernel_mirror.R:165)     */
  int32_t nadapt_len;
  int32_t mvn_method;   /* ADAPT: FMCMC_MVN_*                                      */
  const double* constr; /* RAM: k x k column-major mask or NULL (R/kernel_ram.R:149-150) */
} fmcmc_kernel_spec;

/*
 * Per-chain mutable kernel state, round-tripped on every fmcmc_run so that a
 * kernel object can be reused across bulks / calls exactly like the R env
 * (SURVEY §5 "resume by value").  Flat layout, per chain c:
 *   istate[c*FMCMC_ISTATE_LEN + 0] abs_iter
 *   istate[c*FMCMC_ISTATE_LEN + 1] flags  (bit0: ADAPT Mean_t_prev is set;
 *                                          bit1: state initialised (k known);
 *                                          bits 2-3: MIRROR obs_arate 0=NULL 1=scalar 2=vector)
 *   istate[c*FMCMC_ISTATE_LEN + 2] nerrors (RAM)
 *   istate[c*FMCMC_ISTATE_LEN + 3] n_changed: rows of the current run whose state differs
 *                                  from the previous row (scratch, reset every run)
 *   dstate[c*dlen ...]  ADAPT : Sigma[kf*kf] (col-major), Mean_t_prev[kf]
 *                       RAM   : Sigma[kf*kf] (used as the factor S)
 *                       MIRROR: mu[k], scale[k], obs_arate[k]
 *                       others: empty (dlen = 0)
 * where kf = number of non-fixed parameters, dlen = fmcmc_kernel_state_len().
 * A fresh state is all zeros (flags bit1 clear): the run initialises it the way
 * the R closures do on their first proposal call.
 */
#define FMCMC_ISTATE_LEN 4
#define FMCMC_STATE_HAS_MEAN   1
#define FMCMC_STATE_INIT       2
#define FMCMC_STATE_OBS_SHIFT  2

int64_t fmcmc_kernel_state_len(int32_t type, int32_t k, int32_t k_free);

typedef struct fmcmc_kernel_state {
  int64_t* istate;  /* [nchains][FMCMC_ISTATE_LEN] */
  double*  dstate;  /* [nchains][dlen]             */
} fmcmc_kernel_state;

/* ---- random streams ------------------------------------------------------ */
enum {
  FMCMC_STREAM_PHILOX = 0, /* production: Philox4x32-10 keyed by (seed, global chain id),
                              counter (run_index, row, slot): independent of #GPUs       */
  FMCMC_STREAM_FED    = 1  /* verification: host uploads R's own draws   */
};

typedef struct fmcmc_stream_spec {
  int32_t  mode;
  int32_t  kdraw;      /* FED: slots per row in `z` (k_free for joint, 1 otherwise)     */
  uint64_t seed;       /* PHILOX                                                        */
  uint64_t run_index;  /* PHILOX: bulk counter, so consecutive bulks use fresh counters */
  const double* logu;  /* FED: [nchains][nsteps]   log(runif(nsteps)); [.,0] is unused  */
  const double* z;     /* FED: [nchains][nsteps][kdraw]; row 0 unused.  Standard normals
                          (NORMAL*, ADAPT, NMIRROR), U(0,1) (UNIF*, UMIRROR) or U=qfun(k)
                          (RAM).  The device applies mu + scale*z / a + (b-a)*u unfused. */
} fmcmc_stream_spec;

/* ---- one call of MCMC_without_conv_checker ------------------------------- */
#define FMCMC_RUN_NO_DRAWS   1u /* do not copy `draws` back (pointer may be NULL)       */
#define FMCMC_RUN_COLMAJOR   2u /* outputs as R matrices: [chain][param][row]           */
#define FMCMC_RUN_APPEND     4u /* append the kept rows to the handle's sample store
                                   (append_chains, R/mcmc.R:947) for fmcmc_gelman       */
#define FMCMC_RUN_NO_OUTPUT  8u /* keep results on the device only (store / last state) */
#define FMCMC_RUN_DEVICE_STATE 16u /* kernel state stays resident in HBM: `state` is neither
                                   uploaded nor downloaded; the state the previous fmcmc_run left
                                   on this model is used (bulk loop, R/mcmc.R:901-947)       */

#define FMCMC_RUN_KEEP_STATE 32u /* upload `state` as usual but do not download it afterwards: the first bulk of a loop whose
                                   later bulks run with FMCMC_RUN_DEVICE_STATE (fmcmc_kernel_state_fetch at the end)   */

typedef struct fmcmc_run_spec {
  int64_t nsteps;        /* rows, including the initial state (R semantics)  */
  int64_t burnin;
  int64_t thin;
  int32_t nchains;       /* chains handled by THIS call / GPU                */
  uint32_t flags;
  int64_t chain_offset;  /* global id of local chain 0 (Philox keying)       */
  const double* initial; /* [nchains][k] row-major; NULL => continue from the
                            last state kept on the device (next bulk)        */
  int64_t nchains_total; /* chains of the whole job over all GPUs (0 => nchains).  The stepping path is chosen
                            from THIS count, so a chain's log-posterior bits do not depend on how many
                            GPUs the job was sharded over                                              */
  int64_t out_rows_total; /* 0: ans / draws / logpost hold exactly rows_kept rows per chain.  > 0: they hold this many
                            rows per chain ([nchains][out_rows_total][k], or [nchains][k][out_rows_total] with
                            FMCMC_RUN_COLMAJOR) and this call writes rows out_row_offset .. out_row_offset + rows_kept:
                            a bulk loop fills ONE set of arrays bulk after bulk (append_chains without the copies)  */
  int64_t out_row_offset;
} fmcmc_run_spec;

typedef struct fmcmc_run_report {
  int64_t rows_kept;     /* T' = rows after burnin/thin                      */
  int64_t first_iter;    /* iteration number of the first kept row (mcpar)   */
  int64_t last_iter;
  int64_t nan_chain;     /* FMCMC_ENAN: local chain / row (1-based, R's i)   */
  int64_t nan_step;
  int64_t n_accept;      /* accepted proposals over all chains               */
  int64_t n_launches;    /* CUDA kernels launched by this call               */
  double  device_ms;     /* CUDA-event time of the stepping kernels          */
  int32_t path;          /* 1 = chain-resident fused kernel, 2 = observation-tiled (DFMA), 3 = (DMMA), 4 = (tcgen05 int8 slices) */
  int32_t reserved;
  double  hot_ms;        /* summed CUDA-event time of the dominant kernel's launches ... */
  int64_t hot_launches;  /* ... and how many of them were timed: the first and every 16th likelihood launch of a call
                            (a timed launch is bracketed by event records and therefore does not overlap its set-up with
                            the preceding head kernel - programmatic dependent launch -, the others do)               */
  int64_t h2d_bytes;     /* bytes this call copied host -> device                        */
  int64_t d2h_bytes;     /* bytes this call copied device -> host                        */
} fmcmc_run_report;

typedef struct fmcmc_model fmcmc_model;   /* opaque: device-resident X, y, buffers */

int fmcmc_version(void);
int fmcmc_device_count(void);

/* Upload X/y once (the reference re-ships them to PSOCK workers every bulk,
 * R/mcmc.R:550-577).  `device` is the CUDA ordinal (LOCAL_RANK). */
int fmcmc_model_create(const fmcmc_model_desc* desc, int device,
                       fmcmc_model** out, char* err, size_t errlen);
/* Same, but X/y/group are DEVICE pointers already resident in HBM (borrowed,
 * must outlive the model).  Used by the bench's HBM-resident leg. */
int fmcmc_model_create_device(const fmcmc_model_desc* desc_with_device_ptrs, int device,
                              fmcmc_model** out, char* err, size_t errlen);
void fmcmc_model_free(fmcmc_model* m);
/* Device-clock stopwatch on the library's launch stream: slots 0..7.  A timed region that spans several calls
 * (stepping + R-hat statistics + host finish) is bracketed by two marks (bench.py). */
int fmcmc_event_mark(fmcmc_model* m, int slot);
int fmcmc_event_elapsed_ms(fmcmc_model* m, int from_slot, int to_slot, double* ms);
/* Frees the device copies of X the given stepping path does not read (path 4 reads only its int8 slice tiles,
 * path 3 its tile-major FP64 copy): at BASELINE configs[4] that returns 20.7 of 28.4 GB per GPU.  Afterwards
 * the model runs that path only (anything else -> FMCMC_EUNSUP); borrowed device pointers are never freed. */
int fmcmc_model_trim(fmcmc_model* m, int path, char* err, size_t errlen);

/*
 * Run `nsteps` rows for `nchains` chains.  Outputs (host, caller-allocated, may
 * be NULL with FMCMC_RUN_NO_OUTPUT):
 *   ans     [nchains][rows_kept][k]  accepted states   (R/mcmc.R:778)
 *   draws   [nchains][rows_kept][k]  proposals         (R/mcmc.R:752)
 *   logpost [nchains][rows_kept]     f(proposal), quirk D1 (R/mcmc.R:754)
 * rows_kept = fmcmc_rows_kept(nsteps, burnin, thin).
 */
int fmcmc_run(fmcmc_model* m, const fmcmc_run_spec* run, const fmcmc_kernel_spec* kernel,
              fmcmc_kernel_state* state, const fmcmc_stream_spec* stream,
              double* ans, double* draws, double* logpost,
              fmcmc_run_report* report, char* err, size_t errlen);

int64_t fmcmc_rows_kept(int64_t nsteps, int64_t burnin, int64_t thin);

/* Downloads the kernel state the last fmcmc_run left on the device (shapes of that run).  A bulk loop runs its bulks
 * after the first with FMCMC_RUN_DEVICE_STATE - no state traffic at all - and calls this once at the end for the
 * write-back into the kernel objects (R/mcmc.R:629-631). */
int fmcmc_kernel_state_fetch(fmcmc_model* m, fmcmc_kernel_state* state, char* err, size_t errlen);

/* Force a stepping path (0 = auto, 1 = chain-resident, 2 = observation-tiled DFMA kernel,
 * 3 = observation-tiled FP64 tensor-core (DMMA) kernel, 4 = observation-tiled split-integer kernel: X.Theta
 * exact on int8 slices on the tcgen05 tensor cores, FP64 pipe for the log-density only).  Auto picks 1 for data
 * that fits on chip, 2 for p_x <= 16, 4 for more than 128 likelihood columns (chains), else 3; 4 falls back to 3
 * when X holds non-finite values.  No equivalent in the reference (one CPU code path). */
int fmcmc_set_path(fmcmc_model* m, int path);

/* Evaluate the family's log-posterior for `nchains` parameter vectors
 * ([nchains][k] row-major) on the device: the `f(theta)` of R/mcmc.R:742. */
int fmcmc_logpost(fmcmc_model* m, int32_t nchains, const double* theta, double* out,
                  char* err, size_t errlen);

/* ---- sample store + Gelman-Rubin (R/convergence.R:191-246) -------------- */
/* Reserve room for `capacity_rows` appended rows of nchains x k. */
int fmcmc_store_reset(fmcmc_model* m, int32_t nchains, int32_t k, int64_t capacity_rows,
                      char* err, size_t errlen);
int64_t fmcmc_store_rows(const fmcmc_model* m);

/*
 * Per-GPU partial statistics of rows [row_begin,row_end) of the store, free
 * parameters only (free_mask[k], 1 = keep):
 *   xbar   [nchains][kf]   per-chain means
 *   s2     [nchains][kf]   per-chain variances (divisor N-1)
 *   wsum   [kf*kf]         sum over local chains of the chain covariance matrices
 * Outputs are DEVICE pointers when dev_out != 0 (so torch.distributed can
 * all_gather / all_reduce them over NCCL), host pointers otherwise.
 */
int fmcmc_gelman_partials(fmcmc_model* m, int64_t row_begin, int64_t row_end,
                          const uint8_t* free_mask, double* xbar, double* s2, double* wsum,
                          int dev_out, char* err, size_t errlen);

/*
 * Finish coda::gelman.diag from (gathered) statistics of `nchains_total`
 * chains and `niter` rows: psrf point estimates [kf] and mpsrf (NaN if kf==1).
 * Returns FMCMC_ENOTPD when chol(W) fails (the R wrapper turns that into a
 * warning + FALSE, R/convergence.R:207-217).  Inputs are DEVICE pointers when
 * dev_in != 0.
 */
int fmcmc_gelman_finish(fmcmc_model* m, int64_t niter, int64_t nchains_total, int32_t kf,
                        const double* xbar, const double* s2, const double* wsum, int dev_in,
                        double* psrf, double* mpsrf, char* err, size_t errlen);

/* Single-GPU convenience = what `conv_checker(ans)` costs the R glue in ONE call (R/mcmc.R:968 ->
 * R/convergence.R:207 coda::gelman.diag): coda's autoburnin window + partials + finish on everything in the
 * store.  `start_iter` / `thin` are the store's mcpar (first kept iteration, thinning); the window is
 * fmcmc_gelman_window_begin(start_iter, thin, rows) .. rows, i.e. coda only drops the first half when
 * start < end/2 and snaps to the next kept iteration.  Fewer than 2 rows in the window -> FMCMC_EINVAL. */
int fmcmc_gelman(fmcmc_model* m, const uint8_t* free_mask, int64_t start_iter, int64_t thin, double* psrf,
                 double* mpsrf, int64_t* niter_used, char* err, size_t errlen);
int64_t fmcmc_gelman_window_begin(int64_t start_iter, int64_t thin, int64_t rows);

/* rm_invariant (R/convergence.R:169-186, quirk D9): the reference tests ONE pooled number, sd(rbind(all chains))^2 over every
 * accumulated row and every (free) parameter, against 1e-10.  out[3] = (count, mean, M2 = sum (x - mean)^2) of this GPU's part
 * of the store; several GPUs combine the triples (Chan et al.) - fmcmc_b200/dist.py does. */
int fmcmc_store_pooled(fmcmc_model* m, const uint8_t* free_mask, double* out, char* err, size_t errlen);

/* Effective sample size of every (local chain, free parameter) series over rows [row_begin, row_end) of the store, on the
 * device: autocovariances up to max_lag (<= 0: min(rows - 1, 2000)) and Geyer's initial positive sequence,
 * ESS = N / (-1 + 2 sum_m [rho(2m) + rho(2m + 1)]) up to the first non-positive pair.  ess: host [nchains][kf].
 * *truncated = 1 when some series still had positive pairs at max_lag (its ESS is then an over-estimate).  The reference
 * has no ESS of its own (README.md:191-194 prints coda's time-series SE; coda::effectiveSize is third-party): this is the
 * "ESS/sec" of BASELINE.json's metric and the MCSE of the distributional checks. */
int fmcmc_store_ess(fmcmc_model* m, int64_t row_begin, int64_t row_end, const uint8_t* free_mask, int32_t max_lag,
                    double* ess, int32_t* truncated, char* err, size_t errlen);

/* Host-only helper of the Gelman finish, exported for the CPU test-suite: largest eigenvalue of a symmetric
 * p x p matrix (col-major; Householder tridiagonalisation + Sturm bisection). */
int fmcmc_host_sym_eigmax(int32_t p, const double* A, double* emax);

/* ---- observation sharding across GPUs (few chains, huge n) -------------------------
 * Every rank creates its model on a ROW SLICE of X / y and runs the SAME chains with the same streams.  The
 * likelihood kernel of each rank stores its per-slice partial sums straight into every rank's exchange buffer
 * (peer stores over NVLink) and raises a per-step flag; each rank's head kernel waits for all flags and reduces
 * all ranks' partials in one fixed order, so every rank takes bit-identical accept/reject decisions with no
 * host-side collective on the step path.  This replaces nothing in the reference (its chains are never split
 * over workers, R/mcmc.R:593-627): it is the multi-GPU mode for the reference's typical few-chain usage.
 *   1. every rank:  fmcmc_shard_alloc()   -> its exchange buffer + CUDA IPC handles
 *   2. all_gather the handles (torch.distributed / MPI / anything)
 *   3. every rank:  fmcmc_shard_attach()  -> opens the peers' buffers; from now on fmcmc_run() is collective:
 *      all ranks must call it with the same run / kernel / stream arguments, within 60 s of each other
 *      (a rank that waits longer for a peer's partial sums gives up with FMCMC_EPEER).                    */
typedef struct fmcmc_shard_handles {
  unsigned char partial[64]; /* cudaIpcMemHandle_t of the partial-sum exchange buffer */
  unsigned char flags[64];   /* cudaIpcMemHandle_t of the step flags                  */
  void* partial_ptr;         /* the same two buffers as raw device pointers, for peers living in  */
  void* flags_ptr;           /* the SAME process (IPC handles cannot be opened by their creator)   */
  int32_t device;            /* CUDA ordinal owning them                                            */
  int32_t pid;               /* creator process id                                                  */
} fmcmc_shard_handles;

/* max_cols: the largest number of likelihood columns a run will use (nchains, x2 for kernel_ram);
 * n_total: observations over all ranks (the Gaussian families' normalising term needs it). */
int fmcmc_shard_alloc(fmcmc_model* m, int world, int max_cols, int64_t n_total, fmcmc_shard_handles* out,
                      char* err, size_t errlen);
/* all[world]: every rank's handles in rank order (all[rank] must be this model's own). */
int fmcmc_shard_attach(fmcmc_model* m, int rank, int world, const fmcmc_shard_handles* all,
                       char* err, size_t errlen);

/* ---- exported helpers of the reference, device implementations ---------- */
/* R/recursive.R:124-139 / 63-120 applied to `rows` consecutive rows of X
 * ([rows][k] row-major), starting from (mean_prev, cov_prev, t).  Outputs the
 * last mean [k] and covariance [k*k]. */
int fmcmc_cov_recursive(int device, int32_t k, int64_t rows, const double* X,
                        const double* mean_prev, const double* cov_prev, double t,
                        double eps, double Sd, const double* Ik,
                        double* mean_out, double* cov_out, char* err, size_t errlen);
/* R/kernel.R:450-493 on `count` vectors of length k ([count][k]). */
int fmcmc_reflect(int device, int32_t k, int64_t count, double* x, const double* lb,
                  const double* ub, const uint8_t* which, char* err, size_t errlen);

/* ---- measurement helper (bench.py): FP64 FMA peak of the device ---------- */
/* Times a register-resident DFMA loop on every SM; returns TFLOP/s (2 flop per FMA).
 * MEASURED_PEAKS.json has no FP64 figure and the hot path is FP64-pipe bound. */
int fmcmc_measure_fp64_peak(int device, double* dfma_tflops, char* err, size_t errlen);

/* Self-test hook: out[i] = log(1 + exp(-a[i])) evaluated by the device routine the logistic
 * epilogue uses (csrc/softplus.h), so tests can bound its error against mpmath. */
int fmcmc_test_softplus(int device, int64_t n, const double* a, double* out, char* err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif /* FMCMC_B200_H */
