// fmcmc_b200.cu — C ABI (include/fmcmc_b200.h) + host orchestration.
// Single translation unit; device code lives in the .cuh files next to it.
#include <cuda_runtime.h>
#include <unistd.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "../../include/fmcmc_b200.h"
#include "common.cuh"
#include "families.cuh"
#include "gelman.cuh"
#include "propose.cuh"
#include "resident.cuh"
#include "tiled.cuh"
#include "tiled_mma.cuh"
#include "tiled_i8.cuh"

// --------------------------------------------------------------------------------
// small host utilities
// --------------------------------------------------------------------------------
static void set_err(char* err, size_t errlen, const char* fmt, ...) {
  if (!err || !errlen) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err, errlen, fmt, ap);
  va_end(ap);
}

#define CU_CHECK(call)                                                                        \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      set_err(err, errlen, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__,   \
              __LINE__, #call);                                                               \
      return FMCMC_ECUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define FM_HOT_EVENTS 64
#define FM_HOT_EVERY 16   // one likelihood launch in 16 is bracketed by CUDA events (and runs without launch overlap)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

static cudaError_t ensure(DevBuf& b, size_t bytes) {
  if (bytes == 0) bytes = 16;
  if (bytes <= b.cap) return cudaSuccess;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  cudaError_t e = cudaMalloc(&b.p, bytes);
  if (e == cudaSuccess) b.cap = bytes;
  return e;
}
static void release(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

struct fmcmc_model {
  int device = 0;
  ModelParams mp{};
  bool borrowed = false;
  DevBuf X, y, group, sp_tab, sp_tab4, sp_tab8, sp_tab8m, sp_tab6r, Xt, Xq, xq_bad, xq_aux;
  int xt_PB = 0;          // padded width the tile-major copy Xt was built for (0 = not built)
  int xq_NS = 0, xq_KB = 0;  // slices / 32-column blocks the int8 tile copy Xq was built for (0 = not built; -1 = X not sliceable)
  int i8_slices = 0;      // int8 slices per operand of path 4: 0 = automatic (5 8-bit digits; 6 for kernel_ram and for n < 65536)
                          // (FMCMC_I8_SLICES = 6 | 7 overrides; tiled_i8.cuh has the error bound)
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;        // output rows leave the device while later rows are still being computed
  std::vector<cudaEvent_t> chunk_ev;         // "kept rows of chunk q are final" (recorded on `stream`)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t marks[8] = {};                 // fmcmc_event_mark
  int sm_count = 148;
  int smem_optin = 0;
  int forced_path = 0;
  int trimmed_to = 0;     // fmcmc_model_trim: the only stepping path whose copy of X is still resident (0 = all)
  int tiled_default = 3;  // tiled variant picked when p_x <= 32 (FMCMC_TILED_VARIANT=2|3 overrides; tuning only)
  int tiled_many = 4;     // tiled variant for > 128 likelihood columns (FMCMC_TILED_MANY=3|4 overrides; tuning only)
  int l2_bytes = 0;       // L2 cache size of the device
  int l2_keep_mb = -1;    // path 3, few chains: MB of X kept L2-resident across rows (-1: a third of L2; FMCMC_L2_KEEP_MB)
  int i8_gsl = 0;         // path 4: observation slices per CTA (0: automatic, up to 4; FMCMC_I8_GSL=1|2|4 overrides; A/B measurements)
  bool head_cta = true;   // few chains + kernel_adapt: one CTA per chain in the head kernel (FMCMC_HEAD_CTA=0: the warp-per-chain head; A/B measurements)
  bool pdl = true;        // programmatic dependent launch between the two kernels of an MH row (FMCMC_PDL=0 disables; A/B measurements)
  int mma_wide = 0;       // DMMA tile-shape variant (mma_shape(); FMCMC_MMA_VARIANT, tuning only)
  // run buffers (grow-only)
  DevBuf ans, draws, logpost, cur_theta, cur_f, prop, prop_u, istate, dstate, colsum, ubuf, work, cflags,
      errbuf, nacc, spec, fed_logu, fed_z, initial, partial, out_ans, out_draws, out_lp, tmp;
  int state_nchains = 0, state_k = 0, state_type = 0, state_kf = 0;  // shape of cur_theta / kernel state (valid after a run)
  std::vector<cudaEvent_t> hot_ev;
  // sample store (append_chains): [rows][C][k]
  DevBuf store;
  int store_C = 0, store_k = 0;
  long long store_cap = 0, store_rows = 0;
  // gelman scratch
  DevBuf g_xbar, g_s2, g_wsum, g_wpart, g_mask;
  // observation sharding (fmcmc_shard_*): exchange buffers are plain cudaMalloc so they can be IPC-exported
  int shard_world = 0, shard_rank = 0, shard_max_cols = 0;
  double* shard_partial = nullptr;               // [2][world * sm_count][max_cols]
  unsigned long long* shard_flags = nullptr;     // [world]
  unsigned int* shard_done = nullptr;
  double* shard_peer_partial[FM_MAX_PEERS] = {};
  unsigned long long* shard_peer_flags[FM_MAX_PEERS] = {};
  bool shard_ipc_opened[FM_MAX_PEERS] = {};
  unsigned long long shard_step = 0;
};

static int count_free(const fmcmc_kernel_spec* ks, std::vector<int>& free_idx) {
  free_idx.clear();
  for (int j = 0; j < ks->k; j++)
    if (!(ks->fixed && ks->fixed[j])) free_idx.push_back(j);
  return (int)free_idx.size();
}

// --------------------------------------------------------------------------------
// plain accessors
// --------------------------------------------------------------------------------
extern "C" int fmcmc_version(void) { return FMCMC_ABI_VERSION; }

extern "C" int fmcmc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" int32_t fmcmc_model_nparams(const fmcmc_model_desc* d) {
  if (!d) return -1;
  switch (d->family) {
    case FMCMC_FAMILY_GAUSSIAN_LM: return d->p_x + ((d->flags & FMCMC_MODEL_INTERCEPT) ? 1 : 0) + 1;
    case FMCMC_FAMILY_LOGISTIC: return d->p_x;
    case FMCMC_FAMILY_HIER_NORMAL: return d->n_groups + 1 + ((d->flags & FMCMC_MODEL_SCALES) ? 2 : 0);
  }
  return -1;
}

extern "C" int64_t fmcmc_kernel_state_len(int32_t type, int32_t k, int32_t kf) {
  switch (type) {
    case FMCMC_KERNEL_ADAPT: return (int64_t)kf * kf + kf;
    case FMCMC_KERNEL_RAM: return (int64_t)kf * kf;
    case FMCMC_KERNEL_NMIRROR:
    case FMCMC_KERNEL_UMIRROR: return 3 * (int64_t)k;
  }
  return 0;
}

extern "C" int64_t fmcmc_rows_kept(int64_t nsteps, int64_t burnin, int64_t thin) {
  int64_t m = nsteps - burnin;
  if (m < 0) return 0;
  if (thin < 1) thin = 1;
  return m / thin;
}

// --------------------------------------------------------------------------------
// model
// --------------------------------------------------------------------------------
static int model_create_impl(const fmcmc_model_desc* d, int device, bool device_ptrs, fmcmc_model** out, char* err,
                             size_t errlen) {
  if (!d || !out) { set_err(err, errlen, "null argument"); return FMCMC_EINVAL; }
  const int k = fmcmc_model_nparams(d);
  if (k < 1 || d->n < 1) { set_err(err, errlen, "bad model description (family %d, n %lld)", d->family, (long long)d->n); return FMCMC_EINVAL; }
  if (d->family != FMCMC_FAMILY_HIER_NORMAL && (d->p_x < 1 || !d->X)) { set_err(err, errlen, "X is required"); return FMCMC_EINVAL; }
  if (!d->y) { set_err(err, errlen, "y is required"); return FMCMC_EINVAL; }
  if (d->family == FMCMC_FAMILY_HIER_NORMAL && (!d->group || d->n_groups < 1)) { set_err(err, errlen, "group is required"); return FMCMC_EINVAL; }
  if (d->family == FMCMC_FAMILY_LOGISTIC && !(d->hyper[0] > 0)) { set_err(err, errlen, "logistic prior sd must be > 0"); return FMCMC_EINVAL; }
  if (d->family == FMCMC_FAMILY_HIER_NORMAL && !(d->hyper[1] > d->hyper[0])) { set_err(err, errlen, "hier_normal needs gamma bounds lo < hi"); return FMCMC_EINVAL; }
  int ndev = 0;
  CU_CHECK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { set_err(err, errlen, "CUDA device %d not available (%d visible)", device, ndev); return FMCMC_ECUDA; }
  CU_CHECK(cudaSetDevice(device));
  fmcmc_model* m = new fmcmc_model();
  m->device = device;
  const long long n = d->n, ld = (n + 1) & ~1LL;  // even leading dimension: 16-byte aligned columns for TMA
  ModelParams& mp = m->mp;
  mp.family = d->family; mp.flags = d->flags; mp.n = n; mp.ld = ld; mp.p_x = d->p_x; mp.n_groups = d->n_groups;
  mp.k = k; mp.h0 = d->hyper[0]; mp.h1 = d->hyper[1]; mp.n_total = n;
  const cudaMemcpyKind kind = device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
#define MC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err(err, errlen, "CUDA error %s (%s)", cudaGetErrorString(e_), #call); fmcmc_model_free(m); return FMCMC_ECUDA; } } while (0)
  MC(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  MC(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  MC(cudaEventCreate(&m->ev0));
  MC(cudaEventCreate(&m->ev1));
  MC(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device));
  MC(cudaDeviceGetAttribute(&m->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  MC(cudaDeviceGetAttribute(&m->l2_bytes, cudaDevAttrL2CacheSize, device));
  const bool aligned16 = (((uintptr_t)d->X | (uintptr_t)d->y) & 15) == 0;   // TMA bulk copies and vector loads need 16-byte aligned columns
  if (device_ptrs && ld == n && aligned16) {  // borrow
    m->borrowed = true;
    mp.X = d->X; mp.y = d->y; mp.group = d->group;
  } else {
    if (d->p_x > 0) {
      MC(ensure(m->X, (size_t)d->p_x * ld * 8));
      MC(cudaMemset(m->X.p, 0, (size_t)d->p_x * ld * 8));
      MC(cudaMemcpy2D(m->X.p, ld * 8, d->X, n * 8, n * 8, d->p_x, kind));
      mp.X = m->X.as<double>();
    }
    MC(ensure(m->y, (size_t)ld * 8));
    MC(cudaMemset(m->y.p, 0, (size_t)ld * 8));
    MC(cudaMemcpy(m->y.p, d->y, n * 8, kind));
    mp.y = m->y.as<double>();
    if (d->group) {
      MC(ensure(m->group, (size_t)ld * 4));
      MC(cudaMemset(m->group.p, 0, (size_t)ld * 4));
      MC(cudaMemcpy(m->group.p, d->group, n * 4, kind));
      mp.group = m->group.as<int>();
    }
  }
  if (d->family == FMCMC_FAMILY_HIER_NORMAL && !device_ptrs) {
    for (long long i = 0; i < n; i++)
      if (d->group[i] < 0 || d->group[i] >= d->n_groups) {
        set_err(err, errlen, "group[%lld] = %d out of range", i, d->group[i]);
        fmcmc_model_free(m);
        return FMCMC_EINVAL;
      }
  }
  if (d->family == FMCMC_FAMILY_LOGISTIC) {  // enables the integer-select epilogue (logistic_term_binary)
    std::vector<double> hy;
    const double* yh = d->y;
    if (device_ptrs) {
      hy.resize((size_t)n);
      MC(cudaMemcpy(hy.data(), d->y, (size_t)n * 8, cudaMemcpyDeviceToHost));
      yh = hy.data();
    }
    int bin = 1;
    for (long long i = 0; i < n && bin; i++) {
      int64_t bits;
      memcpy(&bits, &yh[i], 8);
      bin = (bits == 0 || bits == 0x3FF0000000000000LL);
    }
    mp.y_binary = bin;
    std::vector<double> tab(2 * (size_t)FM_SP_ENTRIES);  // softplus table of the tiled kernel's epilogue
    fm_softplus_table_fill(tab.data());
    MC(ensure(m->sp_tab, tab.size() * 8));
    MC(cudaMemcpy(m->sp_tab.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice));
    mp.sp_tab = m->sp_tab.as<double>();
    std::vector<double> tab4(2 * (size_t)FM_SP4_ENTRIES);  // finer table of the split-integer kernel (path 4)
    fm_softplus_table4_fill(tab4.data());
    MC(ensure(m->sp_tab4, tab4.size() * 8));
    MC(cudaMemcpy(m->sp_tab4.p, tab4.data(), tab4.size() * 8, cudaMemcpyHostToDevice));
    mp.sp_tab4 = m->sp_tab4.as<double>();
    std::vector<double> tab8(2 * (size_t)FM_SP8_ENTRIES);
    fm_lcosh_table8_fill(tab8.data());
    MC(ensure(m->sp_tab8, tab8.size() * 8));
    MC(cudaMemcpy(m->sp_tab8.p, tab8.data(), tab8.size() * 8, cudaMemcpyHostToDevice));
    mp.sp_tab8 = m->sp_tab8.as<double>();
    fm_lcosh_table8m_fill(tab8.data());
    MC(ensure(m->sp_tab8m, tab8.size() * 8));
    MC(cudaMemcpy(m->sp_tab8m.p, tab8.data(), tab8.size() * 8, cudaMemcpyHostToDevice));
    mp.sp_tab8m = m->sp_tab8m.as<double>();
    std::vector<double> tab6((size_t)FM_LC6_ENTRIES_MAX * (FM_LC6_POINT_BYTES / 8));
    fm_lcosh_table6r_fill(tab6.data());
    MC(ensure(m->sp_tab6r, tab6.size() * 8));
    MC(cudaMemcpy(m->sp_tab6r.p, tab6.data(), tab6.size() * 8, cudaMemcpyHostToDevice));
    mp.sp_tab6r = m->sp_tab6r.as<double>();
  }
  MC(ensure(m->errbuf, 4 * sizeof(int)));
  MC(ensure(m->nacc, sizeof(unsigned long long)));
#undef MC
  if (const char* v = getenv("FMCMC_TILED_VARIANT")) { if (atoi(v) == 2 || atoi(v) == 3) m->tiled_default = atoi(v); }
  if (const char* v = getenv("FMCMC_TILED_MANY")) { if (atoi(v) == 3 || atoi(v) == 4) m->tiled_many = atoi(v); }
  if (const char* v = getenv("FMCMC_PDL")) m->pdl = atoi(v) != 0;
  if (const char* v = getenv("FMCMC_HEAD_CTA")) m->head_cta = atoi(v) != 0;
  if (const char* v = getenv("FMCMC_L2_KEEP_MB")) m->l2_keep_mb = atoi(v);
  if (const char* v = getenv("FMCMC_I8_GSL")) { const int g = atoi(v); if (g == 1 || g == 2 || g == 4) m->i8_gsl = g; }
  if (const char* v = getenv("FMCMC_MMA_VARIANT")) m->mma_wide = atoi(v);
  if (const char* v = getenv("FMCMC_PATH")) { if (atoi(v) >= 1 && atoi(v) <= 4) m->forced_path = atoi(v); }  // tuning / profiling only
  if (const char* v = getenv("FMCMC_I8_SLICES")) { if (atoi(v) >= I8_NS_LO && atoi(v) <= I8_NS_LO + 1) m->i8_slices = atoi(v); }
  *out = m;
  return FMCMC_OK;
}

extern "C" int fmcmc_model_create(const fmcmc_model_desc* d, int device, fmcmc_model** out, char* err, size_t errlen) {
  return model_create_impl(d, device, false, out, err, errlen);
}
extern "C" int fmcmc_model_create_device(const fmcmc_model_desc* d, int device, fmcmc_model** out, char* err,
                                         size_t errlen) {
  return model_create_impl(d, device, true, out, err, errlen);
}

extern "C" void fmcmc_model_free(fmcmc_model* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  DevBuf* bufs[] = {&m->X, &m->y, &m->group, &m->sp_tab, &m->sp_tab4, &m->sp_tab8, &m->sp_tab8m, &m->sp_tab6r, &m->Xt, &m->Xq, &m->xq_bad, &m->xq_aux, &m->ans, &m->draws, &m->logpost, &m->cur_theta, &m->cur_f, &m->prop,
                    &m->prop_u, &m->istate, &m->dstate, &m->colsum, &m->ubuf, &m->work, &m->cflags, &m->errbuf,
                    &m->nacc, &m->spec, &m->fed_logu, &m->fed_z, &m->initial, &m->partial, &m->out_ans,
                    &m->out_draws, &m->out_lp, &m->tmp, &m->store, &m->g_xbar, &m->g_s2, &m->g_wsum, &m->g_wpart,
                    &m->g_mask};
  for (DevBuf* b : bufs) release(*b);
  for (int g = 0; g < FM_MAX_PEERS; g++)
    if (m->shard_ipc_opened[g]) { cudaIpcCloseMemHandle(m->shard_peer_partial[g]); cudaIpcCloseMemHandle(m->shard_peer_flags[g]); }
  if (m->shard_partial) cudaFree(m->shard_partial);
  if (m->shard_flags) cudaFree(m->shard_flags);
  if (m->shard_done) cudaFree(m->shard_done);
  for (auto& e : m->hot_ev) cudaEventDestroy(e);
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  for (auto& e : m->chunk_ev) cudaEventDestroy(e);
  for (auto& e : m->marks) if (e) cudaEventDestroy(e);
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

// --------------------------------------------------------------------------------
// observation sharding across GPUs
// --------------------------------------------------------------------------------
extern "C" int fmcmc_shard_alloc(fmcmc_model* m, int world, int max_cols, int64_t n_total, fmcmc_shard_handles* out,
                                 char* err, size_t errlen) {
  if (!m || !out || world < 2 || world > FM_MAX_PEERS || max_cols < 1 || n_total < m->mp.n) {
    set_err(err, errlen, "fmcmc_shard_alloc: bad argument (world 2..%d, max_cols >= 1, n_total >= local n)", FM_MAX_PEERS);
    return FMCMC_EINVAL;
  }
  if (m->shard_partial) { set_err(err, errlen, "fmcmc_shard_alloc: the model is already sharded"); return FMCMC_EINVAL; }
  CU_CHECK(cudaSetDevice(m->device));
  const size_t pbytes = (size_t)2 * world * m->sm_count * max_cols * 8;
  CU_CHECK(cudaMalloc((void**)&m->shard_partial, pbytes));
  CU_CHECK(cudaMalloc((void**)&m->shard_flags, (size_t)FM_MAX_PEERS * 8));
  CU_CHECK(cudaMalloc((void**)&m->shard_done, 16));
  CU_CHECK(cudaMemset(m->shard_partial, 0, pbytes));
  CU_CHECK(cudaMemset(m->shard_flags, 0, (size_t)FM_MAX_PEERS * 8));
  CU_CHECK(cudaMemset(m->shard_done, 0, 16));
  CU_CHECK(cudaDeviceSynchronize());
  memset(out, 0, sizeof(*out));
  cudaIpcMemHandle_t h;
  CU_CHECK(cudaIpcGetMemHandle(&h, m->shard_partial));
  memcpy(out->partial, &h, sizeof(h));
  CU_CHECK(cudaIpcGetMemHandle(&h, m->shard_flags));
  memcpy(out->flags, &h, sizeof(h));
  out->partial_ptr = m->shard_partial;
  out->flags_ptr = m->shard_flags;
  out->device = m->device;
  out->pid = (int32_t)getpid();
  m->shard_world = -world;  // allocated, not attached yet
  m->shard_max_cols = max_cols;
  m->mp.n_total = n_total;
  return FMCMC_OK;
}

extern "C" int fmcmc_shard_attach(fmcmc_model* m, int rank, int world, const fmcmc_shard_handles* all, char* err,
                                  size_t errlen) {
  if (!m || !all || m->shard_world != -world || rank < 0 || rank >= world) {
    set_err(err, errlen, "fmcmc_shard_attach: call fmcmc_shard_alloc with the same world first");
    return FMCMC_EINVAL;
  }
  if (all[rank].partial_ptr != m->shard_partial || all[rank].pid != (int32_t)getpid()) {
    set_err(err, errlen, "fmcmc_shard_attach: all[rank] is not this model's own handle");
    return FMCMC_EINVAL;
  }
  CU_CHECK(cudaSetDevice(m->device));
  for (int g = 0; g < world; g++) {
    if (g == rank) {
      m->shard_peer_partial[g] = m->shard_partial;
      m->shard_peer_flags[g] = m->shard_flags;
    } else if (all[g].pid == (int32_t)getpid()) {  // peer model in this process: direct peer access
      int can = 0;
      CU_CHECK(cudaDeviceCanAccessPeer(&can, m->device, all[g].device));
      if (!can) { set_err(err, errlen, "GPU %d cannot access GPU %d's memory (no NVLink / P2P)", m->device, all[g].device); return FMCMC_ECUDA; }
      cudaError_t pe = cudaDeviceEnablePeerAccess(all[g].device, 0);
      if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CU_CHECK(pe);
      cudaGetLastError();
      m->shard_peer_partial[g] = (double*)all[g].partial_ptr;
      m->shard_peer_flags[g] = (unsigned long long*)all[g].flags_ptr;
    } else {  // peer process: CUDA IPC
      cudaIpcMemHandle_t h;
      void* p = nullptr;
      memcpy(&h, all[g].partial, sizeof(h));
      CU_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      m->shard_peer_partial[g] = (double*)p;
      memcpy(&h, all[g].flags, sizeof(h));
      CU_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      m->shard_peer_flags[g] = (unsigned long long*)p;
      m->shard_ipc_opened[g] = true;
    }
  }
  m->shard_world = world;
  m->shard_rank = rank;
  return FMCMC_OK;
}

// Device-clock stopwatch on the library's launch stream (bench.py: a timed region that spans several ABI calls -
// stepping, the R-hat statistics, the host finish - is bracketed by two marks; torch.cuda.Event would only see torch's stream).
extern "C" int fmcmc_event_mark(fmcmc_model* m, int slot) {
  if (!m || slot < 0 || slot > 7) return FMCMC_EINVAL;
  if (cudaSetDevice(m->device) != cudaSuccess) return FMCMC_ECUDA;
  if (!m->marks[slot] && cudaEventCreate(&m->marks[slot]) != cudaSuccess) return FMCMC_ECUDA;
  return cudaEventRecord(m->marks[slot], m->stream) == cudaSuccess ? FMCMC_OK : FMCMC_ECUDA;
}
extern "C" int fmcmc_event_elapsed_ms(fmcmc_model* m, int from, int to, double* ms) {
  if (!m || !ms || from < 0 || from > 7 || to < 0 || to > 7 || !m->marks[from] || !m->marks[to]) return FMCMC_EINVAL;
  if (cudaSetDevice(m->device) != cudaSuccess || cudaEventSynchronize(m->marks[to]) != cudaSuccess) return FMCMC_ECUDA;
  float t = 0.f;
  if (cudaEventElapsedTime(&t, m->marks[from], m->marks[to]) != cudaSuccess) return FMCMC_ECUDA;
  *ms = t;
  return FMCMC_OK;
}

extern "C" int fmcmc_set_path(fmcmc_model* m, int path) {
  if (!m || path < 0 || path > 4) return FMCMC_EINVAL;
  m->forced_path = path;
  return FMCMC_OK;
}

extern "C" int fmcmc_model_trim(fmcmc_model* m, int path, char* err, size_t errlen) {
  if (!m || (path != 3 && path != 4)) { set_err(err, errlen, "fmcmc_model_trim: path must be 3 or 4"); return FMCMC_EINVAL; }
  if ((path == 4 && m->xq_NS <= 0) || (path == 3 && m->xt_PB == 0)) {
    set_err(err, errlen, "fmcmc_model_trim: path %d has not run on this model yet (its copy of X is built by the first run)", path);
    return FMCMC_EINVAL;
  }
  if (m->shard_world > 1 && path != 3) { set_err(err, errlen, "observation-sharded models run path 3"); return FMCMC_EINVAL; }
  CU_CHECK(cudaSetDevice(m->device));
  CU_CHECK(cudaStreamSynchronize(m->stream));
  if (path == 4) { release(m->Xt); m->xt_PB = 0; m->mp.Xt = nullptr; }
  if (path == 3) { release(m->Xq); m->xq_NS = 0; m->xq_KB = 0; m->mp.Xq = nullptr; }
  if (!m->borrowed) { release(m->X); m->mp.X = nullptr; }   // y stays: the Gaussian epilogue of path 4 and every head read it
  m->trimmed_to = path;
  m->forced_path = path;
  return FMCMC_OK;
}

// --------------------------------------------------------------------------------
// validation: same conditions / message substrings as the R code
// --------------------------------------------------------------------------------
static int validate_run(const fmcmc_model* m, const fmcmc_run_spec* run, const fmcmc_kernel_spec* ks, int kf,
                        char* err, size_t errlen) {
  const int k = m->mp.k;
  if (run->nchains < 1) { set_err(err, errlen, "`nchains` must be an integer greater than 1."); return FMCMC_EINVAL; }
  if (run->burnin >= run->nsteps) {  // R/mcmc.R:512-513
    set_err(err, errlen, "-burnin- (%lld) cannot be >= than -nsteps- (%lld).", (long long)run->burnin, (long long)run->nsteps);
    return FMCMC_EINVAL;
  }
  if (run->thin >= run->nsteps) {  // R/mcmc.R:516-517
    set_err(err, errlen, "-thin- (%lld) cannot be > than -nsteps- (%lld).", (long long)run->thin, (long long)run->nsteps);
    return FMCMC_EINVAL;
  }
  if (run->thin < 1) { set_err(err, errlen, "-thin- should be >= 1."); return FMCMC_EINVAL; }  // :519-520
  if (run->burnin < 0) { set_err(err, errlen, "-burnin- cannot be negative."); return FMCMC_EINVAL; }
  if (ks->k != k) { set_err(err, errlen, "Incorrect length of -initial-: the kernel has k = %d, the family needs %d parameters.", ks->k, k); return FMCMC_EINVAL; }
  if (ks->type < FMCMC_KERNEL_NORMAL || ks->type > FMCMC_KERNEL_UMIRROR) { set_err(err, errlen, "unknown kernel type %d", ks->type); return FMCMC_EINVAL; }
  if (kf == 0) {  // R/kernel.R:125-128
    set_err(err, errlen, "The number of parameters to update, i.e. not fixed, cannot be zero. Check the value -fixed- in the kernel initialization.");
    return FMCMC_EINVAL;
  }
  const bool bounded = !(ks->type == FMCMC_KERNEL_NORMAL || ks->type == FMCMC_KERNEL_UNIF);
  if (bounded)
    for (int j = 0; j < k; j++)
      if (ks->ub[j] <= ks->lb[j]) { set_err(err, errlen, "-ub- cannot be <= than -lb-."); return FMCMC_EINVAL; }  // R/kernel_normal.R:134-135
  if (ks->type == FMCMC_KERNEL_UNIF || ks->type == FMCMC_KERNEL_UNIF_REFLECTIVE)
    for (int j = 0; j < k; j++)
      if (ks->max_[j] <= ks->min_[j]) { set_err(err, errlen, "-max.- cannot be <= than -min.-."); return FMCMC_EINVAL; }
  if (ks->type == FMCMC_KERNEL_UMIRROR && (kf != k || ks->scheme != FMCMC_SCHEME_JOINT)) {
    set_err(err, errlen, "kernel_umirror with fixed parameters or a non-joint scheme is ill-defined in the reference (quirk D11)");
    return FMCMC_EUNSUP;
  }
  if (ks->type == FMCMC_KERNEL_ADAPT) {
    if (ks->bw > 0 && ks->bw > ks->warmup) { set_err(err, errlen, "The `warmup` parameter must be greater than `bw`."); return FMCMC_EINVAL; }
    if (ks->mvn_method != FMCMC_MVN_CHOLESKY && ks->mvn_method != FMCMC_MVN_EIGEN) { set_err(err, errlen, "unknown mvn_method %d", ks->mvn_method); return FMCMC_EINVAL; }
    if (ks->freq < 1) { set_err(err, errlen, "-freq- must be >= 1."); return FMCMC_EINVAL; }
  }
  if (ks->type == FMCMC_KERNEL_RAM && ks->freq < 1) { set_err(err, errlen, "-freq- must be >= 1."); return FMCMC_EINVAL; }
  const bool has_scheme = ks->type == FMCMC_KERNEL_NORMAL || ks->type == FMCMC_KERNEL_NORMAL_REFLECTIVE ||
                          ks->type == FMCMC_KERNEL_UNIF || ks->type == FMCMC_KERNEL_UNIF_REFLECTIVE ||
                          ks->type == FMCMC_KERNEL_NMIRROR || ks->type == FMCMC_KERNEL_UMIRROR;
  if (has_scheme) {
    if (ks->scheme < FMCMC_SCHEME_JOINT || ks->scheme > FMCMC_SCHEME_EXPLICIT) {
      set_err(err, errlen, "-scheme- update must be either an integer sequence, 'joint', 'ordered', or 'random'.");
      return FMCMC_EINVAL;
    }
    if (ks->scheme == FMCMC_SCHEME_EXPLICIT) {  // R/kernel.R:69-91
      if (ks->order_len != kf || !ks->order) {
        set_err(err, errlen, "When setting the update scheme, it should have the same length as the number of variables that will not be fixed. Right now length(scheme) = %d while sum(!fixed) = %d.", ks->order_len, kf);
        return FMCMC_EINVAL;
      }
      for (int j = 0; j < k; j++) {
        if (ks->fixed && ks->fixed[j]) continue;
        bool found = false;
        for (int q = 0; q < ks->order_len; q++) found |= (ks->order[q] == j + 1);
        if (!found) {
          set_err(err, errlen, "One or more variables was not included in the ordering sequence. Only variables that are not fixed can be included in this list.");
          return FMCMC_EINVAL;
        }
      }
    }
    if (ks->scheme == FMCMC_SCHEME_RANDOM && ks->seq && ks->seq_len < run->nsteps) {
      set_err(err, errlen, "the planned random update sequence (%lld rows) is shorter than nsteps (%lld) (quirk D7: subscript out of bounds in the reference)", (long long)ks->seq_len, (long long)run->nsteps);
      return FMCMC_EUNSUP;
    }
  }
  return FMCMC_OK;
}

// --------------------------------------------------------------------------------
// output gather (burnin / thin, R/mcmc.R:786-813) and store append
// --------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const double* __restrict__ src, double* __restrict__ dst, int C, int k,
                                   long long keep, long long burnin, long long thin, int colmajor) {
  const long long total = (long long)C * keep * k;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    long long c, r;
    int j;
    if (colmajor) { r = e % keep; j = (int)((e / keep) % k); c = e / (keep * k); }
    else { j = (int)(e % k); r = (e / k) % keep; c = e / ((long long)k * keep); }
    const long long srow = burnin + (r + 1) * thin - 1;
    dst[e] = src[((size_t)srow * C + c) * k + j];
  }
}
// kept rows [r0, r1) only, into the same [C][keep][k] layout (streamed outputs)
__global__ void gather_rows_range_kernel(const double* __restrict__ src, double* __restrict__ dst, int C, int k,
                                         long long keep, long long r0, long long r1, long long burnin, long long thin,
                                         int colmajor) {
  const long long nr = r1 - r0, total = (long long)C * nr * k;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    if (colmajor) {  // R matrices: [chain][param][row]
      const long long r = r0 + e % nr, j = (e / nr) % k, c = e / (nr * k);
      const long long srow = burnin + (r + 1) * thin - 1;
      dst[((size_t)c * k + j) * keep + r] = src[((size_t)srow * C + c) * k + j];
    } else {
      const int j = (int)(e % k);
      const long long r = r0 + (e / k) % nr, c = e / ((long long)k * nr);
      const long long srow = burnin + (r + 1) * thin - 1;
      dst[((size_t)c * keep + r) * k + j] = src[((size_t)srow * C + c) * k + j];
    }
  }
}
__global__ void append_rows_kernel(const double* __restrict__ src, double* __restrict__ dst, long long rowlen,
                                   long long keep, long long burnin, long long thin) {
  const long long total = keep * rowlen;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / rowlen, o = e % rowlen;
    dst[e] = src[(size_t)(burnin + (r + 1) * thin - 1) * rowlen + o];
  }
}

// --------------------------------------------------------------------------------
// fmcmc_run
// --------------------------------------------------------------------------------
struct Blob {  // host staging of the small kernel-spec arrays -> one H2D copy
  std::vector<unsigned char> h;
  size_t add(const void* p, size_t bytes) {
    size_t off = (h.size() + 15) & ~(size_t)15;
    h.resize(off + bytes);
    if (p) memcpy(h.data() + off, p, bytes);
    return off;
  }
};

// Launch with (pdl) or without the programmatic-stream-serialization attribute: with it the kernel's CTAs may become resident
// while the previous kernel of the stream is still running; they order themselves behind it with griddepcontrol.wait (pdl_wait,
// common.cuh) before touching anything it wrote.
template <typename... KArgs, typename... Args>
static cudaError_t launch_chained(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

template <int FAMILY, bool YBIN>
static cudaError_t launch_tiled_loglik(fmcmc_model* m, int PB, dim3 grid, const RunBuffers& rb, const TiledBuffers& tb, bool pdl) {
  const size_t smem = tiled_smem_bytes(PB, FAMILY);
#define TL_CASE(P)                                                                                               \
  case P: {                                                                                                      \
    static bool attr_done[64] = {};                                                                              \
    if (!attr_done[m->device]) {                                                                                 \
      cudaError_t e = cudaFuncSetAttribute(tiled_loglik_kernel<FAMILY, P, YBIN>,                                     \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
      if (e != cudaSuccess) return e;                                                                            \
      attr_done[m->device] = true;                                                                               \
    }                                                                                                            \
    { cudaError_t e = launch_chained(tiled_loglik_kernel<FAMILY, P, YBIN>, grid, dim3(TL_THREADS), smem, m->stream, pdl, \
                                     m->mp, rb.prop, rb.prop_u, rb.nchains, tb, rb.err);                          \
      if (e != cudaSuccess) return e; }                                                                          \
    break;                                                                                                       \
  }
  switch (PB) {
    TL_CASE(8)
    TL_CASE(16)
    TL_CASE(32)
    default: return cudaErrorInvalidValue;
  }
#undef TL_CASE
  return cudaGetLastError();
}

// DMMA variant (tiled_mma.cuh).  Geometry per padded width PB: (warps, chain tiles per warp).
struct MmaShape { int PB, warps, NT, MO, osplit, pipe; };
// Few columns (chains): all warps share NT*8 chains and split the observations (HBM-bound mapping);
// many: every warp owns NT*8 chains (FP64-bound mapping).  tiled_mma.cuh explains both.
static MmaShape mma_shape(int p_x, int ncols, int variant) {
  if (p_x <= 32) {
    if (ncols <= 8) return MmaShape{32, 8, 1, 2, 1};
    if (ncols <= 16) return MmaShape{32, 8, 2, 2, 1};
    if (ncols <= 128) return MmaShape{32, 8, 4, 2, 1};
    switch (variant) {  // measured on B200, cfg3, ms per launch (profiles/r01_*): 3.30 default; 3.33 / 3.38 / 3.41 / 3.59 / 3.61 / 3.62
      case 1: return MmaShape{32, 8, 4, 2, 0, 0};
      case 2: return MmaShape{32, 8, 8, 1, 0, 0};
      case 3: return MmaShape{32, 8, 4, 4, 0, 0};
      case 4: return MmaShape{32, 8, 4, 1, 0, 1};
      case 6: return MmaShape{32, 16, 4, 1, 0, 1};
      case 7: return MmaShape{32, 16, 4, 1, 0, 0};
      default: return MmaShape{32, 8, 4, 2, 0, 1};  // software-pipelined: DMMAs of block t+1 ahead of the epilogue of block t
    }
  }
  if (p_x <= 64) return ncols <= 32 ? MmaShape{64, 8, 4, 1, 1} : MmaShape{64, 8, 4, 1, 0};
  if (ncols <= 16) return MmaShape{128, 4, 2, 1, 1};
  return variant == 1 ? MmaShape{128, 8, 2, 1, 0} : MmaShape{128, 8, 2, 2, 0};
}
// Builds the tile-major copy of X / y (tiled_mma.cuh) once per model.
static cudaError_t ensure_packed_tiles(fmcmc_model* m, int PB) {
  if (m->xt_PB == PB) return cudaSuccess;
  const ModelParams& mp = m->mp;
  const int TR = PB <= 32 ? 128 : (PB == 64 ? 64 : 32);
  const long long ntiles = (mp.n + TR - 1) / TR;
  const size_t stage_doubles = (size_t)PB * (TR + 4) + TR;
  cudaError_t e = ensure(m->Xt, (size_t)ntiles * stage_doubles * 8);
  if (e != cudaSuccess) return e;
  double* xt = m->Xt.as<double>();
  switch (PB) {
    case 32: pack_tiles_kernel<32><<<(unsigned)ntiles, 256, 0, m->stream>>>(mp.X, mp.y, mp.n, mp.ld, mp.p_x, xt); break;
    case 64: pack_tiles_kernel<64><<<(unsigned)ntiles, 256, 0, m->stream>>>(mp.X, mp.y, mp.n, mp.ld, mp.p_x, xt); break;
    case 128: pack_tiles_kernel<128><<<(unsigned)ntiles, 256, 0, m->stream>>>(mp.X, mp.y, mp.n, mp.ld, mp.p_x, xt); break;
    default: return cudaErrorInvalidValue;
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  m->mp.Xt = xt;
  m->xt_PB = PB;
  return cudaSuccess;
}

template <int FAMILY, bool YBIN>
static cudaError_t launch_tiled_mma(fmcmc_model* m, const MmaShape& sh, dim3 grid, const RunBuffers& rb,
                                    const TiledBuffers& tb, bool pdl) {
#define TM_CASE(P, W, N, O, S, PP)                                                                                      \
  if (sh.PB == P && sh.warps == W && sh.NT == N && sh.MO == O && sh.osplit == (S ? 1 : 0) && sh.pipe == (PP ? 1 : 0)) {                                                             \
    const size_t smem = tiled_mma_smem_bytes<P>(FAMILY);                                                         \
    static bool attr_done[64] = {};                                                                              \
    if (!attr_done[m->device]) {                                                                                 \
      cudaError_t e = cudaFuncSetAttribute(tiled_loglik_mma_kernel<FAMILY, P, YBIN, W, N, O, S, PP>,                    \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
      if (e != cudaSuccess) return e;                                                                            \
      attr_done[m->device] = true;                                                                               \
    }                                                                                                            \
    return launch_chained(tiled_loglik_mma_kernel<FAMILY, P, YBIN, W, N, O, S, PP>, grid, dim3(W * 32), smem, m->stream, pdl, \
                          m->mp, rb.prop, rb.prop_u, rb.nchains, tb, rb.err);                                    \
  }
  TM_CASE(32, 8, 1, 2, true, false)
  TM_CASE(32, 8, 2, 2, true, false)
  TM_CASE(32, 8, 4, 2, true, false)
  TM_CASE(32, 8, 4, 2, false, false)
  TM_CASE(32, 8, 4, 4, false, false)
  TM_CASE(32, 8, 8, 1, false, false)
  TM_CASE(32, 16, 4, 1, false, false)
  TM_CASE(32, 8, 4, 1, false, true)
  TM_CASE(32, 8, 4, 2, false, true)
  TM_CASE(32, 16, 4, 1, false, true)
  TM_CASE(64, 8, 4, 1, true, false)
  TM_CASE(64, 8, 4, 1, false, false)
  TM_CASE(128, 4, 2, 1, true, false)
  TM_CASE(128, 8, 2, 1, false, false)
  TM_CASE(128, 8, 2, 2, false, false)
#undef TM_CASE
  return cudaErrorInvalidValue;
}

// ---- path 4: split-integer tensor-core kernel (tiled_i8.cuh) ------------------------------------------
static int i8_kblocks(int p_x) { return p_x <= 32 ? 1 : (p_x <= 64 ? 2 : 4); }
#if I8_DIGIT_BITS == 8
#define I8_FOR_SHAPES(X) X(5, 1) X(5, 2) X(5, 4) X(6, 1) X(6, 2) X(6, 4)
#else
#define I8_FOR_SHAPES(X) X(6, 1) X(6, 2) X(6, 4) X(7, 1) X(7, 2) X(7, 4)
#endif
// observations per pipeline stage / tile of Xq (the Gaussian family at K = 128 packs 64-observation blocks: tiled_i8.cuh, i8_blk)
static int i8_tile_rows(int family, int NS, int KB) {
#define I8_CASE(N, K) if (NS == N && KB == K) return family == FMCMC_FAMILY_GAUSSIAN_LM ? I8Geom<N, K, i8_blk<N, K>(FMCMC_FAMILY_GAUSSIAN_LM)>::TO : I8Geom<N, K, i8_blk<N, K>(FMCMC_FAMILY_LOGISTIC)>::TO;
  I8_FOR_SHAPES(I8_CASE)
#undef I8_CASE
  return 32;
}
// Builds the int8 slice tiles of X once per model.  Returns cudaErrorNotSupported when X holds non-finite
// values (or magnitudes beyond the exponent window): the caller falls back to the FP64 kernels.
static cudaError_t ensure_packed_i8(fmcmc_model* m, int NS, int KB) {
  if (m->xq_NS == NS && m->xq_KB == KB) return cudaSuccess;
  if (m->xq_NS == -1) return cudaErrorNotSupported;
  const ModelParams& mp = m->mp;
  const bool gauss = mp.family == FMCMC_FAMILY_GAUSSIAN_LM;
  const int TO = i8_tile_rows(mp.family, NS, KB);
  const long long ntiles = (mp.n + TO - 1) / TO;
  size_t stage_bytes = 0;
#define I8_CASE(N, K) if (NS == N && KB == K) stage_bytes = gauss ? I8Geom<N, K, i8_blk<N, K>(FMCMC_FAMILY_GAUSSIAN_LM)>::STAGE_BYTES : I8Geom<N, K, i8_blk<N, K>(FMCMC_FAMILY_LOGISTIC)>::STAGE_BYTES;
  I8_FOR_SHAPES(I8_CASE)
#undef I8_CASE
  if (!stage_bytes) return cudaErrorInvalidValue;
  cudaError_t e = ensure(m->Xq, (size_t)ntiles * stage_bytes);
  if (e != cudaSuccess) return e;
  e = ensure(m->xq_bad, sizeof(int));
  if (e != cudaSuccess) return e;
  // aux: [p_x] column maxima + [1] largest squared row norm (u64 bit patterns of non-negative doubles) | [p_x] sxy (f64) |
  // [p_x] column exponents (i32)
  const size_t px = (size_t)mp.p_x;
  e = ensure(m->xq_aux, (px + 1) * 8 + px * 8 + px * 4);
  if (e != cudaSuccess) return e;
  unsigned long long* colmax = m->xq_aux.as<unsigned long long>();
  double* sxy = reinterpret_cast<double*>(colmax + px + 1);
  int* cexp = reinterpret_cast<int*>(sxy + px);
  e = cudaMemsetAsync(m->xq_bad.p, 0, sizeof(int), m->stream);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(m->xq_aux.p, 0, px * 20 + 8, m->stream);
  if (e != cudaSuccess) return e;
  unsigned char* xq = m->Xq.as<unsigned char>();
  const unsigned ychunks = (unsigned)std::min<long long>(64, (mp.n + 255) / 256);
  i8_colmax_kernel<<<dim3((unsigned)mp.p_x, ychunks), 256, 0, m->stream>>>(mp.X, mp.n, mp.ld, colmax, m->xq_bad.as<int>());
  i8_colexp_kernel<<<(mp.p_x + 127) / 128, 128, 0, m->stream>>>(colmax, mp.p_x, cexp);
  i8_rownorm_kernel<<<(unsigned)std::min<long long>(4096, (mp.n + 255) / 256), 256, 0, m->stream>>>(mp.X, mp.n, mp.ld, mp.p_x, colmax + px);
  i8_sxy_kernel<<<(unsigned)mp.p_x, 1024, 0, m->stream>>>(mp.X, mp.y, mp.n, mp.ld, sxy);
#define I8_CASE(N, K)                                                                                                      \
  if (NS == N && KB == K) {                                                                                                \
    constexpr int BW = i8_blk<N, K>(FMCMC_FAMILY_GAUSSIAN_LM), BL = i8_blk<N, K>(FMCMC_FAMILY_LOGISTIC);                   \
    if (gauss) pack_i8_kernel<N, K, BW><<<(unsigned)ntiles, TO, 0, m->stream>>>(mp.X, mp.n, mp.ld, mp.p_x, cexp, xq);      \
    else pack_i8_kernel<N, K, BL><<<(unsigned)ntiles, TO, 0, m->stream>>>(mp.X, mp.n, mp.ld, mp.p_x, cexp, xq);            \
  }
  I8_FOR_SHAPES(I8_CASE)
#undef I8_CASE
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  int bad = 0;
  e = cudaMemcpyAsync(&bad, m->xq_bad.p, sizeof(int), cudaMemcpyDeviceToHost, m->stream);
  if (e != cudaSuccess) return e;
  e = cudaStreamSynchronize(m->stream);
  if (e != cudaSuccess) return e;
  if (bad) { m->xq_NS = -1; release(m->Xq); return cudaErrorNotSupported; }
  m->mp.i8_cexp = cexp;
  m->mp.i8_sxy = sxy;
  m->mp.i8_cmax = reinterpret_cast<const double*>(colmax);
  m->mp.Xq = xq;
  m->xq_NS = NS;
  m->xq_KB = KB;
  return cudaSuccess;
}

#ifndef I8_EPI_WARPS
#define I8_EPI_WARPS 16   // 4 epilogue warps per TMEM lane quarter, 8 observations each per block (B200: 2.37 ms per cfg3
#define I8_EPI_CHUNK 8    // launch; 8 warps x 2 chunks of 8: 2.78 ms; 16 x 2 chunks of 4: 2.75 ms - profiles/r01_i8_*.txt;
                          // on the final 1.79 ms kernel: 8 warps 2.01 ms, 8 warps with the chunk loop unrolled 1.92 ms)
#endif
template <int FAMILY, bool YBIN>
static cudaError_t launch_tiled_i8(fmcmc_model* m, int NS, int KB, dim3 grid, const RunBuffers& rb, const TiledBuffers& tb, bool pdl) {
#define I8_CASE(N, K)                                                                                             \
  if (NS == N && KB == K) {                                                                                       \
    const size_t smem = tiled_i8_smem_bytes<N, K>(FAMILY, YBIN);                                                        \
    static bool attr_done[64] = {};                                                                               \
    if (!attr_done[m->device]) {                                                                                  \
      cudaError_t e = cudaFuncSetAttribute(tiled_loglik_i8_kernel<FAMILY, YBIN, N, K, I8_EPI_WARPS, I8_EPI_CHUNK>, \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);               \
      if (e != cudaSuccess) { cudaGetLastError(); return e; }                                                     \
      attr_done[m->device] = true;                                                                                \
    }                                                                                                             \
    return launch_chained(tiled_loglik_i8_kernel<FAMILY, YBIN, N, K, I8_EPI_WARPS, I8_EPI_CHUNK>, grid,          \
                          dim3((I8_EPI_WARPS + 2) * 32), smem, m->stream, pdl, m->mp, rb.prop, rb.prop_u, rb.nchains, tb, rb.err); \
  }
  I8_FOR_SHAPES(I8_CASE)
#undef I8_CASE
  return cudaErrorInvalidValue;
}

extern "C" int fmcmc_run(fmcmc_model* m, const fmcmc_run_spec* run, const fmcmc_kernel_spec* ks,
                         fmcmc_kernel_state* state, const fmcmc_stream_spec* stream, double* ans_out,
                         double* draws_out, double* logpost_out, fmcmc_run_report* report, char* err,
                         size_t errlen) {
  if (!m || !run || !ks || !stream) { set_err(err, errlen, "null argument"); return FMCMC_EINVAL; }
  CU_CHECK(cudaSetDevice(m->device));
  std::vector<int> free_idx;
  const int kf = count_free(ks, free_idx);
  int rc = validate_run(m, run, ks, kf, err, errlen);
  if (rc) return rc;
  const int k = m->mp.k, C = run->nchains;
  const long long T = run->nsteps;
  const long long keep = fmcmc_rows_kept(T, run->burnin, run->thin);
  const long long dlen = fmcmc_kernel_state_len(ks->type, k, kf);
  if (!run->initial && (m->state_nchains != C || m->state_k != k)) {
    set_err(err, errlen, "initial is NULL but the model holds no state for %d chains", C);
    return FMCMC_EINVAL;
  }
  if (stream->mode == FMCMC_STREAM_FED && (!stream->logu || !stream->z || stream->kdraw < 1)) {
    set_err(err, errlen, "fed stream needs logu, z and kdraw");
    return FMCMC_EINVAL;
  }
  {
    const int need = (ks->type == FMCMC_KERNEL_ADAPT || ks->type == FMCMC_KERNEL_RAM) ? kf
                     : (ks->scheme == FMCMC_SCHEME_JOINT ? kf : 1);
    if (stream->mode == FMCMC_STREAM_FED && stream->kdraw < need) {
      set_err(err, errlen, "fed stream has %d draws per row, the kernel needs %d", stream->kdraw, need);
      return FMCMC_EINVAL;
    }
  }
  if (report) {
    memset(report, 0, sizeof(*report));
    report->rows_kept = keep;
    report->first_iter = run->burnin + run->thin;
    report->last_iter = run->burnin + keep * run->thin;
  }

  long long h2d = 0, d2h = 0;
  // host output arrays may be larger than this call's rows (a bulk loop fills ONE set of arrays, bulk after bulk)
  const long long host_rows = run->out_rows_total > 0 ? run->out_rows_total : keep, host_off = run->out_rows_total > 0 ? run->out_row_offset : 0;
  if (host_off < 0 || host_off + keep > host_rows) {
    set_err(err, errlen, "out_row_offset %lld + %lld kept rows exceed out_rows_total %lld", host_off, (long long)keep, host_rows);
    return FMCMC_EINVAL;
  }
  const bool dev_state = (run->flags & FMCMC_RUN_DEVICE_STATE) != 0;
  if (dev_state && (m->state_nchains != C || m->state_k != k || m->state_type != ks->type)) {
    set_err(err, errlen, "FMCMC_RUN_DEVICE_STATE: the model holds no kernel state of this shape");
    return FMCMC_EINVAL;
  }

  // ---- kernel spec -> device ----------------------------------------------------------
  Blob blob;
  KParams kp{};
  kp.type = ks->type; kp.k = k; kp.kf = kf; kp.scheme = ks->scheme; kp.order_len = ks->order_len;
  kp.nadapt_len = ks->nadapt_len; kp.seq_len = ks->seq_len; kp.mvn_method = ks->mvn_method;
  kp.warmup = ks->warmup; kp.freq = ks->freq < 1 ? 1 : ks->freq; kp.bw = ks->bw;
  kp.until = ks->until; kp.eps = ks->eps; kp.Sd = ks->Sd; kp.arate = ks->arate; kp.dlen = dlen;
  std::vector<double> dflt0(k, 0.0), dflt1(k, 1.0), dfltlo(k, -1.79769313486231570815e308), dflthi(k, 1.79769313486231570815e308);
  std::vector<unsigned char> fixed0(k, 0);
  const size_t o_mu = blob.add(ks->mu ? ks->mu : dflt0.data(), k * 8);
  const size_t o_scale = blob.add(ks->scale ? ks->scale : dflt1.data(), k * 8);
  const size_t o_min = blob.add(ks->min_ ? ks->min_ : dflt0.data(), k * 8);
  const size_t o_max = blob.add(ks->max_ ? ks->max_ : dflt1.data(), k * 8);
  const size_t o_lb = blob.add(ks->lb ? ks->lb : dfltlo.data(), k * 8);
  const size_t o_ub = blob.add(ks->ub ? ks->ub : dflthi.data(), k * 8);
  const size_t o_fixed = blob.add(ks->fixed ? ks->fixed : fixed0.data(), k);
  const size_t o_free = blob.add(free_idx.data(), kf * sizeof(int));
  const size_t o_order = (ks->scheme == FMCMC_SCHEME_EXPLICIT) ? blob.add(ks->order, ks->order_len * sizeof(int)) : 0;
  const size_t o_nadapt = ks->nadapt_len > 0 ? blob.add(ks->nadapt, ks->nadapt_len * sizeof(long long)) : 0;
  const size_t o_constr = ks->constr ? blob.add(ks->constr, (size_t)k * k * 8) : 0;
  const bool fed_seq = ks->scheme == FMCMC_SCHEME_RANDOM && ks->seq;
  const size_t o_seq = fed_seq ? blob.add(ks->seq, (size_t)C * ks->seq_len * sizeof(int)) : 0;
  CU_CHECK(ensure(m->spec, blob.h.size()));
  CU_CHECK(cudaMemcpyAsync(m->spec.p, blob.h.data(), blob.h.size(), cudaMemcpyHostToDevice, m->stream));
  h2d += (long long)blob.h.size();
  unsigned char* sb = m->spec.as<unsigned char>();
  kp.mu = (const double*)(sb + o_mu); kp.scale = (const double*)(sb + o_scale);
  kp.min_ = (const double*)(sb + o_min); kp.max_ = (const double*)(sb + o_max);
  kp.lb = (const double*)(sb + o_lb); kp.ub = (const double*)(sb + o_ub);
  kp.fixed = sb + o_fixed; kp.free_idx = (const int*)(sb + o_free);
  kp.order = (ks->scheme == FMCMC_SCHEME_EXPLICIT) ? (const int*)(sb + o_order) : nullptr;
  kp.nadapt = ks->nadapt_len > 0 ? (const long long*)(sb + o_nadapt) : nullptr;
  kp.constr = ks->constr ? (const double*)(sb + o_constr) : nullptr;
  kp.seq = fed_seq ? (const int*)(sb + o_seq) : nullptr;

  // ---- stream ----------------------------------------------------------------------------
  StreamParams sp{};
  sp.mode = stream->mode; sp.kdraw = stream->kdraw; sp.seed = stream->seed; sp.run = (unsigned int)stream->run_index;
  if (stream->mode == FMCMC_STREAM_FED) {
    CU_CHECK(ensure(m->fed_logu, (size_t)C * T * 8));
    CU_CHECK(ensure(m->fed_z, (size_t)C * T * stream->kdraw * 8));
    CU_CHECK(cudaMemcpyAsync(m->fed_logu.p, stream->logu, (size_t)C * T * 8, cudaMemcpyHostToDevice, m->stream));
    CU_CHECK(cudaMemcpyAsync(m->fed_z.p, stream->z, (size_t)C * T * stream->kdraw * 8, cudaMemcpyHostToDevice, m->stream));
    h2d += (long long)C * T * 8 * (1 + stream->kdraw);
    sp.logu = m->fed_logu.as<double>();
    sp.z = m->fed_z.as<double>();
  }

  // ---- run buffers ---------------------------------------------------------------------
  const bool is_ram = ks->type == FMCMC_KERNEL_RAM;
  const long long worklen = is_ram ? 4LL * kf * kf
                                  : (ks->type == FMCMC_KERNEL_ADAPT ? (ks->mvn_method == FMCMC_MVN_EIGEN ? 3LL : 1LL) * kf * kf : 0);  // L | eigen: A, V
  CU_CHECK(ensure(m->ans, (size_t)T * C * k * 8));
  CU_CHECK(ensure(m->draws, (size_t)T * C * k * 8));
  CU_CHECK(ensure(m->logpost, (size_t)T * C * 8));
  if (m->state_nchains != C || m->state_k != k) {
    // shape change: state buffers are reallocated (grow-only helper keeps old data otherwise)
    m->state_nchains = 0;
  }
  CU_CHECK(ensure(m->cur_theta, (size_t)C * k * 8));
  CU_CHECK(ensure(m->cur_f, (size_t)C * 8));
  CU_CHECK(ensure(m->prop, (size_t)C * k * 8));
  CU_CHECK(ensure(m->prop_u, (size_t)C * k * 8));
  CU_CHECK(ensure(m->istate, (size_t)C * FMCMC_ISTATE_LEN * 8));
  CU_CHECK(ensure(m->dstate, (size_t)C * (dlen ? dlen : 1) * 8));
  CU_CHECK(ensure(m->colsum, (size_t)C * kf * 16));
  CU_CHECK(ensure(m->ubuf, (size_t)C * kf * 8));
  CU_CHECK(ensure(m->work, (size_t)C * (worklen ? worklen : 1) * 8));
  CU_CHECK(ensure(m->cflags, (size_t)C * sizeof(int)));
  CU_CHECK(cudaMemsetAsync(m->errbuf.p, 0, 4 * sizeof(int), m->stream));
  CU_CHECK(cudaMemsetAsync(m->nacc.p, 0, sizeof(unsigned long long), m->stream));
  CU_CHECK(cudaMemsetAsync(m->cflags.p, 0, (size_t)C * sizeof(int), m->stream));
  if (dev_state) {
    // resident: nothing to move
  } else if (state && state->istate) {
    CU_CHECK(cudaMemcpyAsync(m->istate.p, state->istate, (size_t)C * FMCMC_ISTATE_LEN * 8, cudaMemcpyHostToDevice, m->stream));
    h2d += (long long)C * FMCMC_ISTATE_LEN * 8;
  } else {
    CU_CHECK(cudaMemsetAsync(m->istate.p, 0, (size_t)C * FMCMC_ISTATE_LEN * 8, m->stream));
  }
  if (dlen && !dev_state) {
    if (state && state->dstate) {
      CU_CHECK(cudaMemcpyAsync(m->dstate.p, state->dstate, (size_t)C * dlen * 8, cudaMemcpyHostToDevice, m->stream));
      h2d += (long long)C * dlen * 8;
    } else {
      CU_CHECK(cudaMemsetAsync(m->dstate.p, 0, (size_t)C * dlen * 8, m->stream));
    }
  }
  const double* d_initial = nullptr;
  if (run->initial) {
    CU_CHECK(ensure(m->initial, (size_t)C * k * 8));
    CU_CHECK(cudaMemcpyAsync(m->initial.p, run->initial, (size_t)C * k * 8, cudaMemcpyHostToDevice, m->stream));
    d_initial = m->initial.as<double>();
    h2d += (long long)C * k * 8;
  }
  RunBuffers rb{};
  rb.nchains = C; rb.chain_offset = run->chain_offset; rb.T = T;
  rb.ans = m->ans.as<double>(); rb.draws = m->draws.as<double>(); rb.logpost = m->logpost.as<double>();
  rb.cur_theta = m->cur_theta.as<double>(); rb.cur_f = m->cur_f.as<double>();
  rb.prop = m->prop.as<double>(); rb.prop_u = m->prop_u.as<double>();
  rb.istate = m->istate.as<long long>(); rb.dstate = m->dstate.as<double>();
  rb.colsum = m->colsum.as<double>(); rb.ubuf = m->ubuf.as<double>();
  rb.work = m->work.as<double>(); rb.worklen = worklen;
  rb.chain_flags = m->cflags.as<int>(); rb.err = m->errbuf.as<int>();
  rb.n_accept = m->nacc.as<unsigned long long>();

  // ---- choose the stepping path --------------------------------------------------------
  // 1 = chain-resident fused kernel; 2 = observation-tiled, lane<->chain DFMA kernel (p_x <= 32);
  // 3 = observation-tiled, DMMA kernel (p_x <= 128)
  const ModelParams& mp = m->mp;
  const size_t data_bytes = (size_t)mp.p_x * mp.ld * 8 + (size_t)mp.ld * 8 + (mp.group ? (size_t)mp.ld * 4 + 16 : 0);
  const bool lm_or_logit = mp.family == FMCMC_FAMILY_GAUSSIAN_LM || mp.family == FMCMC_FAMILY_LOGISTIC;
  const bool tiled_ok = lm_or_logit && mp.p_x <= 128;
  int path = m->forced_path;
  if (m->shard_world > 1) path = 3;
  if (path == 0) {
    // narrow X: the DFMA kernel's 8 / 16-column tiers do no padded work.  Otherwise, many likelihood columns (chains):
    // the split-integer tcgen05 kernel (path 4), whose FP64 pipe only runs the family epilogue; few columns: the
    // DMMA kernel's observation-split mapping, which is HBM-bound (tiled_mma.cuh)
    const long long ctot = run->nchains_total > C ? run->nchains_total : C;  // the whole job's chains: sharding must not change the path
    const long long ncols_auto = is_ram ? 2 * ctot : ctot;
    path = (tiled_ok && data_bytes > 96 * 1024)
               ? (mp.p_x <= 16 ? 2 : ((ncols_auto > 128 && m->tiled_default != 2) ? m->tiled_many : (mp.p_x > 32 ? 3 : m->tiled_default)))
               : 1;
  }
  if (m->trimmed_to && path != m->trimmed_to) {
    set_err(err, errlen, "the model was trimmed to stepping path %d (fmcmc_model_trim); this run needs path %d", m->trimmed_to, path);
    return FMCMC_EUNSUP;
  }
  if ((path == 2 && !(lm_or_logit && mp.p_x <= 32)) || ((path == 3 || path == 4) && !tiled_ok)) {
    set_err(err, errlen, "the observation-tiled paths support gaussian_lm / logistic with p_x <= 32 (path 2) or <= 128 (paths 3, 4); got family %d, p_x %d", mp.family, mp.p_x);
    return FMCMC_EUNSUP;
  }
  // 6 slices leave ~1e-12 |theta x|max in eta: averaged over >= 65536 observations the log-posterior is ~1e-14 relative.
  // kernel_ram (its adaptation consumes f itself) and short data (less averaging, and the tensor work is negligible
  // there anyway) run on 7 slices: ~1e-14 in eta, ~1e-15 relative in f.
  const int i8_NS = m->i8_slices ? m->i8_slices : ((is_ram || mp.n_total < 65536) ? I8_NS_LO + 1 : I8_NS_LO), i8_KB = i8_kblocks(mp.p_x);
  if (m->trimmed_to == 4 && (m->xq_NS != i8_NS || m->xq_KB != i8_KB)) {
    set_err(err, errlen, "the model was trimmed to path 4 with %d int8 slices; this run needs %d (kernel_ram / short data use one more) and X is gone", m->xq_NS, i8_NS);
    return FMCMC_EUNSUP;
  }
  if (path == 4) {  // int8 slice tiles of X (once per model); X with non-finite entries cannot be sliced
    cudaError_t pe = ensure_packed_i8(m, i8_NS, i8_KB);
    if (pe == cudaErrorNotSupported) {
      if (m->forced_path == 4) { set_err(err, errlen, "path 4 (split-integer tensor cores) needs finite X with |x| < 2^480"); return FMCMC_EUNSUP; }
      path = 3;
    } else if (pe == cudaErrorMemoryAllocation) {
      set_err(err, errlen, "out of device memory for the int8 slice tiles of X");
      return FMCMC_ENOMEM;
    } else if (pe != cudaSuccess) {
      set_err(err, errlen, "CUDA error %s (pack_i8)", cudaGetErrorString(pe));
      return FMCMC_ECUDA;
    }
  }
  long long launches = 0;
  long long stream_chunk = 0, stream_nchunks = 0;   // streamed outputs (tiled paths): kept rows per chunk, chunks
  int hot_timed = 0;
  CU_CHECK(cudaEventRecord(m->ev0, m->stream));
  if (path == 1) {
    // ---- path 1: chain-resident fused kernel, one launch per bulk -------------------------
    const bool wpc = C >= 2 * m->sm_count;
    // shared-memory matrix scratch of the adaptive kernels (4 kf^2 doubles per chain) when it is small enough
    const bool adaptive = ks->type == FMCMC_KERNEL_ADAPT || ks->type == FMCMC_KERNEL_RAM;
    const size_t mat_bytes = adaptive ? (size_t)4 * kf * kf * 8 : 0;
    const int mat_doubles = (mat_bytes && mat_bytes <= (wpc ? (size_t)8 * 1024 : (size_t)40 * 1024)) ? 4 * kf * kf : 0;
    const int chain_smem_doubles = 7 * k + mat_doubles;
    int threads, chains_per_block;
    if (wpc) { threads = 256; chains_per_block = threads / 32; }
    else {
      long long want = (mp.n + 3) / 4;
      threads = (int)std::min<long long>(256, std::max<long long>(32, ((want + 31) / 32) * 32));
      if (const char* v = getenv("FMCMC_RES_THREADS")) { const int t = atoi(v); if (t >= 32 && t <= 256 && t % 32 == 0) threads = t; }  // tuning only
      chains_per_block = 1;
    }
    const size_t chain_bytes = (size_t)chains_per_block * chain_smem_doubles * 8 + 2 * RES_MAX_WARPS * 8 + 128;
    const bool in_smem = data_bytes + chain_bytes + 64 <= (size_t)m->smem_optin;
    const size_t smem = chain_bytes + (in_smem ? data_bytes + 16 : 0);
    if (smem > (size_t)m->smem_optin) {
      set_err(err, errlen, "k = %d needs %zu bytes of shared memory per CTA (limit %d)", k, smem, m->smem_optin);
      return FMCMC_EUNSUP;
    }
    const int blocks = (C + chains_per_block - 1) / chains_per_block;
#define RES_LAUNCH(W, K)                                                                                         \
  do {                                                                                                           \
    CU_CHECK(cudaFuncSetAttribute(mh_resident_kernel<W, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    mh_resident_kernel<W, K><<<blocks, threads, smem, m->stream>>>(mp, kp, sp, rb, d_initial, in_smem ? 1 : 0,   \
                                                                   chain_smem_doubles, mat_doubles);             \
  } while (0)
#define RES_DISPATCH(W)                                                   \
  switch (kernel_class(ks->type)) {                                      \
    case KC_ADAPT: RES_LAUNCH(W, KC_ADAPT); break;                       \
    case KC_RAM: RES_LAUNCH(W, KC_RAM); break;                           \
    case KC_MIRROR: RES_LAUNCH(W, KC_MIRROR); break;                     \
    default: RES_LAUNCH(W, KC_PLAIN); break;                             \
  }
    if (wpc) { RES_DISPATCH(true) } else { RES_DISPATCH(false) }
#undef RES_DISPATCH
#undef RES_LAUNCH
    CU_CHECK(cudaGetLastError());
    launches += 1;
  } else {
    // ---- path 2: observation-tiled, two launches per row ------------------------------------
    const int PB = mp.p_x <= 8 ? 8 : (mp.p_x <= 16 ? 16 : 32);
    const int ncols_all = is_ram ? 2 * C : C;
    const MmaShape msh = mma_shape(mp.p_x, ncols_all, m->mma_wide);
    const int cpb = path == 4 ? I8_CHAINS : (path == 3 ? (msh.osplit ? msh.NT * 8 : msh.warps * msh.NT * 8) : TL_CHAINS);   // chains per CTA
    const int tile_rows = path == 4 ? i8_tile_rows(mp.family, i8_NS, i8_KB) : (path == 3 ? (msh.PB <= 32 ? 128 : (msh.PB == 64 ? 64 : 32)) : TL_TILE);
    if (path == 3) {
      cudaError_t pe = ensure_packed_tiles(m, msh.PB);
      if (pe == cudaErrorMemoryAllocation) { set_err(err, errlen, "out of device memory for the tile-major copy of X"); return FMCMC_ENOMEM; }
      if (pe != cudaSuccess) { set_err(err, errlen, "CUDA error %s (pack_tiles)", cudaGetErrorString(pe)); return FMCMC_ECUDA; }
    }
    TiledBuffers tb{};
    tb.ncols = is_ram ? 2 * C : C;
    tb.exact_core = is_ram ? 1 : 0;
    const int chain_blocks = (tb.ncols + cpb - 1) / cpb;
    const long long ntiles = (mp.ld + tile_rows - 1) / tile_rows;
    // One observation slice per SM, whatever the number of chains: chain_blocks waves of sm_count CTAs.  Keeping
    // gx independent of C makes the summation tree - hence every log-posterior bit - independent of how the
    // chains are sharded over calls / GPUs (as long as the same mapping, chain- or observation-split, is used).
    int gx = m->sm_count;
    if (gx > ntiles) gx = (int)ntiles;
    tb.gx = gx;
    tb.cb = chain_blocks;
#ifdef FMCMC_I8_TUNE_HOOKS
    if (const char* v = getenv("FMCMC_I8_TUNE")) tb.tune = atoi(v);
#endif
    const bool sharded = m->shard_world > 1;
    if (sharded) {  // observation sharding: partial sums live in the exchange buffer, one block per step parity
      if (path != 3 || gx != m->sm_count || tb.ncols > m->shard_max_cols) {
        set_err(err, errlen, "observation sharding needs the DMMA path, >= %d observation tiles per rank and <= %d likelihood columns (got path %d, %d slices, %d columns)",
                m->sm_count, m->shard_max_cols, path, gx, tb.ncols);
        return FMCMC_EINVAL;
      }
      tb.sx.world = m->shard_world; tb.sx.rank = m->shard_rank; tb.sx.done = m->shard_done;
      tb.sx.parity_stride = (long long)m->shard_world * m->sm_count * m->shard_max_cols;
      for (int g = 0; g < m->shard_world; g++) { tb.sx.peer_partial[g] = m->shard_peer_partial[g]; tb.sx.peer_flags[g] = m->shard_peer_flags[g]; }
      tb.gx_total = m->shard_world * gx;
      tb.partial = m->shard_partial;
    } else {
      CU_CHECK(ensure(m->partial, (size_t)gx * tb.ncols * 8));
      tb.partial = m->partial.as<double>();
      tb.gx_total = gx;
    }
    // path 4: a CTA walks gsl consecutive slices (flushed one by one: the partial sums, hence every bit, do not depend on gsl) so
    // that its set-up - launch, tensor-memory allocation, Theta slicing, the table's bulk copy, pipeline fill / drain - is paid
    // once per gsl slices; the CTA count stays a whole number of waves (gx / gsl * chain_blocks, gsl dividing both)
    int gsl = 1;
    if (path == 4 && m->i8_gsl != 1)
      for (int g = (m->i8_gsl > 0 ? m->i8_gsl : 4); g > 1; g >>= 1)
        if (gx % g == 0 && chain_blocks % g == 0) { gsl = g; break; }
    tb.gsl = gsl;
    // path 3 with one chain block (the few-chain, HBM-bound regime) on data larger than L2: keep a part of the tile-major X
    // L2-resident from row to row (tiled_mma.cuh); FMCMC_L2_KEEP_MB overrides the size (0 disables; A/B measurements)
    if (path == 3 && chain_blocks == 1 && !sharded) {
      const size_t stage_bytes = (msh.PB <= 32 ? MmaGeom<32>::STAGE_DOUBLES : (msh.PB == 64 ? MmaGeom<64>::STAGE_DOUBLES : MmaGeom<128>::STAGE_DOUBLES)) * sizeof(double), total = (size_t)ntiles * stage_bytes;
      size_t keep_bytes = (size_t)(0.33 * m->l2_bytes);   // B200, n = 1e6, p = 32, 4 chains: 0 / 40 / 60 / 75 / 90 / 110 MB -> 68.7 / 64.7 / 65.1 / 65.3 / 68.1 / 69.8 us per MH step
      if (m->l2_keep_mb >= 0) keep_bytes = (size_t)m->l2_keep_mb << 20;
      if (total > (size_t)m->l2_bytes && keep_bytes > 0) tb.l2_keep_tiles = (long long)(keep_bytes / stage_bytes);
    }
    const dim3 lgrid = path >= 3 ? dim3((unsigned)(gx / gsl) * chain_blocks, 1) : dim3(gx, chain_blocks);
    const int hblocks = (C + TL_HEAD_WARPS - 1) / TL_HEAD_WARPS;
    const bool adaptive = ks->type == FMCMC_KERNEL_ADAPT || ks->type == FMCMC_KERNEL_RAM;
    // kernel_adapt factorises in 2 kf^2 doubles of shared memory (A, L), kernel_ram needs 4 kf^2: at k = 32 that is 16 KB per
    // warp instead of 32, so three CTAs instead of one share an SM and 1 024 chains are one wave of the head kernel, not two
    const int mat_mats = ks->type == FMCMC_KERNEL_ADAPT ? 2 : 4;
    const size_t mat_bytes = adaptive ? (size_t)mat_mats * kf * kf * 8 : 0;
    const int mat_doubles = (mat_bytes && mat_bytes <= (size_t)40 * 1024) ? mat_mats * kf * kf : 0;
    const size_t hsmem = (size_t)TL_HEAD_WARPS * (4 * k + mat_doubles) * 8;
    const int kclass = kernel_class(ks->type);
    // Programmatic dependent launch chains head(row) -> likelihood(row) -> head(row + 1) ...: each kernel's CTAs may set
    // themselves up while the previous one drains (pdl_wait() in the kernels orders the data).  Not when sharded over
    // observations (the peer-flag protocol paces the kernels there), and not across the event records of a timed launch.
    const bool pdl_ok = m->pdl && !(m->shard_world > 1);
    bool pdl_head = false;
    // few chains, kernel_adapt with the Cholesky draw: a CTA per chain (tiled.cuh, tiled_head_adapt_cta_kernel) - same results bit for bit
    // (with more chains than SMs the same kernel at 128 threads per chain - 8 chains per CTA - was measured SLOWER than the
    // warp-per-chain head at 1 024 chains, 58 against 33 us per row: profiles/r02_findings.md, section 20)
    const bool head_cta = m->head_cta && kclass == KC_ADAPT && ks->mvn_method != FMCMC_MVN_EIGEN && ks->bw <= 0 && kf <= 32 && C <= m->sm_count;
    if (head_cta) {
      cudaError_t ae = cudaFuncSetAttribute(tiled_head_adapt_cta_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled_headc_smem_bytes(1024));
      if (ae != cudaSuccess) { set_err(err, errlen, "CUDA error %s (tiled_head_adapt_cta attributes)", cudaGetErrorString(ae)); return FMCMC_ECUDA; }
    }
    auto head_launch = [&](long long row) -> cudaError_t {
      if (head_cta)
        return launch_chained(tiled_head_adapt_cta_kernel<1024>, dim3(C), dim3(TL_HEADC_THREADS), tiled_headc_smem_bytes(1024), m->stream, pdl_head,
                              mp, kp, sp, rb, tb, d_initial, row);
#define HEAD_CASE(K)                                                                                              \
  case K:                                                                                                          \
    if (hsmem > 48 * 1024) {                                                                                       \
      cudaError_t e_ = cudaFuncSetAttribute(tiled_head_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem); \
      if (e_ != cudaSuccess) return e_;                                                                            \
    }                                                                                                              \
    { cudaError_t e_ = launch_chained(tiled_head_kernel<K>, dim3(hblocks), dim3(TL_HEAD_WARPS * 32), hsmem, m->stream, pdl_head, \
                                      mp, kp, sp, rb, tb, d_initial, row, mat_doubles);                            \
      if (e_ != cudaSuccess) return e_; }                                                                          \
    break;
      switch (kclass) {
        HEAD_CASE(KC_ADAPT)
        HEAD_CASE(KC_RAM)
        HEAD_CASE(KC_MIRROR)
        default:
        HEAD_CASE(KC_PLAIN)
      }
#undef HEAD_CASE
      return cudaSuccess;
    };
    if (m->hot_ev.empty()) {
      m->hot_ev.resize(2 * FM_HOT_EVENTS);
      for (auto& e : m->hot_ev) CU_CHECK(cudaEventCreate(&e));
    }
    // Streamed outputs: the stepping launches are all queued first; while the GPU works through them the host copies
    // the kept rows of finished chunks (gather on a second stream + strided D2H) - only the last chunk's copy is exposed.
    {
      const bool want_out = !(run->flags & FMCMC_RUN_NO_OUTPUT) && keep > 0 && (ans_out || draws_out || logpost_out);
      const size_t row_bytes = (size_t)C * k * 8 * ((ans_out ? 1 : 0) + ((draws_out && !(run->flags & FMCMC_RUN_NO_DRAWS)) ? 1 : 0)) +
                               (logpost_out ? (size_t)C * 8 : 0);
      if (want_out && row_bytes * (size_t)keep >= ((size_t)8 << 20)) {
        stream_chunk = (long long)std::max<size_t>(1, ((size_t)4 << 20) / std::max<size_t>(row_bytes, 1));
        stream_nchunks = (keep + stream_chunk - 1) / stream_chunk;
        if (stream_nchunks < 2) stream_chunk = stream_nchunks = 0;
      }
      if (stream_nchunks) {
        if (ans_out) CU_CHECK(ensure(m->out_ans, (size_t)C * keep * k * 8));
        if (draws_out && !(run->flags & FMCMC_RUN_NO_DRAWS)) CU_CHECK(ensure(m->out_draws, (size_t)C * keep * k * 8));
        if (logpost_out) CU_CHECK(ensure(m->out_lp, (size_t)C * keep * 8));
        while ((long long)m->chunk_ev.size() < stream_nchunks) {
          cudaEvent_t cev;
          CU_CHECK(cudaEventCreateWithFlags(&cev, cudaEventDisableTiming));
          m->chunk_ev.push_back(cev);
        }
      }
    }
    long long next_chunk = 0, hot_seen = 0;
    for (long long row = 1; row <= T + 1; row++) {
      bool head_direct = true;   // no event record between this row's head kernel and its likelihood launch
      { cudaError_t he = head_launch(row); if (he != cudaSuccess) { set_err(err, errlen, "CUDA error %s (tiled_head)", cudaGetErrorString(he)); return FMCMC_ECUDA; } }
      launches += 1;
      if (next_chunk < stream_nchunks) {  // head(row) finalises source row row - 2 (0-based): is chunk `next_chunk` complete?
        const long long r1 = std::min<long long>(keep, (next_chunk + 1) * stream_chunk);
        const long long last_src = run->burnin + r1 * run->thin - 1;
        if (row - 2 >= last_src) { cudaEventRecord(m->chunk_ev[next_chunk], m->stream); next_chunk++; head_direct = false; }
      }
      if (row <= T && !(row == 1 && !d_initial)) {  // f(theta0) of a continued run is already on the device
        // the hot kernel is bracketed by events on every 16th launch of a call (the first included): a timed launch runs
        // alone - its time is the kernel's own -, the others overlap their set-up with the head kernel
        const bool timed = hot_timed < FM_HOT_EVENTS && (hot_seen++ % FM_HOT_EVERY) == 0;
        const bool pdl = pdl_ok && !timed && head_direct;
        if (sharded) {
          tb.sx.step = ++m->shard_step;
          tb.partial = m->shard_partial + (size_t)(tb.sx.step & 1ULL) * tb.sx.parity_stride;
        }
        if (timed) cudaEventRecord(m->hot_ev[2 * hot_timed], m->stream);
        cudaError_t e;
        if (path == 4)
          e = (mp.family == FMCMC_FAMILY_LOGISTIC)
                  ? (mp.y_binary ? launch_tiled_i8<FMCMC_FAMILY_LOGISTIC, true>(m, i8_NS, i8_KB, lgrid, rb, tb, pdl)
                                 : launch_tiled_i8<FMCMC_FAMILY_LOGISTIC, false>(m, i8_NS, i8_KB, lgrid, rb, tb, pdl))
                  : launch_tiled_i8<FMCMC_FAMILY_GAUSSIAN_LM, false>(m, i8_NS, i8_KB, lgrid, rb, tb, pdl);
        else if (path == 3)
          e = (mp.family == FMCMC_FAMILY_LOGISTIC)
                  ? (mp.y_binary ? launch_tiled_mma<FMCMC_FAMILY_LOGISTIC, true>(m, msh, lgrid, rb, tb, pdl)
                                 : launch_tiled_mma<FMCMC_FAMILY_LOGISTIC, false>(m, msh, lgrid, rb, tb, pdl))
                  : launch_tiled_mma<FMCMC_FAMILY_GAUSSIAN_LM, false>(m, msh, lgrid, rb, tb, pdl);
        else
          e = (mp.family == FMCMC_FAMILY_LOGISTIC)
                  ? (mp.y_binary ? launch_tiled_loglik<FMCMC_FAMILY_LOGISTIC, true>(m, PB, lgrid, rb, tb, pdl)
                                 : launch_tiled_loglik<FMCMC_FAMILY_LOGISTIC, false>(m, PB, lgrid, rb, tb, pdl))
                  : launch_tiled_loglik<FMCMC_FAMILY_GAUSSIAN_LM, false>(m, PB, lgrid, rb, tb, pdl);
        if (e != cudaSuccess) { set_err(err, errlen, "CUDA launch error %s (tiled_loglik)", cudaGetErrorString(e)); return FMCMC_ECUDA; }
        if (timed) cudaEventRecord(m->hot_ev[2 * hot_timed + 1], m->stream), hot_timed++;
        pdl_head = pdl;   // the next head kernel follows this launch directly unless an event record sits between them
        launches += 1;
      }
    }
    CU_CHECK(cudaGetLastError());
  }
  CU_CHECK(cudaEventRecord(m->ev1, m->stream));

  if (stream_nchunks) {  // overlap: chunk q leaves while the rows after it are still being computed
    const int gb = m->sm_count;
    const bool want_draws = draws_out && !(run->flags & FMCMC_RUN_NO_DRAWS);
    for (long long q = 0; q < stream_nchunks; q++) {
      const long long r0 = q * stream_chunk, r1 = std::min<long long>(keep, r0 + stream_chunk);
      CU_CHECK(cudaStreamWaitEvent(m->copy_stream, m->chunk_ev[q], 0));
      // row-major [c][r][k]: chain c's rows r0..r1 are one run of (r1 - r0) k doubles, C runs a pitch of keep k apart;
      // column-major [c][j][r]: (r1 - r0) doubles per (chain, parameter), C k runs a pitch of keep apart
      // (the host arrays may hold out_rows_total >= keep rows per chain, this call's rows starting at out_row_offset: hpitch)
      const int cm = (run->flags & FMCMC_RUN_COLMAJOR) ? 1 : 0;
      const size_t pitch = cm ? (size_t)keep * 8 : (size_t)keep * k * 8;
      const size_t hpitch = cm ? (size_t)host_rows * 8 : (size_t)host_rows * k * 8;
      const size_t width = cm ? (size_t)(r1 - r0) * 8 : (size_t)(r1 - r0) * k * 8;
      const size_t off = cm ? (size_t)r0 : (size_t)r0 * k, height = cm ? (size_t)C * k : (size_t)C;
      const size_t hoff = cm ? (size_t)(host_off + r0) : (size_t)(host_off + r0) * k;
      if (ans_out) {
        gather_rows_range_kernel<<<gb, 256, 0, m->copy_stream>>>(rb.ans, m->out_ans.as<double>(), C, k, keep, r0, r1, run->burnin, run->thin, cm);
        CU_CHECK(cudaMemcpy2DAsync(ans_out + hoff, hpitch, m->out_ans.as<double>() + off, pitch, width, height, cudaMemcpyDeviceToHost, m->copy_stream));
        launches += 1;
      }
      if (want_draws) {
        gather_rows_range_kernel<<<gb, 256, 0, m->copy_stream>>>(rb.draws, m->out_draws.as<double>(), C, k, keep, r0, r1, run->burnin, run->thin, cm);
        CU_CHECK(cudaMemcpy2DAsync(draws_out + hoff, hpitch, m->out_draws.as<double>() + off, pitch, width, height, cudaMemcpyDeviceToHost, m->copy_stream));
        launches += 1;
      }
      if (logpost_out) {
        gather_rows_range_kernel<<<gb, 256, 0, m->copy_stream>>>(rb.logpost, m->out_lp.as<double>(), C, 1, keep, r0, r1, run->burnin, run->thin, 0);
        CU_CHECK(cudaMemcpy2DAsync(logpost_out + host_off + r0, (size_t)host_rows * 8, m->out_lp.as<double>() + r0, (size_t)keep * 8,
                                   (size_t)(r1 - r0) * 8, C, cudaMemcpyDeviceToHost, m->copy_stream));
        launches += 1;
      }
      CU_CHECK(cudaStreamSynchronize(m->copy_stream));
    }
    d2h += (long long)C * keep * k * 8 * ((ans_out ? 1 : 0) + (want_draws ? 1 : 0)) + (logpost_out ? (long long)C * keep * 8 : 0);
  }

  // ---- error check, outputs -----------------------------------------------------------------
  int herr[4] = {0, 0, 0, 0};
  unsigned long long hacc = 0;
  CU_CHECK(cudaMemcpyAsync(herr, m->errbuf.p, sizeof(herr), cudaMemcpyDeviceToHost, m->stream));
  CU_CHECK(cudaMemcpyAsync(&hacc, m->nacc.p, sizeof(hacc), cudaMemcpyDeviceToHost, m->stream));
  CU_CHECK(cudaStreamSynchronize(m->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, m->ev0, m->ev1);
  double hot_ms = 0.0;
  for (int q = 0; q < hot_timed; q++) {
    float t = 0.f;
    cudaEventElapsedTime(&t, m->hot_ev[2 * q], m->hot_ev[2 * q + 1]);
    hot_ms += t;
  }
  d2h += (long long)sizeof(int) * 4 + 8;
  if (report) {
    report->hot_ms = hot_ms;
    report->hot_launches = hot_timed;
    report->device_ms = ms;
    report->n_accept = (int64_t)hacc;
    report->path = path;
    report->n_launches = launches;
  }
  if (herr[0] != 0) {
    m->state_nchains = 0;
    if (report) { report->nan_chain = herr[1]; report->nan_step = herr[2]; }
    switch (herr[0]) {
      case FMCMC_ENAN:
        set_err(err, errlen, "fun(par) is undefined (NaN). Check either -fun- or the -lb- and -ub- parameters. This error ocurred during step i = %d (chain %d).", herr[2], herr[1]);
        break;
      case FMCMC_ENANRATIO:
        set_err(err, errlen, "missing value where TRUE/FALSE needed (f1 - f0 is NaN) at step i = %d (chain %d).", herr[2], herr[1]);
        break;
      case FMCMC_ENOTPD:
        set_err(err, errlen, "'Sigma' is not positive definite (step i = %d, chain %d).", herr[2], herr[1]);
        break;
      case FMCMC_EPEER:
        set_err(err, errlen, "observation sharding: a peer GPU did not publish its partial sums within 60 s (step i = %d)", herr[2]);
        break;
      case FMCMC_EUNSUP:
        set_err(err, errlen, "the kernel reached a state the reference itself mishandles at step i = %d, chain %d (SURVEY App. D: mirror quirk D8 / adapt update range / t. = 0)", herr[2], herr[1]);
        break;
      default:
        set_err(err, errlen, "device error %d at step i = %d (chain %d).", herr[0], herr[2], herr[1]);
    }
    return herr[0];
  }
  m->state_nchains = C;
  m->state_k = k;
  m->state_type = ks->type;
  m->state_kf = kf;

  const int gblocks = m->sm_count * 4;
  if (run->flags & FMCMC_RUN_APPEND) {
    if (m->store_C != C || m->store_k != k || m->store_rows + keep > m->store_cap) {
      set_err(err, errlen, "sample store not reserved for %d chains x %d params x %lld more rows (call fmcmc_store_reset)", C, k, (long long)keep);
      return FMCMC_EINVAL;
    }
    if (keep > 0) {
      append_rows_kernel<<<gblocks, 256, 0, m->stream>>>(rb.ans, m->store.as<double>() + (size_t)m->store_rows * C * k,
                                                         (long long)C * k, keep, run->burnin, run->thin);
      launches += 1;
    }
    m->store_rows += keep;
  }
  if (!(run->flags & FMCMC_RUN_NO_OUTPUT) && keep > 0 && !stream_nchunks) {
    const int colmajor = (run->flags & FMCMC_RUN_COLMAJOR) ? 1 : 0;
    // device staging [c][keep][k] (or [c][k][keep]) -> host arrays of host_rows rows per chain, this call's at host_off
    const size_t dpitch = colmajor ? (size_t)keep * 8 : (size_t)keep * k * 8;
    const size_t hpitch = colmajor ? (size_t)host_rows * 8 : (size_t)host_rows * k * 8;
    const size_t hoff = colmajor ? (size_t)host_off : (size_t)host_off * k, height = colmajor ? (size_t)C * k : (size_t)C;
    if (ans_out) {
      CU_CHECK(ensure(m->out_ans, (size_t)C * keep * k * 8));
      gather_rows_kernel<<<gblocks, 256, 0, m->stream>>>(rb.ans, m->out_ans.as<double>(), C, k, keep, run->burnin, run->thin, colmajor);
      CU_CHECK(cudaMemcpy2DAsync(ans_out + hoff, hpitch, m->out_ans.p, dpitch, dpitch, height, cudaMemcpyDeviceToHost, m->stream));
      d2h += (long long)C * keep * k * 8;
      launches += 1;
    }
    if (draws_out && !(run->flags & FMCMC_RUN_NO_DRAWS)) {
      CU_CHECK(ensure(m->out_draws, (size_t)C * keep * k * 8));
      gather_rows_kernel<<<gblocks, 256, 0, m->stream>>>(rb.draws, m->out_draws.as<double>(), C, k, keep, run->burnin, run->thin, colmajor);
      CU_CHECK(cudaMemcpy2DAsync(draws_out + hoff, hpitch, m->out_draws.p, dpitch, dpitch, height, cudaMemcpyDeviceToHost, m->stream));
      d2h += (long long)C * keep * k * 8;
      launches += 1;
    }
    if (logpost_out) {
      CU_CHECK(ensure(m->out_lp, (size_t)C * keep * 8));
      gather_rows_kernel<<<gblocks, 256, 0, m->stream>>>(rb.logpost, m->out_lp.as<double>(), C, 1, keep, run->burnin, run->thin, 0);
      CU_CHECK(cudaMemcpy2DAsync(logpost_out + host_off, (size_t)host_rows * 8, m->out_lp.p, (size_t)keep * 8, (size_t)keep * 8, C,
                                 cudaMemcpyDeviceToHost, m->stream));
      d2h += (long long)C * keep * 8;
      launches += 1;
    }
  }
  const bool keep_state = (run->flags & FMCMC_RUN_KEEP_STATE) != 0;
  if (state && state->istate && !dev_state && !keep_state) {
    CU_CHECK(cudaMemcpyAsync(state->istate, m->istate.p, (size_t)C * FMCMC_ISTATE_LEN * 8, cudaMemcpyDeviceToHost, m->stream));
    d2h += (long long)C * FMCMC_ISTATE_LEN * 8;
  }
  if (state && state->dstate && dlen && !dev_state && !keep_state) {
    CU_CHECK(cudaMemcpyAsync(state->dstate, m->dstate.p, (size_t)C * dlen * 8, cudaMemcpyDeviceToHost, m->stream));
    d2h += (long long)C * dlen * 8;
  }
  CU_CHECK(cudaStreamSynchronize(m->stream));
  CU_CHECK(cudaGetLastError());
  if (report) { report->n_launches = launches; report->h2d_bytes = h2d; report->d2h_bytes = d2h; }
  return FMCMC_OK;
}

// The kernel state the last fmcmc_run left resident on the device -> host (FMCMC_RUN_DEVICE_STATE bulks neither upload
// nor download it; the bulk loop fetches it ONCE, after its last bulk, for the write-back of R/mcmc.R:629-631).
extern "C" int fmcmc_kernel_state_fetch(fmcmc_model* m, fmcmc_kernel_state* state, char* err, size_t errlen) {
  if (!m || !state || !state->istate) { set_err(err, errlen, "null argument"); return FMCMC_EINVAL; }
  if (m->state_nchains < 1) { set_err(err, errlen, "the model holds no kernel state yet (no fmcmc_run so far)"); return FMCMC_EINVAL; }
  CU_CHECK(cudaSetDevice(m->device));
  const int C = m->state_nchains;
  const long long dlen = fmcmc_kernel_state_len(m->state_type, m->state_k, m->state_kf);
  CU_CHECK(cudaMemcpyAsync(state->istate, m->istate.p, (size_t)C * FMCMC_ISTATE_LEN * 8, cudaMemcpyDeviceToHost, m->stream));
  if (state->dstate && dlen > 0)
    CU_CHECK(cudaMemcpyAsync(state->dstate, m->dstate.p, (size_t)C * dlen * 8, cudaMemcpyDeviceToHost, m->stream));
  CU_CHECK(cudaStreamSynchronize(m->stream));
  return FMCMC_OK;
}

// --------------------------------------------------------------------------------
// f(theta) on the device
// --------------------------------------------------------------------------------
extern "C" int fmcmc_logpost(fmcmc_model* m, int32_t nchains, const double* theta, double* out, char* err,
                             size_t errlen) {
  if (!m || !theta || !out || nchains < 1) { set_err(err, errlen, "bad argument"); return FMCMC_EINVAL; }
  if (m->trimmed_to) { set_err(err, errlen, "fmcmc_logpost reads the FP64 copy of X, which fmcmc_model_trim released"); return FMCMC_EUNSUP; }
  CU_CHECK(cudaSetDevice(m->device));
  const int k = m->mp.k;
  CU_CHECK(ensure(m->initial, (size_t)nchains * k * 8));
  CU_CHECK(ensure(m->tmp, (size_t)nchains * 8));
  CU_CHECK(cudaMemcpyAsync(m->initial.p, theta, (size_t)nchains * k * 8, cudaMemcpyHostToDevice, m->stream));
  logpost_kernel<<<nchains, 256, 0, m->stream>>>(m->mp, m->initial.as<double>(), m->tmp.as<double>(), nchains);
  CU_CHECK(cudaGetLastError());
  CU_CHECK(cudaMemcpyAsync(out, m->tmp.p, (size_t)nchains * 8, cudaMemcpyDeviceToHost, m->stream));
  CU_CHECK(cudaStreamSynchronize(m->stream));
  return FMCMC_OK;
}

// --------------------------------------------------------------------------------
// sample store + Gelman-Rubin
// --------------------------------------------------------------------------------
extern "C" int fmcmc_store_reset(fmcmc_model* m, int32_t nchains, int32_t k, int64_t capacity_rows, char* err,
                                 size_t errlen) {
  if (!m || nchains < 1 || k < 1 || capacity_rows < 0) { set_err(err, errlen, "bad argument"); return FMCMC_EINVAL; }
  CU_CHECK(cudaSetDevice(m->device));
  CU_CHECK(ensure(m->store, (size_t)capacity_rows * nchains * k * 8));
  m->store_C = nchains; m->store_k = k; m->store_cap = capacity_rows; m->store_rows = 0;
  return FMCMC_OK;
}
extern "C" int64_t fmcmc_store_rows(const fmcmc_model* m) { return m ? m->store_rows : -1; }

extern "C" int fmcmc_gelman_partials(fmcmc_model* m, int64_t row_begin, int64_t row_end, const uint8_t* free_mask,
                                     double* xbar, double* s2, double* wsum, int dev_out, char* err, size_t errlen) {
  if (!m || !xbar || !s2 || !wsum) { set_err(err, errlen, "null argument"); return FMCMC_EINVAL; }
  CU_CHECK(cudaSetDevice(m->device));
  const int C = m->store_C, k = m->store_k;
  if (row_begin < 0 || row_end > m->store_rows || row_end - row_begin < 2) {
    set_err(err, errlen, "bad window [%lld, %lld) of %lld stored rows", (long long)row_begin, (long long)row_end, (long long)m->store_rows);
    return FMCMC_EINVAL;
  }
  std::vector<int> fidx;
  for (int j = 0; j < k; j++)
    if (!free_mask || free_mask[j]) fidx.push_back(j);
  const int kf = (int)fidx.size();
  if (kf < 1) { set_err(err, errlen, "no free parameters"); return FMCMC_EINVAL; }
  CU_CHECK(ensure(m->g_mask, kf * sizeof(int)));
  CU_CHECK(cudaMemcpyAsync(m->g_mask.p, fidx.data(), kf * sizeof(int), cudaMemcpyHostToDevice, m->stream));
  const bool tiled = kf <= 144;  // moments + tiled SYRK (gelman.cuh); beyond that the per-pair kernel
  const int nblocks = tiled ? std::min(C, m->sm_count) : std::min(C, 4 * m->sm_count);
  CU_CHECK(ensure(m->g_wpart, (size_t)nblocks * kf * kf * 8));
  double *dx = xbar, *ds = s2, *dw = wsum;
  if (!dev_out) {
    CU_CHECK(ensure(m->g_xbar, (size_t)C * kf * 8));
    CU_CHECK(ensure(m->g_s2, (size_t)C * kf * 8));
    CU_CHECK(ensure(m->g_wsum, (size_t)kf * kf * 8));
    dx = m->g_xbar.as<double>(); ds = m->g_s2.as<double>(); dw = m->g_wsum.as<double>();
  }
  if (tiled) {
    const double* st = m->store.as<double>();
    const int* fi = m->g_mask.as<int>();
    double* wp = m->g_wpart.as<double>();
    gelman_chain_moments_kernel<<<std::min(C, 8 * m->sm_count), 256, 2 * 256 * 8, m->stream>>>(st, C, k, row_begin, row_end, fi, kf, dx, ds);
    if (kf <= 16) gelman_syrk_kernel<1><<<nblocks, 256, 0, m->stream>>>(st, C, k, row_begin, row_end, fi, kf, dx, wp);
    else if (kf <= 32) gelman_syrk_kernel<2><<<nblocks, 256, 0, m->stream>>>(st, C, k, row_begin, row_end, fi, kf, dx, wp);
    else if (kf <= 64) gelman_syrk_kernel<4><<<nblocks, 256, 0, m->stream>>>(st, C, k, row_begin, row_end, fi, kf, dx, wp);
    else if (kf <= 128) gelman_syrk_kernel<8><<<nblocks, 256, 0, m->stream>>>(st, C, k, row_begin, row_end, fi, kf, dx, wp);
    else gelman_syrk_kernel<9><<<nblocks, 256, 0, m->stream>>>(st, C, k, row_begin, row_end, fi, kf, dx, wp);
  } else {
    const size_t smem = (size_t)kf * 8;
    gelman_chain_stats_kernel<<<nblocks, 256, smem, m->stream>>>(m->store.as<double>(), C, k, row_begin, row_end,
                                                                 m->g_mask.as<int>(), kf, dx, ds, m->g_wpart.as<double>());
  }
  gelman_wsum_kernel<<<(kf * kf + 255) / 256, 256, 0, m->stream>>>(m->g_wpart.as<double>(), nblocks, kf * kf, dw);
  CU_CHECK(cudaGetLastError());
  if (!dev_out) {
    CU_CHECK(cudaMemcpyAsync(xbar, dx, (size_t)C * kf * 8, cudaMemcpyDeviceToHost, m->stream));
    CU_CHECK(cudaMemcpyAsync(s2, ds, (size_t)C * kf * 8, cudaMemcpyDeviceToHost, m->stream));
    CU_CHECK(cudaMemcpyAsync(wsum, dw, (size_t)kf * kf * 8, cudaMemcpyDeviceToHost, m->stream));
  }
  CU_CHECK(cudaStreamSynchronize(m->stream));
  return FMCMC_OK;
}

// Host finish of coda::gelman.diag's multivariate part: largest eigenvalue of
// L^-1 B L^-T with W = L L' (== backsolve(CW, t(backsolve(CW, B, transpose=TRUE)), transpose=TRUE)).
static int host_sym_eigmax(int p, std::vector<double>& Mx, double* emax);
static int host_gelman_emax(int p, const std::vector<double>& W, const std::vector<double>& B, double* emax) {
  std::vector<double> L((size_t)p * p, 0.0), Y((size_t)p * p), Mx((size_t)p * p);
  for (int j = 0; j < p; j++) {
    double s = W[j + (size_t)j * p];
    for (int q = 0; q < j; q++) s -= L[j + (size_t)q * p] * L[j + (size_t)q * p];
    if (!(s > 0.0)) return 1;
    const double ljj = sqrt(s);
    L[j + (size_t)j * p] = ljj;
    for (int i = j + 1; i < p; i++) {
      double v = W[i + (size_t)j * p];
      for (int q = 0; q < j; q++) v -= L[i + (size_t)q * p] * L[j + (size_t)q * p];
      L[i + (size_t)j * p] = v / ljj;
    }
  }
  for (int c = 0; c < p; c++)
    for (int a = 0; a < p; a++) {
      double s = B[a + (size_t)c * p];
      for (int q = 0; q < a; q++) s -= L[a + (size_t)q * p] * Y[q + (size_t)c * p];
      Y[a + (size_t)c * p] = s / L[a + (size_t)a * p];
    }
  for (int c = 0; c < p; c++)
    for (int a = 0; a < p; a++) {
      double s = Y[c + (size_t)a * p];
      for (int q = 0; q < a; q++) s -= L[a + (size_t)q * p] * Mx[q + (size_t)c * p];
      Mx[a + (size_t)c * p] = s / L[a + (size_t)a * p];
    }
  for (int a = 0; a < p; a++)
    for (int b = 0; b < a; b++) {
      const double v = 0.5 * (Mx[a + (size_t)b * p] + Mx[b + (size_t)a * p]);
      Mx[a + (size_t)b * p] = Mx[b + (size_t)a * p] = v;
    }
  return host_sym_eigmax(p, Mx, emax);
}

// largest eigenvalue of the symmetric p x p Mx (col-major, destroyed): Householder tridiagonalisation (O(4/3 p^3)) + Sturm-count bisection.
static int host_sym_eigmax(int p, std::vector<double>& Mx, double* emax) {
  // (The cyclic Jacobi this replaces took ~40 ms at p = 128; this is ~1 ms.)
  std::vector<double> d(p), e(p > 1 ? p - 1 : 1, 0.0), v(p), w(p);
  for (int c = 0; c + 2 < p; c++) {
    const int nn = p - c - 1;                       // x = Mx[c+1 .., c]
    double nrm = 0.0;
    for (int r = 0; r < nn; r++) nrm += Mx[(c + 1 + r) + (size_t)c * p] * Mx[(c + 1 + r) + (size_t)c * p];
    nrm = sqrt(nrm);
    d[c] = Mx[c + (size_t)c * p];
    const double x0 = Mx[(c + 1) + (size_t)c * p];
    const double alpha = x0 > 0.0 ? -nrm : nrm;
    e[c] = alpha;
    double vn = 0.0;
    for (int r = 0; r < nn; r++) { v[r] = Mx[(c + 1 + r) + (size_t)c * p]; }
    v[0] -= alpha;
    for (int r = 0; r < nn; r++) vn += v[r] * v[r];
    if (!(vn > 0.0)) continue;                      // column already in tridiagonal form
    vn = sqrt(vn);
    for (int r = 0; r < nn; r++) v[r] /= vn;
    double K = 0.0;                                  // A22 <- H A22 H, H = I - 2 v v'
    for (int r = 0; r < nn; r++) {
      double acc = 0.0;
      for (int q = 0; q < nn; q++) acc += Mx[(c + 1 + r) + (size_t)(c + 1 + q) * p] * v[q];
      w[r] = 2.0 * acc;
      K += v[r] * w[r];
    }
    for (int r = 0; r < nn; r++) w[r] -= K * v[r];
    for (int q = 0; q < nn; q++)
      for (int r = 0; r < nn; r++) Mx[(c + 1 + r) + (size_t)(c + 1 + q) * p] -= v[r] * w[q] + w[r] * v[q];
  }
  if (p >= 2) {
    d[p - 2] = Mx[(p - 2) + (size_t)(p - 2) * p];
    e[p - 2] = Mx[(p - 1) + (size_t)(p - 2) * p];
  }
  d[p - 1] = Mx[(p - 1) + (size_t)(p - 1) * p];
  double lo = d[0], hi = d[0];                       // Gershgorin
  for (int a = 0; a < p; a++) {
    const double rad = (a > 0 ? fabs(e[a - 1]) : 0.0) + (a + 1 < p ? fabs(e[a]) : 0.0);
    lo = std::min(lo, d[a] - rad);
    hi = std::max(hi, d[a] + rad);
  }
  const double tiny = 2.2250738585072014e-308;
  auto below = [&](double x) {                       // number of eigenvalues < x
    int cnt = 0;
    double q = d[0] - x;
    if (q < 0.0) cnt++;
    for (int a = 1; a < p; a++) {
      if (fabs(q) < tiny) q = q < 0.0 ? -tiny : tiny;
      q = d[a] - x - e[a - 1] * e[a - 1] / q;
      if (q < 0.0) cnt++;
    }
    return cnt;
  };
  for (int it = 0; it < 200 && hi - lo > 4.0 * 2.220446049250313e-16 * std::max(fabs(lo), fabs(hi)); it++) {
    const double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;
    if (below(mid) >= p) hi = mid; else lo = mid;   // all p eigenvalues below mid -> the largest is below mid
  }
  *emax = 0.5 * (lo + hi);
  return 0;
}

extern "C" int fmcmc_host_sym_eigmax(int32_t p, const double* A, double* emax) {
  if (p < 1 || !A || !emax) return FMCMC_EINVAL;
  std::vector<double> Mx(A, A + (size_t)p * p);
  return host_sym_eigmax(p, Mx, emax);
}

extern "C" int fmcmc_gelman_finish(fmcmc_model* m, int64_t niter, int64_t nchains_total, int32_t kf,
                                   const double* xbar, const double* s2, const double* wsum, int dev_in,
                                   double* psrf, double* mpsrf, char* err, size_t errlen) {
  if (!m || !xbar || !s2 || !wsum || !psrf || !mpsrf || kf < 1) { set_err(err, errlen, "bad argument"); return FMCMC_EINVAL; }
  if (nchains_total < 2) {  // R/convergence.R:239
    set_err(err, errlen, "Convergence test with the Gelman is only available when `nchains` > 1L.");
    return FMCMC_EINVAL;
  }
  CU_CHECK(cudaSetDevice(m->device));
  const double *dx = xbar, *ds = s2, *dw = wsum;
  if (!dev_in) {
    CU_CHECK(ensure(m->g_xbar, (size_t)nchains_total * kf * 8));
    CU_CHECK(ensure(m->g_s2, (size_t)nchains_total * kf * 8));
    CU_CHECK(ensure(m->g_wsum, (size_t)kf * kf * 8));
    CU_CHECK(cudaMemcpyAsync(m->g_xbar.p, xbar, (size_t)nchains_total * kf * 8, cudaMemcpyHostToDevice, m->stream));
    CU_CHECK(cudaMemcpyAsync(m->g_s2.p, s2, (size_t)nchains_total * kf * 8, cudaMemcpyHostToDevice, m->stream));
    CU_CHECK(cudaMemcpyAsync(m->g_wsum.p, wsum, (size_t)kf * kf * 8, cudaMemcpyHostToDevice, m->stream));
    dx = m->g_xbar.as<double>(); ds = m->g_s2.as<double>(); dw = m->g_wsum.as<double>();
  }
  // device: cross-chain moments + between-chain scatter (fixed order); host: O(kf^3) finish
  const int nb = (int)std::min<long long>(nchains_total, 2LL * m->sm_count);
  const size_t nK = (size_t)kf * kf;
  CU_CHECK(ensure(m->tmp, ((size_t)kf * 8 + (size_t)nb * nK + nK) * 8));
  double* d_mom = m->tmp.as<double>();
  double* d_bpart = d_mom + (size_t)kf * 8;
  double* d_b = d_bpart + (size_t)nb * nK;
  gelman_moments_kernel<<<kf, 256, 0, m->stream>>>(dx, ds, (long long)nchains_total, kf, d_mom);
  gelman_between_kernel<<<nb, 256, 0, m->stream>>>(dx, (long long)nchains_total, kf, d_mom, d_bpart);
  gelman_wsum_kernel<<<(int)((nK + 255) / 256), 256, 0, m->stream>>>(d_bpart, nb, (int)nK, d_b);
  CU_CHECK(cudaGetLastError());
  std::vector<double> mom((size_t)kf * 8), B(nK), W(nK);
  CU_CHECK(cudaMemcpyAsync(mom.data(), d_mom, mom.size() * 8, cudaMemcpyDeviceToHost, m->stream));
  CU_CHECK(cudaMemcpyAsync(B.data(), d_b, nK * 8, cudaMemcpyDeviceToHost, m->stream));
  CU_CHECK(cudaMemcpyAsync(W.data(), dw, nK * 8, cudaMemcpyDeviceToHost, m->stream));
  CU_CHECK(cudaStreamSynchronize(m->stream));
  const double N = (double)niter, M = (double)nchains_total;
  for (size_t e = 0; e < nK; e++) { W[e] /= M; B[e] = N * (B[e] / (M - 1.0)); }
  int status = FMCMC_OK;
  *mpsrf = NAN;
  if (kf > 1) {
    double emax = 0.0;
    if (host_gelman_emax(kf, W, B, &emax)) {
      set_err(err, errlen, "chol(W) failed: the within-chain covariance is not positive definite");
      status = FMCMC_ENOTPD;
    } else {
      *mpsrf = sqrt((1.0 - 1.0 / N) + (1.0 + 1.0 / (double)kf) * emax / N);
    }
  }
  for (int a = 0; a < kf; a++) {  // univariate point estimates (coda::gelman.diag)
    const double w = W[a + (size_t)a * kf], b = B[a + (size_t)a * kf];
    const double muhat = mom[(size_t)a * 8], var_s2 = mom[(size_t)a * 8 + 2];
    const double c1 = mom[(size_t)a * 8 + 3], c2 = mom[(size_t)a * 8 + 4];
    const double var_w = var_s2 / M;
    const double var_b = (2.0 * b * b) / (M - 1.0);
    const double cov_wb = (N / M) * (c1 - 2.0 * muhat * c2);
    const double V = (N - 1.0) * w / N + (1.0 + 1.0 / M) * b / N;
    const double var_V = ((N - 1.0) * (N - 1.0) * var_w + (1.0 + 1.0 / M) * (1.0 + 1.0 / M) * var_b +
                          2.0 * (N - 1.0) * (1.0 + 1.0 / M) * cov_wb) / (N * N);
    const double df_V = (2.0 * V * V) / var_V;
    const double df_adj = (df_V + 3.0) / (df_V + 1.0);
    psrf[a] = sqrt(df_adj * ((N - 1.0) / N + (1.0 + 1.0 / M) * (1.0 / N) * (b / w)));
  }
  return status;
}

// rm_invariant's pooled moments of THIS GPU's part of the store (every stored row, local chains, free parameters):
// out = (count, mean, M2 = sum (x - mean)^2).  Several GPUs: combine the triples with Chan's formula (dist.py).
extern "C" int fmcmc_store_pooled(fmcmc_model* m, const uint8_t* free_mask, double* out, char* err, size_t errlen) {
  if (!m || !out) { set_err(err, errlen, "null argument"); return FMCMC_EINVAL; }
  const long long rows = m->store_rows;
  const int C = m->store_C, k = m->store_k;
  if (rows < 1 || C < 1) { set_err(err, errlen, "the sample store is empty"); return FMCMC_EINVAL; }
  std::vector<int> fidx;
  for (int j = 0; j < k; j++)
    if (!free_mask || free_mask[j]) fidx.push_back(j);
  const int kf = (int)fidx.size();
  if (kf < 1) { set_err(err, errlen, "no free parameters"); return FMCMC_EINVAL; }
  CU_CHECK(cudaSetDevice(m->device));
  CU_CHECK(ensure(m->g_mask, kf * sizeof(int)));
  CU_CHECK(cudaMemcpyAsync(m->g_mask.p, fidx.data(), kf * sizeof(int), cudaMemcpyHostToDevice, m->stream));
  const long long total = rows * (long long)C * kf;
  const int nb = (int)std::min<long long>((total + 255) / 256, 8LL * m->sm_count);
  CU_CHECK(ensure(m->g_wpart, (size_t)nb * 3 * 8));
  gelman_pooled_kernel<<<nb, 256, 0, m->stream>>>(m->store.as<double>(), C, k, rows, m->g_mask.as<int>(), kf, m->g_wpart.as<double>());
  CU_CHECK(cudaGetLastError());
  std::vector<double> part((size_t)nb * 3);
  double shift = 0.0;
  CU_CHECK(cudaMemcpyAsync(part.data(), m->g_wpart.p, part.size() * 8, cudaMemcpyDeviceToHost, m->stream));
  CU_CHECK(cudaMemcpyAsync(&shift, m->store.as<double>() + fidx[0], 8, cudaMemcpyDeviceToHost, m->stream));
  CU_CHECK(cudaStreamSynchronize(m->stream));
  double n = 0.0, sd = 0.0, sdd = 0.0;
  for (int b = 0; b < nb; b++) { n += part[(size_t)b * 3]; sd += part[(size_t)b * 3 + 1]; sdd += part[(size_t)b * 3 + 2]; }
  if (n != (double)total) { set_err(err, errlen, "internal: pooled count %.0f != %lld", n, total); return FMCMC_ECUDA; }
  out[0] = n;
  out[1] = shift + sd / n;
  out[2] = sdd - sd * sd / n;
  if (out[2] < 0.0) out[2] = 0.0;
  return FMCMC_OK;
}

// Effective sample size of every (chain, free parameter) series in rows [row_begin, row_end) of the store.
extern "C" int fmcmc_store_ess(fmcmc_model* m, int64_t row_begin, int64_t row_end, const uint8_t* free_mask, int32_t max_lag,
                               double* ess, int32_t* truncated, char* err, size_t errlen) {
  if (!m || !ess) { set_err(err, errlen, "null argument"); return FMCMC_EINVAL; }
  const int C = m->store_C, k = m->store_k;
  const long long N = row_end - row_begin;
  if (row_begin < 0 || row_end > m->store_rows || N < 4) {
    set_err(err, errlen, "bad window [%lld, %lld) of %lld stored rows (at least 4 rows are needed)", (long long)row_begin, (long long)row_end, (long long)m->store_rows);
    return FMCMC_EINVAL;
  }
  std::vector<int> fidx;
  for (int j = 0; j < k; j++)
    if (!free_mask || free_mask[j]) fidx.push_back(j);
  const int kf = (int)fidx.size();
  if (kf < 1) { set_err(err, errlen, "no free parameters"); return FMCMC_EINVAL; }
  if (max_lag <= 0) max_lag = (int)std::min<long long>(N - 1, 2000);
  const int L = (int)std::min<long long>(max_lag, N - 1);
  const size_t smem = ((size_t)N + (size_t)L + 2) * 8;
  CU_CHECK(cudaSetDevice(m->device));
  if (smem > (size_t)m->smem_optin) {
    set_err(err, errlen, "a window of %lld rows with %d lags needs %zu bytes of shared memory per series (limit %d): thin the window", N, L, smem, m->smem_optin);
    return FMCMC_EUNSUP;
  }
  CU_CHECK(cudaFuncSetAttribute(store_ess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CU_CHECK(ensure(m->g_mask, kf * sizeof(int)));
  CU_CHECK(cudaMemcpyAsync(m->g_mask.p, fidx.data(), kf * sizeof(int), cudaMemcpyHostToDevice, m->stream));
  CU_CHECK(ensure(m->g_xbar, (size_t)C * kf * 8 + 16));
  int* d_tr = reinterpret_cast<int*>(m->g_xbar.as<double>() + (size_t)C * kf);
  CU_CHECK(cudaMemsetAsync(d_tr, 0, sizeof(int), m->stream));
  store_ess_kernel<<<C * kf, 256, smem, m->stream>>>(m->store.as<double>(), C, k, row_begin, row_end, m->g_mask.as<int>(), kf, L,
                                                    m->g_xbar.as<double>(), d_tr);
  CU_CHECK(cudaGetLastError());
  int tr = 0;
  CU_CHECK(cudaMemcpyAsync(ess, m->g_xbar.p, (size_t)C * kf * 8, cudaMemcpyDeviceToHost, m->stream));
  CU_CHECK(cudaMemcpyAsync(&tr, d_tr, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  CU_CHECK(cudaStreamSynchronize(m->stream));
  if (truncated) *truncated = tr;
  return FMCMC_OK;
}

// 0-based first store row kept by coda::gelman.diag's autoburnin on an mcmc.list with mcpar = (start, start + (rows-1) thin, thin):
//   if (autoburnin && start(x) < end(x)/2) x <- window(x, start = end(x)/2 + 1)
// window.mcmc() snaps a start that is not on the iteration grid UP to the next kept iteration (ts.eps = 1e-5), then
// keeps rows from trunc((start' - start)/thin + 1.5) (1-based).  Third-party coda, restated from its published source;
// the Python glue's window_first_row (fmcmc_b200/coda.py) is the same arithmetic and tests/test_host_logic.py pins both.
extern "C" int64_t fmcmc_gelman_window_begin(int64_t start_iter, int64_t thin, int64_t rows) {
  if (rows < 1 || thin < 1) return 0;
  const double start = (double)start_iter, th = (double)thin;
  const double end = start + th * (double)(rows - 1);
  if (!(start < end / 2.0)) return 0;
  double ns = end / 2.0 + 1.0;
  if (ns < start) ns = start;
  const double q = (ns - start) / th;
  const double near = start + th * floor(q + 0.5);                  // closest grid point
  if (fabs(near - ns) > fabs(ns) * 1e-5) ns = start + th * (floor(q) + 1.0);  // not on the grid: next kept iteration
  long long first = (long long)((ns - start) / th + 1.5);            // 1-based
  if (first < 1) first = 1;
  if (first > rows) first = rows;
  return first - 1;
}

extern "C" int fmcmc_gelman(fmcmc_model* m, const uint8_t* free_mask, int64_t start_iter, int64_t thin, double* psrf,
                            double* mpsrf, int64_t* niter_used, char* err, size_t errlen) {
  if (!m) { set_err(err, errlen, "null model"); return FMCMC_EINVAL; }
  const long long rows = m->store_rows;
  const long long begin = fmcmc_gelman_window_begin(start_iter, thin, rows);
  if (rows - begin < 2) {
    set_err(err, errlen, "the Gelman-Rubin window holds %lld row(s); at least 2 are needed", rows - begin);
    return FMCMC_EINVAL;
  }
  int kf = 0;
  for (int j = 0; j < m->store_k; j++) kf += (!free_mask || free_mask[j]) ? 1 : 0;
  if (m->store_C < 2) {
    set_err(err, errlen, "Convergence test with the Gelman is only available when `nchains` > 1L.");
    return FMCMC_EINVAL;
  }
  CU_CHECK(cudaSetDevice(m->device));
  CU_CHECK(ensure(m->g_xbar, (size_t)m->store_C * kf * 8));
  CU_CHECK(ensure(m->g_s2, (size_t)m->store_C * kf * 8));
  CU_CHECK(ensure(m->g_wsum, (size_t)kf * kf * 8));
  int rc = fmcmc_gelman_partials(m, begin, rows, free_mask, m->g_xbar.as<double>(), m->g_s2.as<double>(),
                                 m->g_wsum.as<double>(), 1, err, errlen);
  if (rc) return rc;
  if (niter_used) *niter_used = rows - begin;
  return fmcmc_gelman_finish(m, rows - begin, m->store_C, kf, m->g_xbar.as<double>(), m->g_s2.as<double>(),
                             m->g_wsum.as<double>(), 1, psrf, mpsrf, err, errlen);
}

// --------------------------------------------------------------------------------
// exported reference helpers on the device
// --------------------------------------------------------------------------------
__global__ void cov_recursive_kernel(int k, long long rows, const double* X, const double* mean_prev,
                                     const double* cov_prev, double t, double eps, double Sd, const double* Ik,
                                     double* mean_out, double* cov_out, double* scr) {
  const int lane = threadIdx.x;
  double* m = scr;
  double* mp = scr + k;
  for (int a = lane; a < k; a += FM_WARP) mp[a] = mean_prev[a];
  for (int e = lane; e < k * k; e += FM_WARP) cov_out[e] = cov_prev[e];
  __syncwarp();
  for (long long i = 0; i < rows; i++) {
    const double ti = t + (double)i;
    const double* x = X + i * k;
    for (int a = lane; a < k; a += FM_WARP) m[a] = xdiv(xadd(xmul(mp[a], ti), x[a]), ti + 1.0);
    __syncwarp();
    const double c1 = xdiv(ti - 1.0, ti), c2 = xdiv(Sd, ti);
    for (int e = lane; e < k * k; e += FM_WARP) {
      const int a = e % k, b = e / k;
      double inner = xsub(xmul(ti, xmul(mp[a], mp[b])), xmul(ti + 1.0, xmul(m[a], m[b])));
      inner = xadd(inner, xmul(x[a], x[b]));
      inner = xadd(inner, xmul(eps, Ik ? Ik[e] : (a == b ? 1.0 : 0.0)));
      cov_out[e] = xadd(xmul(c1, cov_out[e]), xmul(c2, inner));
    }
    __syncwarp();
    for (int a = lane; a < k; a += FM_WARP) mp[a] = m[a];
    __syncwarp();
  }
  for (int a = lane; a < k; a += FM_WARP) mean_out[a] = mp[a];
}

extern "C" int fmcmc_cov_recursive(int device, int32_t k, int64_t rows, const double* X, const double* mean_prev,
                                   const double* cov_prev, double t, double eps, double Sd, const double* Ik,
                                   double* mean_out, double* cov_out, char* err, size_t errlen) {
  if (k < 1 || rows < 1 || !X || !mean_prev || !cov_prev || !mean_out || !cov_out) { set_err(err, errlen, "bad argument"); return FMCMC_EINVAL; }
  CU_CHECK(cudaSetDevice(device));
  const size_t nX = (size_t)rows * k, nK = (size_t)k * k;
  double* d = nullptr;
  const size_t total = nX + k + nK + nK + k + nK + 2 * (size_t)k;
  CU_CHECK(cudaMalloc(&d, total * 8));
  double *dX = d, *dm = dX + nX, *dc = dm + k, *dI = dc + nK, *dmo = dI + nK, *dco = dmo + k, *dscr = dco + nK;
  cudaMemcpy(dX, X, nX * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dm, mean_prev, k * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dc, cov_prev, nK * 8, cudaMemcpyHostToDevice);
  if (Ik) cudaMemcpy(dI, Ik, nK * 8, cudaMemcpyHostToDevice);
  cov_recursive_kernel<<<1, 32>>>(k, rows, dX, dm, dc, t, eps, Sd, Ik ? dI : nullptr, dmo, dco, dscr);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(mean_out, dmo, k * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(cov_out, dco, nK * 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  CU_CHECK(e);
  return FMCMC_OK;
}

__global__ void reflect_kernel(int k, long long count, double* x, const double* lb, const double* ub,
                               const unsigned char* which) {
  const long long total = count * k;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % k);
    if (which && !which[j]) continue;
    x[e] = reflect1(x[e], lb[j], ub[j]);
  }
}

extern "C" int fmcmc_reflect(int device, int32_t k, int64_t count, double* x, const double* lb, const double* ub,
                             const uint8_t* which, char* err, size_t errlen) {
  if (k < 1 || count < 1 || !x || !lb || !ub) { set_err(err, errlen, "bad argument"); return FMCMC_EINVAL; }
  CU_CHECK(cudaSetDevice(device));
  double* d = nullptr;
  unsigned char* dw = nullptr;
  const size_t n = (size_t)count * k;
  CU_CHECK(cudaMalloc(&d, (n + 2 * (size_t)k) * 8));
  CU_CHECK(cudaMalloc(&dw, k));
  cudaMemcpy(d, x, n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + n, lb, k * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + n + k, ub, k * 8, cudaMemcpyHostToDevice);
  if (which) cudaMemcpy(dw, which, k, cudaMemcpyHostToDevice);
  reflect_kernel<<<148, 256>>>(k, count, d, d + n, d + n + k, which ? dw : nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(x, d, n * 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  cudaFree(dw);
  CU_CHECK(e);
  return FMCMC_OK;
}

// --------------------------------------------------------------------------------
// FP64 FMA peak (measurement helper for bench.py; SURVEY 8d: the binding roof)
// --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, double a, double b, int iters) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1e-9, x2 = x0 + 2e-9, x3 = x0 + 3e-9, x4 = x0 + 4e-9, x5 = x0 + 5e-9,
         x6 = x0 + 6e-9, x7 = x0 + 7e-9;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

extern "C" int fmcmc_measure_fp64_peak(int device, double* dfma_tflops, char* err, size_t errlen) {
  if (!dfma_tflops) { set_err(err, errlen, "null argument"); return FMCMC_EINVAL; }
  CU_CHECK(cudaSetDevice(device));
  int sms = 0;
  CU_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double* d = nullptr;
  CU_CHECK(cudaMalloc(&d, (size_t)blocks * threads * 8));
  cudaEvent_t e0, e1;
  CU_CHECK(cudaEventCreate(&e0));
  CU_CHECK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, threads>>>(d, 0.999999, 1e-7, iters);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { cudaFree(d); CU_CHECK(e); }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *dfma_tflops = best;
  return FMCMC_OK;
}

__global__ void softplus_test_kernel(const double* a, double* out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = softplus_neg(a[i]);
}
extern "C" int fmcmc_test_softplus(int device, int64_t n, const double* a, double* out, char* err, size_t errlen) {
  if (n < 1 || !a || !out) { set_err(err, errlen, "bad argument"); return FMCMC_EINVAL; }
  CU_CHECK(cudaSetDevice(device));
  double* d = nullptr;
  CU_CHECK(cudaMalloc(&d, (size_t)n * 16));
  cudaMemcpy(d, a, (size_t)n * 8, cudaMemcpyHostToDevice);
  softplus_test_kernel<<<296, 256>>>(d, d + n, n);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(out, d + n, (size_t)n * 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  CU_CHECK(e);
  return FMCMC_OK;
}
