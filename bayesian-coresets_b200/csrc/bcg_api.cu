// C-ABI implementation (include/bcg.h) of the B200-native coreset engine.
// Host side: handle management, uploads, kernel launches on a private stream.  There is no CPU
// compute path: without a CUDA device every entry point fails with BCG_ERR_NO_DEVICE.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <chrono>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <utility>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <unistd.h>
#include <vector>

#include "../../include/bcg.h"
#include "bcg_state.h"
#include "filter_bounds.h"
#include "kernel_args.h"
#include "project_kernels.cuh"
#include "project_fast_kernel.cuh"
#include "project_sum_mma_kernel.cuh"
#include "project_mma_kernel.cuh"
#include "step_kernels.cuh"
#include "audit_kernel.cuh"
#include "lazy_select_kernel.cuh"
#include "pseudo_grad_kernel.cuh"
#include "sampler_kernels.cuh"

using namespace bcg;

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// error setter for the other translation units of the library (bcg_comm.cu)
int bcg_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(BCG_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
  } while (0)

#define RET(call)                  \
  do {                             \
    int r__ = (call);              \
    if (r__ != BCG_OK) return r__; \
  } while (0)

// temporary device / pinned-host buffers that are released on every exit path
template <typename T>
struct DevBuf {
  T* p = nullptr;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
  operator T*() const { return p; }
};
struct EventPair {
  cudaEvent_t e[2] = {nullptr, nullptr};
  ~EventPair() { for (auto x : e) if (x) cudaEventDestroy(x); }
};

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

static int pow2ceil(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// ------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------
struct bcg_ctx {
  int device;
  int sm_count;
  cudaDeviceProp prop;
  cudaStream_t stream;
  float* flush_buf;
  int64_t flush_bytes;
  unsigned char* pin[2];      // pinned staging for host -> device uploads
  cudaEvent_t pin_done[2];
  cudaStream_t copy_stream;   // second stream: uploads that overlap kernels on `stream`
  void* scratch[8];           // per-context device scratch, grown on demand (no malloc/free per projection pass)
  size_t scratch_bytes[8];
  double* sp_tab;             // softplus table of the fast links (softplus_table.h); null with BCG_FAST_LINK=0
  // N-sharding mailbox of this rank and the peers' mailboxes mapped so far.  Owned by the context, not by a solver:
  // opening a peer's IPC handle (and the lazy peer-access enable behind it) costs ~100 ms, so the mappings are made
  // once per process and reused by every solver; sized for kMaxWorld ranks and S = 1024 up front (67 KB), so the
  // allocation -- and with it the handle the peers hold -- never changes
  unsigned char* mail;
  int64_t mail_bytes;
  unsigned long long mail_epoch;   // solvers connected so far: every solver numbers its exchanges from epoch << 32,
                                   // so a value left in a slot by an earlier solver never matches (ranks connect in
                                   // the same order, SPMD, hence agree on the epoch)
  std::vector<std::pair<std::array<unsigned char, 64>, void*>> peer_cache;
  // one-slot cache of the last released matrix buffers: a cudaMalloc of gigabytes costs milliseconds -- far more once peer
  // access is enabled and the allocation has to be mapped for 7 peers -- and SparseVI / repeated coreset constructions
  // allocate a matrix of the same size again and again.  bcg_ctx_trim() releases it.
  float* pool_An;
  size_t pool_An_bytes;
  double* pool_norms;
  size_t pool_norms_bytes;
  uint16_t* pool_An16;
  size_t pool_An16_bytes;
};

struct bcg_vecs {
  bcg_ctx* ctx;
  int64_t n;
  int32_t S, ld;
  float* An;
  double* norms;
  uint16_t* An16;               // float16 copy of An for the pre-filter of the persistent scan (made on demand; may stay null)
  int32_t ld16;                 // halves per row of An16 (multiple of 8)
  size_t An16_bytes;
  size_t An_bytes, norms_bytes; // requested sizes (the buffers may be larger: they come from the context's one-slot cache)
  std::vector<double> colsum;   // S sums + [S] = sum of norms
  uint64_t zero_rows;
  // never-materialising form (bcg_dataset_project_lazy): An == null, rows are re-evaluated from the dataset on demand
  bcg_dataset* lazy_ds;         // borrowed: must outlive this object
  int32_t lazy_model, lazy_d;
  double *lazy_thetaT, *lazy_tt, *lazy_Siginv;   // owned device copies of the samples
};

struct bcg_solver {
  bcg_ctx* ctx;
  bcg_vecs* v;
  SolverState h;      // host mirror (valid after every synchronising call)
  SolverState* d;
  ScanConfig sc;
  bool use_loop;            // persistent cooperative kernel available for this shape (GIGA / FW: greedy_loop_kernel)
  bool use_omp_loop;        // persistent OrthoPursuit kernel available (omp_loop_kernel)
  bool filter16;            // the persistent kernel streams the float16 copy and re-scans by bounds (filter_bounds.h)
  LoopCtl* d_ctl;
  ScanCand* d_cta_cands;
  float* d_cta_lost;
  unsigned int* d_claims;
  int claims_cap;
  NnlsWork nw;              // device buffers of the NNLS factorisation (host copy of the struct)
  NnlsWork* d_nw;
  int nw_cap;
  int trace_on;
  unsigned long long* d_trace;
  int trace_cap, trace_n;
  int events_cap;           // capacity of h.events (kept across build calls)
  int64_t* d_fout;
  void* peer_ptrs[kMaxWorld];
  bool peers_open;
  // timing
  int profiling;
  cudaEvent_t ev0, ev1;
  std::vector<cudaEvent_t> scan_ev;
  float build_ms, scan_ms;
  int scan_launches, step_launches, loop_launches;
};

// context-owned scratch slot `i` of at least `bytes` (contents undefined); calls on one context are serialised
// and every user synchronises the stream before returning, so slots are free again at the next call
static int ctx_scratch(bcg_ctx* ctx, int i, size_t bytes, void** out) {
  if (ctx->scratch_bytes[i] < bytes) {
    if (ctx->scratch[i]) CK(cudaFree(ctx->scratch[i]));
    ctx->scratch[i] = nullptr;
    ctx->scratch_bytes[i] = 0;
    const size_t want = std::max(bytes, (size_t)1 << 20);
    CK(cudaMalloc(&ctx->scratch[i], want));
    ctx->scratch_bytes[i] = want;
  }
  *out = ctx->scratch[i];
  return BCG_OK;
}

static int use_device(bcg_ctx* ctx) {
  if (!ctx) return fail(BCG_ERR_ARG, "null context");
  CK(cudaSetDevice(ctx->device));
  return BCG_OK;
}

// ------------------------------------------------------------------------------------------
// library / context
// ------------------------------------------------------------------------------------------
extern "C" int bcg_abi_version(void) { return BCG_ABI_VERSION; }
extern "C" const char* bcg_last_error(void) { return g_err; }

extern "C" int bcg_device_count(int* count) {
  if (!count) return fail(BCG_ERR_ARG, "null count");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    *count = 0;
    return fail(BCG_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  *count = n;
  return BCG_OK;
}

extern "C" int bcg_ctx_create(int device, bcg_ctx** out) {
  if (!out) return fail(BCG_ERR_ARG, "null out");
  *out = nullptr;
  int n = 0;
  RET(bcg_device_count(&n));
  if (device < 0 || device >= n) return fail(BCG_ERR_ARG, "device %d out of range [0,%d)", device, n);
  bcg_ctx* c = new bcg_ctx();
  c->device = device;
  c->flush_buf = nullptr;
  c->flush_bytes = 0;
  c->pin[0] = c->pin[1] = nullptr;
  for (int i = 0; i < 8; ++i) { c->scratch[i] = nullptr; c->scratch_bytes[i] = 0; }
  c->stream = nullptr;
  c->copy_stream = nullptr;
  c->mail = nullptr;
  c->mail_bytes = 0;
  c->mail_epoch = 0;
  c->sp_tab = nullptr;
  c->pool_An = nullptr; c->pool_An_bytes = 0;
  c->pool_norms = nullptr; c->pool_norms_bytes = 0;
  c->pool_An16 = nullptr; c->pool_An16_bytes = 0;
  auto body = [&]() -> int {
    CK(cudaSetDevice(device));
    CK(cudaGetDeviceProperties(&c->prop, device));
    if (c->prop.major < 10)
      return fail(BCG_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                  c->prop.major, c->prop.minor);
    c->sm_count = c->prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    if (env_int("BCG_FAST_LINK", 1)) {
      std::vector<double> tab(kSpTableDoubles);
      softplus_table_build(tab.data());
      CK(cudaMalloc(&c->sp_tab, kSpTableDoubles * sizeof(double)));
      CK(cudaMemcpy(c->sp_tab, tab.data(), kSpTableDoubles * sizeof(double), cudaMemcpyHostToDevice));
    }
    return BCG_OK;
  };
  const int rc = body();
  if (rc != BCG_OK) { bcg_ctx_destroy(c); return rc; }    // (bcg_ctx_destroy leaves the error message alone)
  *out = c;
  return BCG_OK;
}

extern "C" int bcg_ctx_destroy(bcg_ctx* ctx) {
  if (!ctx) return BCG_OK;
  cudaSetDevice(ctx->device);
  if (ctx->flush_buf) cudaFree(ctx->flush_buf);
  if (ctx->sp_tab) cudaFree(ctx->sp_tab);
  if (ctx->pool_An) cudaFree(ctx->pool_An);
  if (ctx->pool_norms) cudaFree(ctx->pool_norms);
  if (ctx->pool_An16) cudaFree(ctx->pool_An16);
  for (auto& pc : ctx->peer_cache) cudaIpcCloseMemHandle(pc.second);
  if (ctx->mail) cudaFree(ctx->mail);
  for (int i = 0; i < 8; ++i)
    if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
  for (int i = 0; i < 2; ++i)
    if (ctx->pin[i]) { cudaFreeHost(ctx->pin[i]); cudaEventDestroy(ctx->pin_done[i]); }
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
  return BCG_OK;
}

extern "C" int bcg_ctx_info(bcg_ctx* ctx, char* name, int name_cap, int* sm_count, int* cc_major, int* cc_minor,
                            int64_t* total_mem_bytes) {
  if (!ctx) return fail(BCG_ERR_ARG, "null context");
  if (name && name_cap > 0) {
    strncpy(name, ctx->prop.name, name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (sm_count) *sm_count = ctx->sm_count;
  if (cc_major) *cc_major = ctx->prop.major;
  if (cc_minor) *cc_minor = ctx->prop.minor;
  if (total_mem_bytes) *total_mem_bytes = (int64_t)ctx->prop.totalGlobalMem;
  return BCG_OK;
}

extern "C" int bcg_ctx_mem_info(bcg_ctx* ctx, int64_t* free_bytes, int64_t* total_bytes) {
  RET(use_device(ctx));
  size_t f = 0, t = 0;
  CK(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = (int64_t)f;
  if (total_bytes) *total_bytes = (int64_t)t;
  return BCG_OK;
}

extern "C" int bcg_ctx_trim(bcg_ctx* ctx) {
  RET(use_device(ctx));
  if (ctx->pool_An) CK(cudaFree(ctx->pool_An));
  if (ctx->pool_norms) CK(cudaFree(ctx->pool_norms));
  if (ctx->pool_An16) CK(cudaFree(ctx->pool_An16));
  ctx->pool_An = nullptr; ctx->pool_An_bytes = 0;
  ctx->pool_norms = nullptr; ctx->pool_norms_bytes = 0;
  ctx->pool_An16 = nullptr; ctx->pool_An16_bytes = 0;
  return BCG_OK;
}

// take a buffer of at least `bytes` from the one-slot cache when it fits without wasting more than half of it
template <typename T>
static int pool_take(T** slot, size_t* slot_bytes, size_t bytes, T** out) {
  if (*slot && *slot_bytes >= bytes && *slot_bytes <= 2 * bytes + (1 << 20)) {
    *out = *slot;
    *slot = nullptr;
    *slot_bytes = 0;
    return BCG_OK;
  }
  CK(cudaMalloc(out, bytes));
  return BCG_OK;
}
// the most recently released buffer stays (repeated constructions ask for the size they just released)
template <typename T>
static void pool_give(T** slot, size_t* slot_bytes, T* p, size_t bytes) {
  if (!p) return;
  if (*slot) cudaFree(*slot);
  *slot = p;
  *slot_bytes = bytes;
}

extern "C" int bcg_ctx_synchronize(bcg_ctx* ctx) {
  RET(use_device(ctx));
  CK(cudaStreamSynchronize(ctx->stream));
  return BCG_OK;
}

extern "C" int bcg_ctx_flush_l2(bcg_ctx* ctx, int64_t bytes) {
  RET(use_device(ctx));
  if (bytes <= 0) return BCG_OK;
  if (ctx->flush_bytes < bytes) {
    if (ctx->flush_buf) CK(cudaFree(ctx->flush_buf));
    ctx->flush_buf = nullptr;
    CK(cudaMalloc(&ctx->flush_buf, bytes));
    ctx->flush_bytes = bytes;
  }
  fill_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->flush_buf, bytes / 4, 0.f);
  CK(cudaGetLastError());
  return BCG_OK;
}


// ------------------------------------------------------------------------------------------
// host -> device upload of pageable memory: multi-threaded memcpy into two pinned staging
// buffers, overlapped with the asynchronous copies (a plain cudaMemcpy from pageable memory
// runs at a fraction of the PCIe rate)
// ------------------------------------------------------------------------------------------
static const size_t kPinChunk = (size_t)32 << 20;

// Host worker threads for the staging copies of pageable sources, started on first use and kept (one staging copy per
// 16 - 32 MB chunk: spawning threads per chunk cost as much as the copy itself).  Process-wide; re-created after a fork.
struct HostWorkers {
  std::vector<std::thread> th;
  std::mutex m;
  std::condition_variable cv_go, cv_done;
  const char* src = nullptr;
  char* dst = nullptr;
  size_t bytes = 0, per = 0;
  uint64_t gen = 0;
  int pending = 0;
  explicit HostWorkers(int n) {
    for (int t = 0; t < n; ++t) th.emplace_back([this, t]() { loop(t); });
  }
  void loop(int t) {
    uint64_t seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(m);
      cv_go.wait(lk, [&]() { return gen != seen; });
      seen = gen;
      const size_t off = (size_t)t * per;
      const size_t len = off < bytes ? std::min(per, bytes - off) : 0;
      const char* s = src;
      char* d = dst;
      lk.unlock();
      if (len) memcpy(d + off, s + off, len);
      lk.lock();
      if (--pending == 0) cv_done.notify_one();
    }
  }
  void copy(void* d, const void* s, size_t n) {
    std::unique_lock<std::mutex> lk(m);
    dst = (char*)d; src = (const char*)s; bytes = n;
    per = (n / th.size() + 4095) / 4096 * 4096;
    pending = (int)th.size();
    ++gen;
    cv_go.notify_all();
    cv_done.wait(lk, [&]() { return pending == 0; });
  }
};
static HostWorkers* g_workers = nullptr;
static pid_t g_workers_pid = 0;
static std::mutex g_workers_mu;

static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
  unsigned hw = std::thread::hardware_concurrency();
  const int nt = (int)std::max(1u, std::min(8u, hw ? hw / 2 : 1u));
  if (nt == 1 || bytes < ((size_t)4 << 20)) { memcpy(dst, src, bytes); return; }
  std::lock_guard<std::mutex> g(g_workers_mu);
  if (!g_workers || g_workers_pid != getpid()) {          // (a forked child inherits the pointer but not the threads)
    g_workers = new HostWorkers(nt);
    g_workers_pid = getpid();
  }
  g_workers->copy(dst, src, bytes);
}

extern "C" int bcg_host_alloc(int64_t bytes, void** out) {
  if (!out || bytes < 0) return fail(BCG_ERR_ARG, "bad arguments");
  *out = nullptr;
  int n = 0;
  RET(bcg_device_count(&n));
  CK(cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocPortable));
  return BCG_OK;
}

extern "C" int bcg_host_free(void* p) {
  if (p) CK(cudaFreeHost(p));
  return BCG_OK;
}

// page-locked source (bcg_host_alloc / cudaHostRegister): DMA reads it directly, no staging copy
static bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

// the two context-owned pinned staging buffers (kPinChunk bytes each), allocated on first use and kept:
// cudaMallocHost / cudaFreeHost cost milliseconds per call
static int ensure_pins(bcg_ctx* ctx) {
  for (int i = 0; i < 2; ++i)
    if (!ctx->pin[i]) {
      CK(cudaMallocHost(&ctx->pin[i], kPinChunk));
      CK(cudaEventCreateWithFlags(&ctx->pin_done[i], cudaEventDisableTiming));
    }
  return BCG_OK;
}

static int h2d(bcg_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return BCG_OK;
  if (bytes < ((size_t)1 << 20)) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // the source may be a temporary of the caller
    return BCG_OK;
  }
  if (is_pinned(src)) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BCG_OK;
  }
  RET(ensure_pins(ctx));
  int c = 0;
  for (size_t off = 0; off < bytes; off += kPinChunk, ++c) {
    const int i = c & 1;
    const size_t len = std::min(kPinChunk, bytes - off);
    if (c >= 2) CK(cudaEventSynchronize(ctx->pin_done[i]));
    parallel_memcpy(ctx->pin[i], (const char*)src + off, len);
    CK(cudaMemcpyAsync((char*)dst + off, ctx->pin[i], len, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->pin_done[i], ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return BCG_OK;
}

// ------------------------------------------------------------------------------------------
// projection matrix
// ------------------------------------------------------------------------------------------
static int j_for_ld(int ld) {
  int j = pow2ceil((ld + 31) / 32);
  return j;
}

// is the float16 pre-filter available for rows of ld floats (same rule as choose_scan_config)
static bool filter16_shape_ok(int ld) {
  if (!env_int("BCG_FILTER16", 1)) return false;
  const int nchunk = ld / 4;
  const int lpr = std::min(32, pow2ceil(nchunk));
  const int ch = pow2ceil((nchunk + lpr - 1) / lpr);
  return loop_variant_exists(ch, lpr) && loop_variant_ch16(ch, lpr) > 0;
}

// buffer for a float16 copy from the context's one-slot cache or the device; false (no error) when there is no memory
static bool half_take(bcg_ctx* ctx, size_t bytes, uint16_t** out) {
  if (ctx->pool_An16 && ctx->pool_An16_bytes >= bytes && ctx->pool_An16_bytes <= 2 * bytes + (1 << 20)) {
    *out = ctx->pool_An16;
    ctx->pool_An16 = nullptr;
    ctx->pool_An16_bytes = 0;
    return true;
  }
  if (cudaMalloc(out, bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    *out = nullptr;
    return false;
  }
  return true;
}

// rows of padding behind An: the float32 re-scan of the filtered persistent kernel reads whole row batches from global memory
static const int64_t kRowPad = 64;

static int vecs_alloc(bcg_ctx* ctx, int64_t n, int32_t S, bcg_vecs** out, bool lazy = false) {
  if (n < 0 || S <= 0) return fail(BCG_ERR_ARG, "bad shape n=%lld S=%d", (long long)n, S);
  if (S > 1024) return fail(BCG_ERR_UNSUPPORTED, "S=%d > 1024 is not supported", S);
  if (n >= (1ll << 32) - 1) return fail(BCG_ERR_UNSUPPORTED, "more than 2^32-2 local rows");
  bcg_vecs* v = new bcg_vecs();
  v->ctx = ctx;
  v->n = n;
  v->S = S;
  v->ld = (S + 3) / 4 * 4;
  v->An = nullptr;
  v->norms = nullptr;
  v->An16 = nullptr;
  v->ld16 = (S + 7) / 8 * 8;
  v->An16_bytes = 0;
  v->zero_rows = 0;
  v->colsum.assign(S + 1, 0.);
  v->lazy_ds = nullptr;
  v->lazy_model = v->lazy_d = 0;
  v->lazy_thetaT = v->lazy_tt = v->lazy_Siginv = nullptr;
  v->An_bytes = v->norms_bytes = 0;
  if (n > 0) {
    if (!lazy) {
      v->An_bytes = (size_t)(n + kRowPad) * v->ld * sizeof(float);   // (padding: see kRowPad)
      RET(pool_take(&ctx->pool_An, &ctx->pool_An_bytes, v->An_bytes, &v->An));
    }
    v->norms_bytes = (size_t)n * sizeof(double);
    RET(pool_take(&ctx->pool_norms, &ctx->pool_norms_bytes, v->norms_bytes, &v->norms));
  }
  *out = v;
  return BCG_OK;
}

// reduce `nparts` partial column-sum rows on the device and fetch them
static int finish_colsum(bcg_vecs* v, const double* d_partial, int nparts, unsigned long long* d_zero) {
  bcg_ctx* ctx = v->ctx;
  const int S1 = v->S + 1;
  DevBuf<double> d_out;
  CK(d_out.alloc(S1));
  colsum_reduce_kernel<<<(S1 + 31) / 32, dim3(32, kCsrSlices), 0, ctx->stream>>>(d_partial, nparts, S1, d_out);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(v->colsum.data(), d_out, S1 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  unsigned long long z = 0;
  CK(cudaMemcpyAsync(&z, d_zero, sizeof(z), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  v->zero_rows = z;
  return BCG_OK;
}

template <int J>
static int launch_ingest(bcg_ctx* ctx, const double* src, int64_t src_ld, int64_t n, bcg_vecs* v, int64_t row0,
                         double* partial, int grid, unsigned long long* d_zero) {
  const size_t smem = (size_t)kProjWarps * (v->S + 1) * sizeof(double);
  CK(cudaFuncSetAttribute(ingest_kernel<J>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ingest_kernel<J><<<grid, kProjWarps * 32, smem, ctx->stream>>>(src, src_ld, n, v->S, v->ld,
                                                                  v->An + (size_t)row0 * v->ld, v->norms + row0,
                                                                  partial, d_zero);
  CK(cudaGetLastError());
  return BCG_OK;
}

static int dispatch_ingest(bcg_ctx* ctx, const double* src, int64_t src_ld, int64_t n, bcg_vecs* v, int64_t row0,
                           double* partial, int grid, unsigned long long* d_zero) {
  switch (j_for_ld(v->ld)) {
    case 1: return launch_ingest<1>(ctx, src, src_ld, n, v, row0, partial, grid, d_zero);
    case 2: return launch_ingest<2>(ctx, src, src_ld, n, v, row0, partial, grid, d_zero);
    case 4: return launch_ingest<4>(ctx, src, src_ld, n, v, row0, partial, grid, d_zero);
    case 8: return launch_ingest<8>(ctx, src, src_ld, n, v, row0, partial, grid, d_zero);
    case 16: return launch_ingest<16>(ctx, src, src_ld, n, v, row0, partial, grid, d_zero);
    default: return launch_ingest<32>(ctx, src, src_ld, n, v, row0, partial, grid, d_zero);
  }
}

// double-buffered staging through the context's pinned buffers: the host memcpy of chunk c+1 overlaps H2D + ingest of
// chunk c; the host waits only for its staging buffer (copy of chunk c-2 done), the device buffer is ordered by events
static int ingest_host_rows(bcg_ctx* ctx, const double* rows, int64_t n, int32_t S, int64_t ld_host, bcg_vecs* v) {
  const int64_t chunk_rows = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)kPinChunk / ((int64_t)S * 8)));
  const int nchunks = (int)((n + chunk_rows - 1) / chunk_rows);
  const int grid = (int)std::min<int64_t>((chunk_rows + kProjWarps - 1) / kProjWarps, (int64_t)ctx->sm_count * 2);
  cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
  RET(ensure_pins(ctx));
  double* pin[2] = {reinterpret_cast<double*>(ctx->pin[0]), reinterpret_cast<double*>(ctx->pin[1])};
  double* dev[2] = {nullptr, nullptr};
  EventPair copied, kdone;
  DevBuf<double> d_partial;
  DevBuf<unsigned long long> d_zero;
  const size_t celems = (size_t)chunk_rows * S;
  for (int i = 0; i < 2; ++i) {
    RET(ctx_scratch(ctx, 6 + i, celems * sizeof(double), (void**)&dev[i]));
    CK(cudaEventCreateWithFlags(&copied.e[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&kdone.e[i], cudaEventDisableTiming));
  }
  CK(d_partial.alloc((size_t)nchunks * grid * (S + 1)));
  CK(d_zero.alloc(1));
  CK(cudaMemsetAsync(d_zero, 0, sizeof(unsigned long long), st));
  for (int c = 0; c < nchunks; ++c) {
    const int i = c & 1;
    const int64_t r0 = (int64_t)c * chunk_rows;
    const int64_t nr = std::min<int64_t>(chunk_rows, n - r0);
    if (c >= 2) {
      CK(cudaStreamWaitEvent(cs, kdone.e[i], 0));
      CK(cudaEventSynchronize(copied.e[i]));
    }
    if (ld_host == S) {
      parallel_memcpy(pin[i], rows + r0 * ld_host, (size_t)nr * S * sizeof(double));
    } else {
      for (int64_t r = 0; r < nr; ++r) memcpy(pin[i] + r * S, rows + (r0 + r) * ld_host, (size_t)S * sizeof(double));
    }
    CK(cudaMemcpyAsync(dev[i], pin[i], (size_t)nr * S * sizeof(double), cudaMemcpyHostToDevice, cs));
    CK(cudaEventRecord(copied.e[i], cs));
    CK(cudaStreamWaitEvent(st, copied.e[i], 0));
    RET(dispatch_ingest(ctx, dev[i], S, nr, v, r0, d_partial.p + (size_t)c * grid * (S + 1), grid, d_zero));
    CK(cudaEventRecord(kdone.e[i], st));
  }
  RET(finish_colsum(v, d_partial, nchunks * grid, d_zero));
  CK(cudaStreamSynchronize(cs));
  return BCG_OK;
}

extern "C" int bcg_vecs_from_host_f64(bcg_ctx* ctx, const double* rows, int64_t n, int32_t S, int64_t ld_host,
                                      bcg_vecs** out) {
  RET(use_device(ctx));
  if (!out || (n > 0 && !rows) || ld_host < S) return fail(BCG_ERR_ARG, "bad arguments");
  *out = nullptr;
  bcg_vecs* v = nullptr;
  RET(vecs_alloc(ctx, n, S, &v));
  if (n == 0) { *out = v; return BCG_OK; }
  const int rc = ingest_host_rows(ctx, rows, n, S, ld_host, v);
  if (rc != BCG_OK) { bcg_vecs_destroy(v); return rc; }
  *out = v;
  return BCG_OK;
}

struct bcg_dataset {
  bcg_ctx* ctx;
  int64_t n;
  int32_t zld;
  double* Z;      // device, n x zld float64
};

extern "C" int bcg_dataset_create(bcg_ctx* ctx, const double* Z, int64_t n, int32_t zld, bcg_dataset** out) {
  RET(use_device(ctx));
  if (!out || n < 0 || zld <= 0 || (n > 0 && !Z)) return fail(BCG_ERR_ARG, "bad arguments");
  bcg_dataset* ds = new bcg_dataset();
  ds->ctx = ctx; ds->n = n; ds->zld = zld; ds->Z = nullptr;
  if (n > 0) {
    CK(cudaMalloc(&ds->Z, (size_t)n * zld * sizeof(double)));
    RET(h2d(ctx, ds->Z, Z, (size_t)n * zld * sizeof(double)));
  }
  *out = ds;
  return BCG_OK;
}

extern "C" int bcg_dataset_destroy(bcg_dataset* ds) {
  if (!ds) return BCG_OK;
  cudaSetDevice(ds->ctx->device);
  if (ds->Z) cudaFree(ds->Z);
  delete ds;
  return BCG_OK;
}

template <int J>
static int launch_project(bcg_ctx* ctx, const ProjectArgs& a, int grid, size_t smem) {
  CK(cudaFuncSetAttribute(project_kernel<J>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  project_kernel<J><<<grid, kProjWarps * 32, smem, ctx->stream>>>(a);
  CK(cudaGetLastError());
  return BCG_OK;
}

template <int J2, int MODEL>
static int launch_project_fast(bcg_ctx* ctx, const ProjectArgs& a, int grid, size_t smem) {
  CK(cudaFuncSetAttribute(project_fast_kernel<J2, MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  project_fast_kernel<J2, MODEL><<<grid, kProjWarps * 32, smem, ctx->stream>>>(a);
  CK(cudaGetLastError());
  return BCG_OK;
}

template <int J2>
static int launch_project_fast_model(bcg_ctx* ctx, const ProjectArgs& a, int grid, size_t smem) {
  if (a.model == MODEL_LR) return launch_project_fast<J2, MODEL_LR>(ctx, a, grid, smem);
  if (a.model == MODEL_POISSON) return launch_project_fast<J2, MODEL_POISSON>(ctx, a, grid, smem);
  return launch_project_fast<J2, MODEL_LINEAR>(ctx, a, grid, smem);
}

template <int MODEL, int MI>
static int launch_project_mma(bcg_ctx* ctx, const ProjectArgs& a, int grid) {
  const size_t smem = project_mma_smem(a.S, MI);
  CK(cudaFuncSetAttribute(project_mma_kernel<MODEL, MI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  project_mma_kernel<MODEL, MI><<<grid, kPjThreads, smem, ctx->stream>>>(a);   // (CTAs without rows write zero partial sums)
  CK(cudaGetLastError());
  return BCG_OK;
}

template <int MI>
static int launch_project_mma_model(bcg_ctx* ctx, const ProjectArgs& a, int grid) {
  if (a.model == MODEL_LR) return launch_project_mma<MODEL_LR, MI>(ctx, a, grid);
  if (a.model == MODEL_POISSON) return launch_project_mma<MODEL_POISSON, MI>(ctx, a, grid);
  return launch_project_mma<MODEL_LINEAR, MI>(ctx, a, grid);
}

// the DMMA materialising projection (project_mma_kernel.cuh) applies when the unit-row matrix is written, S is 64, 128,
// 256 or 512 and the links come from the table; the partial column sums need room for its grid (see project_mma_grid_cap)
static bool project_mma_ok(const ProjectArgs& a) {
  // BCG_PROJ_MMA: 0 never, 2 whenever the shape allows, 1 (default) where it measured faster than the warp-per-row kernel
  // (profiles/r02_k3_timing.txt: d > 32, where only the general kernel applies otherwise -- 20.9 vs 43.1 ms at d = 200,
  // S = 512, N = 1e6 -- and S <= 256 -- 1.18 vs 1.25 ms at d = 10, S = 256; at S = 512, d = 10 it is 22.9 vs 21.9 ms)
  const int mode = env_int("BCG_PROJ_MMA", 1);
  const bool shape_ok = a.An && !a.out64 && a.ld == a.S && (a.S == 64 || a.S == 128 || a.S == 256 || a.S == 512) &&
                        (a.model == MODEL_LINEAR || a.sp_tab != nullptr);
  if (!mode || !shape_ok) return false;
  return mode >= 2 || a.d > 32 || a.S <= 256;
}

// CTAs (= rows of the partial column-sum buffer) of a projection launch over at most a.n rows
static int project_grid(bcg_ctx* ctx, const ProjectArgs& a) {
  if (project_mma_ok(a)) {
    const int WR = 8 / (a.S / 64);
    const bool resident = a.d <= kPjKT;
    const int bm = (resident ? 8 : 32) * WR;
    return (int)std::min<int64_t>((a.n + bm - 1) / bm, (int64_t)ctx->sm_count * (resident ? 2 : 1));
  }
  return (int)std::min<int64_t>((a.n + kProjWarps - 1) / kProjWarps, (int64_t)ctx->sm_count);
}

// the warp-per-row specialised kernel applies (BCG_PROJ_FAST=0 forces the general kernel)
static bool project_fast_ok(const ProjectArgs& a) {
  if (project_mma_ok(a)) return false;
  const int fast = env_int("BCG_PROJ_FAST", 1);
  return fast && !a.out64 && a.ld == a.S && a.d <= 32 && a.ktile >= a.d && (a.model == MODEL_LINEAR || a.sp_tab != nullptr) &&
         (a.S == 64 || a.S == 128 || a.S == 256 || a.S == 512);
}

// K3 launch: the specialised kernels (project_fast_kernel.cuh) for S in {64, 128, 256, 512} with the whole sample tile
// in shared memory and table links, the general kernel otherwise (BCG_PROJ_FAST=0 forces the general kernel)
static int dispatch_project(bcg_ctx* ctx, const ProjectArgs& a, int grid, size_t smem) {
  if (project_mma_ok(a)) {
    // MI = 1: the whole sample tile resident (d <= 16); MI = 4: streamed k tiles, 32-row blocks (any d)
    if (a.d <= kPjKT) return launch_project_mma_model<1>(ctx, a, grid);
    return launch_project_mma_model<4>(ctx, a, grid);
  }
  if (project_fast_ok(a)) {
    switch (a.S) {
      case 64: return launch_project_fast_model<1>(ctx, a, grid, smem);
      case 128: return launch_project_fast_model<2>(ctx, a, grid, smem);
      case 256: return launch_project_fast_model<4>(ctx, a, grid, smem);
      case 512: return launch_project_fast_model<8>(ctx, a, grid, smem);
      default: break;
    }
  }
  switch (j_for_ld(a.ld)) {
    case 1: return launch_project<1>(ctx, a, grid, smem);
    case 2: return launch_project<2>(ctx, a, grid, smem);
    case 4: return launch_project<4>(ctx, a, grid, smem);
    case 8: return launch_project<8>(ctx, a, grid, smem);
    case 16: return launch_project<16>(ctx, a, grid, smem);
    default: return launch_project<32>(ctx, a, grid, smem);
  }
}

// common driver.  thetaT (d x S, already transposed) and coff (S or null) are host arrays.
// out_vecs: materialise the unit-row matrix; rows64: host n x S float64 centred rows; colsum: host S.
static int project_common(bcg_dataset* ds, const int64_t* rowidx, int64_t nsel, int32_t d, const double* thetaT,
                          const double* coff, int32_t S, int model, bcg_vecs** out_vecs, double* rows64, double* colsum) {
  bcg_ctx* ctx = ds->ctx;
  const int64_t n = rowidx ? nsel : ds->n;
  if (rowidx)
    for (int64_t i = 0; i < nsel; ++i)
      if (rowidx[i] < 0 || rowidx[i] >= ds->n) return fail(BCG_ERR_ARG, "row index %lld out of range", (long long)rowidx[i]);
  if (out_vecs) *out_vecs = nullptr;
  if (d <= 0 || S <= 0) return fail(BCG_ERR_ARG, "d and S must be positive");
  if (S > 1024) return fail(BCG_ERR_UNSUPPORTED, "S=%d > 1024 is not supported", S);
  if ((model == MODEL_POISSON ? d + 1 : d) > ds->zld) return fail(BCG_ERR_ARG, "dataset has too few columns");
  if (n == 0) {
    if (out_vecs) RET(vecs_alloc(ctx, 0, S, out_vecs));
    if (colsum) memset(colsum, 0, (size_t)S * sizeof(double));
    return BCG_OK;
  }
  const auto t_entry = std::chrono::steady_clock::now();
  cudaStream_t st = ctx->stream;
  double *dT = nullptr, *dC = nullptr, *d_partial = nullptr, *d_out = nullptr;
  unsigned long long* d_zero = nullptr;
  int64_t* d_idx = nullptr;
  DevBuf<double> d_rows;
  if (rowidx) {
    RET(ctx_scratch(ctx, 0, (size_t)n * sizeof(int64_t), (void**)&d_idx));
    CK(cudaMemcpyAsync(d_idx, rowidx, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  }
  RET(ctx_scratch(ctx, 1, (size_t)d * S * sizeof(double), (void**)&dT));
  CK(cudaMemcpyAsync(dT, thetaT, (size_t)d * S * sizeof(double), cudaMemcpyHostToDevice, st));
  if (coff) {
    RET(ctx_scratch(ctx, 2, (size_t)S * sizeof(double), (void**)&dC));
    CK(cudaMemcpyAsync(dC, coff, (size_t)S * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  RET(ctx_scratch(ctx, 3, (size_t)(S + 1) * sizeof(double), (void**)&d_out));
  RET(ctx_scratch(ctx, 4, sizeof(unsigned long long), (void**)&d_zero));

  if (!out_vecs && !rows64 && colsum && d >= env_int("BCG_PROJSUM_MIN_D", 24) && n >= 4096) {
    // K3b, GEMM-shaped: float64 tensor-core kernel (project_sum_mma_kernel.cuh)
    const int64_t nrb = (n + kPsBM - 1) / kPsBM;
    const int grid = (int)std::min<int64_t>(nrb, (int64_t)ctx->sm_count);
    RET(ctx_scratch(ctx, 5, (size_t)grid * S * sizeof(double), (void**)&d_partial));
    CK(cudaMemsetAsync(d_partial, 0, (size_t)grid * S * sizeof(double), st));
    ProjectSumArgs pa;
    pa.Z = ds->Z; pa.rowidx = d_idx; pa.thetaT = dT; pa.coff = dC; pa.partial = d_partial; pa.n = n; pa.zld = ds->zld; pa.d = d; pa.S = S;
    pa.model = model; pa.sp_tab = ctx->sp_tab;
    const bool trace = env_int("BCG_PROJ_TRACE", 0) != 0;
    const auto t_launch = std::chrono::steady_clock::now();
    auto launch_sum = [&](auto kern) -> int {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPm2SmemBytes));
      kern<<<grid, kPsThreads, kPm2SmemBytes, st>>>(pa);
      return BCG_OK;
    };
    if (model == MODEL_LR) RET(launch_sum(project_sum_mma_kernel<MODEL_LR>));
    else if (model == MODEL_POISSON) RET(launch_sum(project_sum_mma_kernel<MODEL_POISSON>));
    else RET(launch_sum(project_sum_mma_kernel<MODEL_LINEAR>));
    CK(cudaGetLastError());
    project_sum_finish_kernel<<<1, 256, 0, st>>>(d_partial, grid, S, d_out);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(colsum, d_out, (size_t)S * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (trace) {
      const auto t_done = std::chrono::steady_clock::now();
      fprintf(stderr, "[bcg] project_sum: prep %.3f ms, kernels+readback %.3f ms\n",
              std::chrono::duration<double, std::milli>(t_launch - t_entry).count(),
              std::chrono::duration<double, std::milli>(t_done - t_launch).count());
    }
    return BCG_OK;
  }

  // K3 (and K3b for small d): one warp per row, theta tile in shared memory (project_kernels.cuh)
  const int ld = (S + 3) / 4 * 4;
  const size_t cs_bytes = (size_t)kProjWarps * (S + 1) * sizeof(double);
  const size_t budget = 200 * 1024;
  if (cs_bytes + (size_t)S * sizeof(double) > budget)
    return fail(BCG_ERR_UNSUPPORTED, "projection tile does not fit shared memory for S=%d", S);
  const int ktile = (int)std::min<size_t>(std::min<size_t>(kProjKTile, (size_t)d), (budget - cs_bytes) / ((size_t)S * sizeof(double)));
  const size_t smem = (size_t)ktile * S * sizeof(double) + cs_bytes;
  bcg_vecs* v = nullptr;
  if (out_vecs) RET(vecs_alloc(ctx, n, S, &v));
  int grid = 0;
  auto body = [&]() -> int {
    CK(cudaMemsetAsync(d_zero, 0, sizeof(unsigned long long), st));
    if (rows64) CK(d_rows.alloc((size_t)n * S));
    ProjectArgs a;
    a.Z = ds->Z; a.rowidx = d_idx; a.theta = dT; a.coff = dC; a.An = v ? v->An : nullptr; a.norms = v ? v->norms : nullptr;
    a.out64 = d_rows; a.partial = nullptr; a.zero_rows = d_zero; a.n = n; a.zld = ds->zld; a.d = d; a.S = S;
    a.ld = ld; a.model = model; a.ktile = ktile; a.sp_tab = ctx->sp_tab;
    grid = project_grid(ctx, a);
    RET(ctx_scratch(ctx, 5, (size_t)grid * (S + 1) * sizeof(double), (void**)&d_partial));
    a.partial = d_partial;
    if (env_int("BCG_PROJ_TRACE", 0)) {               // diagnostics: device time of the projection kernel alone
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0, st));
      RET(dispatch_project(ctx, a, grid, smem));
      CK(cudaEventRecord(e1, st));
      CK(cudaEventSynchronize(e1));
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      fprintf(stderr, "[bcg] projection kernel: n=%lld d=%d S=%d model=%d mma=%d grid=%d: %.3f ms\n", (long long)n, d, S, model,
              (int)project_mma_ok(a), grid, ms);
      cudaEventDestroy(e0); cudaEventDestroy(e1);
    } else
    RET(dispatch_project(ctx, a, grid, smem));
    if (v) {
      RET(finish_colsum(v, d_partial, grid, d_zero));
      if (colsum) memcpy(colsum, v->colsum.data(), (size_t)S * sizeof(double));
    } else if (colsum) {
      const int S1 = S + 1;
      colsum_reduce_kernel<<<(S1 + 31) / 32, dim3(32, kCsrSlices), 0, st>>>(d_partial, grid, S1, d_out);
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(colsum, d_out, (size_t)S * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if (rows64) CK(cudaMemcpyAsync(rows64, d_rows, (size_t)n * S * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BCG_OK;
  };
  const int rc = body();
  if (rc != BCG_OK) { if (v) bcg_vecs_destroy(v); return rc; }
  if (out_vecs) *out_vecs = v;
  return BCG_OK;
}

static std::vector<double> transpose_sd(const double* theta, int S, int d) {
  std::vector<double> t((size_t)S * d);
  for (int s = 0; s < S; ++s)
    for (int k = 0; k < d; ++k) t[(size_t)k * S + s] = theta[(size_t)s * d + k];
  return t;
}

static int prepare_model(int model, int d, const double* theta, int S, const double* Siginv, std::vector<double>& tT,
                         std::vector<double>& coff, int* kmodel);

// model: BCG_MODEL_*; theta: host S x d; Siginv: host d x d (Gaussian only).  Any of out_vecs / rows64 /
// colsum may be null; with only colsum the N x S matrix is never written (K3b).
extern "C" int bcg_dataset_project(bcg_dataset* ds, const int64_t* rowidx, int64_t nsel, int32_t model, int32_t d,
                                   const double* theta, int32_t S, const double* Siginv, bcg_vecs** out_vecs,
                                   double* rows64, double* colsum) {
  if (!ds || !theta) return fail(BCG_ERR_ARG, "null argument");
  RET(use_device(ds->ctx));
  std::vector<double> tT, coff;
  int kmodel = 0;
  RET(prepare_model(model, d, theta, S, Siginv, tT, coff, &kmodel));
  return project_common(ds, rowidx, nsel, d, tT.data(), coff.empty() ? nullptr : coff.data(), S, kmodel, out_vecs, rows64,
                        colsum);
}

// Gaussian model with the S x d matrix A = theta Siginv and the offsets c_s = -0.5 theta_s Siginv theta_s
// precomputed by the caller (BLAS on the host instead of the O(S d^2) loop above)
extern "C" int bcg_dataset_project_linear(bcg_dataset* ds, const int64_t* rowidx, int64_t nsel, int32_t d, const double* A,
                                          const double* coff, int32_t S, bcg_vecs** out_vecs, double* rows64,
                                          double* colsum) {
  if (!ds || !A) return fail(BCG_ERR_ARG, "null argument");
  RET(use_device(ds->ctx));
  std::vector<double> tT = transpose_sd(A, S, d);
  return project_common(ds, rowidx, nsel, d, tT.data(), coff, S, MODEL_LINEAR, out_vecs, rows64, colsum);
}

// thetaT (d x S) / coff (S) / kernel model for a C-ABI model id; Gaussian: A = theta Siginv, c = -0.5 theta.A
static int prepare_model(int model, int d, const double* theta, int S, const double* Siginv, std::vector<double>& tT,
                         std::vector<double>& coff, int* kmodel) {
  coff.clear();
  if (model == BCG_MODEL_LR || model == BCG_MODEL_POISSON) {
    tT = transpose_sd(theta, S, d);
    *kmodel = model == BCG_MODEL_LR ? MODEL_LR : MODEL_POISSON;
    return BCG_OK;
  }
  if (model == BCG_MODEL_GAUSSIAN) {
    if (!Siginv) return fail(BCG_ERR_ARG, "Siginv is required for the Gaussian model");
    tT.assign((size_t)S * d, 0.);
    coff.assign(S, 0.);
    for (int s = 0; s < S; ++s) {
      double q = 0.;
      for (int i = 0; i < d; ++i) {
        double m = 0.;
        for (int j = 0; j < d; ++j) m += Siginv[(size_t)i * d + j] * theta[(size_t)s * d + j];
        tT[(size_t)i * S + s] = m;
        q += theta[(size_t)s * d + i] * m;
      }
      coff[s] = -0.5 * q;
    }
    *kmodel = MODEL_LINEAR;
    return BCG_OK;
  }
  return fail(BCG_ERR_ARG, "unknown model %d", model);
}

// Independent float64 audit of one selection pass, recomputed from the raw data (audit_kernel.cuh).  Outputs are
// host arrays and optional: scores (n; needs dirs), norms (n), colsum (S).
template <int J>
static int launch_audit(bcg_ctx* ctx, const AuditArgs& a) {
  const int grid = ctx->sm_count * 8;
  audit_score_kernel<J><<<grid, 256, 0, ctx->stream>>>(a);
  CK(cudaGetLastError());
  return BCG_OK;
}

extern "C" int bcg_dataset_audit(bcg_dataset* ds, int32_t model, int32_t d, const double* theta, int32_t S,
                                 const double* Siginv, int32_t kind, const double* dirs, double* scores, double* norms,
                                 double* colsum) {
  if (!ds || !theta) return fail(BCG_ERR_ARG, "null argument");
  RET(use_device(ds->ctx));
  bcg_ctx* ctx = ds->ctx;
  if (d <= 0 || S <= 0 || S > 1024) return fail(BCG_ERR_ARG, "bad shape d=%d S=%d", d, S);
  if ((model == BCG_MODEL_POISSON ? d + 1 : d) > ds->zld) return fail(BCG_ERR_ARG, "dataset has too few columns");
  if (scores && !dirs) return fail(BCG_ERR_ARG, "scores need the direction(s)");
  const int ndir = kind == BCG_ALG_GIGA ? 2 : 1;
  std::vector<double> tT, coff;
  int kmodel = 0;
  RET(prepare_model(model, d, theta, S, Siginv, tT, coff, &kmodel));
  (void)kmodel;
  std::vector<double> tt(S, 0.);
  for (size_t i = 0; i < coff.size(); ++i) tt[i] = -2. * coff[i];
  const int64_t n = ds->n;
  if (n == 0) {
    if (colsum) memset(colsum, 0, (size_t)S * sizeof(double));
    return BCG_OK;
  }
  cudaStream_t st = ctx->stream;
  DevBuf<double> dT, dtt, dSi, ddir, dsc, dnr, dcs;
  CK(dT.alloc((size_t)d * S));
  CK(cudaMemcpyAsync(dT, tT.data(), (size_t)d * S * sizeof(double), cudaMemcpyHostToDevice, st));
  AuditArgs a;
  a.Z = ds->Z; a.n = n; a.zld = ds->zld; a.d = d; a.S = S; a.model = model; a.kind = kind;
  a.thetaT = dT; a.tt = nullptr; a.Siginv = nullptr; a.dirs = nullptr; a.scores = nullptr; a.norms = nullptr; a.colsum = nullptr;
  if (model == BCG_MODEL_GAUSSIAN) {
    CK(dtt.alloc(S));
    CK(dSi.alloc((size_t)d * d));
    CK(cudaMemcpyAsync(dtt, tt.data(), (size_t)S * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dSi, Siginv, (size_t)d * d * sizeof(double), cudaMemcpyHostToDevice, st));
    a.tt = dtt; a.Siginv = dSi;
  }
  if (dirs) {
    CK(ddir.alloc((size_t)ndir * S));
    CK(cudaMemcpyAsync(ddir, dirs, (size_t)ndir * S * sizeof(double), cudaMemcpyHostToDevice, st));
    a.dirs = ddir;
  }
  if (scores) { CK(dsc.alloc((size_t)n)); a.scores = dsc; }
  if (norms) { CK(dnr.alloc((size_t)n)); a.norms = dnr; }
  if (colsum) { CK(dcs.alloc(S)); CK(cudaMemsetAsync(dcs, 0, (size_t)S * sizeof(double), st)); a.colsum = dcs; }
  switch (pow2ceil((S + 31) / 32)) {
    case 1: RET(launch_audit<1>(ctx, a)); break;
    case 2: RET(launch_audit<2>(ctx, a)); break;
    case 4: RET(launch_audit<4>(ctx, a)); break;
    case 8: RET(launch_audit<8>(ctx, a)); break;
    case 16: RET(launch_audit<16>(ctx, a)); break;
    default: RET(launch_audit<32>(ctx, a)); break;
  }
  if (scores) CK(cudaMemcpyAsync(scores, dsc, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (norms) CK(cudaMemcpyAsync(norms, dnr, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (colsum) CK(cudaMemcpyAsync(colsum, dcs, (size_t)S * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BCG_OK;
}

// The never-materialising projection: nothing but the row norms, the column sums b and the zero-row count are computed
// (one float64 pass of the audit kernel); the returned bcg_vecs has no matrix.  A solver created over it re-evaluates the
// rows from the dataset at every selection pass (lazy_select_kernel.cuh).  `ds` must outlive the result.
extern "C" int bcg_dataset_project_lazy(bcg_dataset* ds, int32_t model, int32_t d, const double* theta, int32_t S,
                                        const double* Siginv, bcg_vecs** out) {
  if (!ds || !theta || !out) return fail(BCG_ERR_ARG, "null argument");
  *out = nullptr;
  RET(use_device(ds->ctx));
  bcg_ctx* ctx = ds->ctx;
  if (d <= 0 || S <= 0) return fail(BCG_ERR_ARG, "d and S must be positive");
  if ((model == BCG_MODEL_POISSON ? d + 1 : d) > ds->zld) return fail(BCG_ERR_ARG, "dataset has too few columns");
  std::vector<double> tT, coff;
  int kmodel = 0;
  RET(prepare_model(model, d, theta, S, Siginv, tT, coff, &kmodel));
  bcg_vecs* v = nullptr;
  RET(vecs_alloc(ctx, ds->n, S, &v, true));
  auto body = [&]() -> int {
    cudaStream_t st = ctx->stream;
    v->lazy_ds = ds; v->lazy_model = model; v->lazy_d = d;
    CK(cudaMalloc(&v->lazy_thetaT, (size_t)d * S * sizeof(double)));
    CK(cudaMemcpyAsync(v->lazy_thetaT, tT.data(), (size_t)d * S * sizeof(double), cudaMemcpyHostToDevice, st));
    if (model == BCG_MODEL_GAUSSIAN) {
      std::vector<double> tt(S);
      for (int i = 0; i < S; ++i) tt[i] = -2. * coff[i];
      CK(cudaMalloc(&v->lazy_tt, (size_t)S * sizeof(double)));
      CK(cudaMalloc(&v->lazy_Siginv, (size_t)d * d * sizeof(double)));
      CK(cudaMemcpyAsync(v->lazy_tt, tt.data(), (size_t)S * sizeof(double), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(v->lazy_Siginv, Siginv, (size_t)d * d * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    CK(cudaStreamSynchronize(st));
    if (ds->n == 0) return BCG_OK;
    DevBuf<double> dcs;
    CK(dcs.alloc(S));
    CK(cudaMemsetAsync(dcs, 0, (size_t)S * sizeof(double), st));
    AuditArgs a;
    a.Z = ds->Z; a.n = ds->n; a.zld = ds->zld; a.d = d; a.S = S; a.model = model; a.kind = BCG_ALG_FW;
    a.thetaT = v->lazy_thetaT; a.tt = v->lazy_tt; a.Siginv = v->lazy_Siginv; a.dirs = nullptr; a.scores = nullptr;
    a.norms = v->norms; a.colsum = dcs;
    switch (pow2ceil((S + 31) / 32)) {
      case 1: RET(launch_audit<1>(ctx, a)); break;
      case 2: RET(launch_audit<2>(ctx, a)); break;
      case 4: RET(launch_audit<4>(ctx, a)); break;
      case 8: RET(launch_audit<8>(ctx, a)); break;
      case 16: RET(launch_audit<16>(ctx, a)); break;
      default: RET(launch_audit<32>(ctx, a)); break;
    }
    std::vector<double> nr((size_t)ds->n);
    CK(cudaMemcpyAsync(v->colsum.data(), dcs, (size_t)S * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nr.data(), v->norms, (size_t)ds->n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    double ns = 0.;
    uint64_t z = 0;
    for (double x : nr) { ns += x; z += !(x > 0.) ? 1 : 0; }
    v->colsum[S] = ns;
    v->zero_rows = z;
    return BCG_OK;
  };
  const int rc = body();
  if (rc != BCG_OK) { bcg_vecs_destroy(v); return rc; }
  *out = v;
  return BCG_OK;
}

// (K, S, dz) datapoint gradients of the K pseudo-points and / or their contraction with (w, resid) -- pseudo_grad_kernel.cuh.
// pts: host K x zld; theta: host S x d; outputs: host arrays, either may be null.
extern "C" int bcg_pseudo_grad(bcg_ctx* ctx, int32_t model, const double* pts, int64_t K, int32_t zld, int32_t d,
                               const double* theta, int32_t S, const double* Siginv, const double* w, const double* resid,
                               double* glls, double* ugrad) {
  RET(use_device(ctx));
  if (!pts || !theta || K < 0 || d <= 0 || S <= 0) return fail(BCG_ERR_ARG, "bad arguments");
  if (ugrad && (!w || !resid)) return fail(BCG_ERR_ARG, "the contraction needs w and resid");
  if (model == BCG_MODEL_GAUSSIAN && !Siginv) return fail(BCG_ERR_ARG, "Siginv is required for the Gaussian model");
  if ((model == BCG_MODEL_POISSON ? d + 1 : d) > zld) return fail(BCG_ERR_ARG, "points have too few columns");
  if (K == 0) return BCG_OK;
  const int dz = model == BCG_MODEL_POISSON ? d + 1 : d;
  // glls[k,s,:] = g_ks U_s + V_k, centred over the last axis (projector.py:26): centre the rows of U and V
  std::vector<double> U((size_t)S * dz, 0.), V;
  if (model == BCG_MODEL_GAUSSIAN) {
    V.assign((size_t)K * dz, 0.);
    for (int s = 0; s < S; ++s)
      for (int j = 0; j < d; ++j) {
        double acc = 0.;
        for (int i = 0; i < d; ++i) acc += theta[(size_t)s * d + i] * Siginv[(size_t)i * d + j];
        U[(size_t)s * dz + j] = acc;
      }
    for (int64_t k = 0; k < K; ++k)
      for (int j = 0; j < d; ++j) {
        double acc = 0.;
        for (int i = 0; i < d; ++i) acc += pts[(size_t)k * zld + i] * Siginv[(size_t)i * d + j];
        V[(size_t)k * dz + j] = -acc;
      }
  } else {
    for (int s = 0; s < S; ++s)
      for (int j = 0; j < d; ++j) U[(size_t)s * dz + j] = theta[(size_t)s * d + j];      // Poisson: zero d/dy column
  }
  auto centre = [&](std::vector<double>& M, int64_t rows) {
    for (int64_t r = 0; r < rows; ++r) {
      double m = 0.;
      for (int j = 0; j < dz; ++j) m += M[(size_t)r * dz + j];
      m /= (double)dz;
      for (int j = 0; j < dz; ++j) M[(size_t)r * dz + j] -= m;
    }
  };
  centre(U, S);
  if (!V.empty()) centre(V, K);
  cudaStream_t st = ctx->stream;
  DevBuf<double> dP, dT, dU, dV, dW, dR, dG, dO;
  auto up = [&](DevBuf<double>& b, const double* src, size_t n) -> int {
    CK(b.alloc(n));
    CK(cudaMemcpyAsync(b, src, n * sizeof(double), cudaMemcpyHostToDevice, st));
    return BCG_OK;
  };
  RET(up(dP, pts, (size_t)K * zld));
  RET(up(dT, theta, (size_t)S * d));
  RET(up(dU, U.data(), U.size()));
  if (!V.empty()) RET(up(dV, V.data(), V.size()));
  if (ugrad) { RET(up(dW, w, (size_t)K)); RET(up(dR, resid, (size_t)S)); CK(dO.alloc((size_t)K * dz)); }
  if (glls) CK(dG.alloc((size_t)K * S * dz));
  PseudoGradArgs a;
  a.pts = dP; a.theta = dT; a.Uc = dU; a.Vc = V.empty() ? nullptr : dV.p; a.w = dW; a.resid = dR; a.glls = dG; a.ugrad = dO;
  a.K = (int32_t)K; a.zld = zld; a.d = d; a.dz = dz; a.S = S; a.model = model;
  const size_t smem = ((size_t)S + d + 2) * sizeof(double);
  CK(cudaFuncSetAttribute(pseudo_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pseudo_grad_kernel<<<(unsigned)K, 256, smem, st>>>(a);
  CK(cudaGetLastError());
  if (glls) CK(cudaMemcpyAsync(glls, dG, (size_t)K * S * dz * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (ugrad) CK(cudaMemcpyAsync(ugrad, dO, (size_t)K * dz * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BCG_OK;
}

// Weighted Gaussian posterior sampler on the device (sampler_kernels.cuh): theta = mup + E U^T with (mup, U) =
// weighted_post(th0, Sig0inv, Siginv, pts, w) (model_gaussian.py:23-30; gaussian/main.py:107-113).  E: host S x d standard
// normals drawn by the caller.  Outputs (host): theta S x d, optionally mup (d) and U (d x d, upper: Sigp = U U^T).
extern "C" int bcg_sampler_gaussian_post(bcg_ctx* ctx, int32_t d, const double* th0, const double* Sig0inv, const double* Siginv,
                                         const double* pts, const double* w, int64_t K, const double* E, int32_t S, double* theta,
                                         double* mup, double* U) {
  RET(use_device(ctx));
  if (!th0 || !Sig0inv || !Siginv || d <= 0 || K < 0 || S < 0 || (K > 0 && (!pts || !w)) || (S > 0 && (!E || !theta)))
    return fail(BCG_ERR_ARG, "bad arguments");
  cudaStream_t st = ctx->stream;
  // one carve-out of a context-owned scratch slot (no cudaMalloc / cudaFree per call: SparseVI calls this 101 times per
  // build iteration)
  const size_t dd = (size_t)d * d;
  const size_t n_doubles = d + 2 * dd + (size_t)K * d + K + (size_t)S * d + 2 * dd + 2 * d + d + (size_t)S * d + 2;
  double* base = nullptr;
  RET(ctx_scratch(ctx, 0, n_doubles * sizeof(double), (void**)&base));
  double* cur = base;
  auto take = [&](size_t n) { double* p = cur; cur += n; return p; };
  auto up = [&](double** dst, const double* src, size_t n) -> int {
    *dst = take(n);
    if (n) CK(cudaMemcpyAsync(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice, st));
    return BCG_OK;
  };
  double *dth0, *dS0, *dS, *dP, *dW, *dE;
  RET(up(&dth0, th0, (size_t)d));
  RET(up(&dS0, Sig0inv, dd));
  RET(up(&dS, Siginv, dd));
  RET(up(&dP, pts, (size_t)K * d));
  RET(up(&dW, w, (size_t)K));
  RET(up(&dE, E, (size_t)S * d));
  double* dL = take(dd); double* dLi = take(dd); double* dR = take((size_t)2 * d); double* dM = take((size_t)d);
  double* dT = take((size_t)S * d);
  int32_t* dStat = reinterpret_cast<int32_t*>(take(1));
  CK(cudaMemsetAsync(dStat, 0, sizeof(int32_t), st));
  GaussPostArgs a;
  a.th0 = dth0; a.Sig0inv = dS0; a.Siginv = dS; a.pts = dP; a.w = dW; a.E = dE; a.L = dL; a.Linv = dLi; a.rhs = dR; a.mup = dM;
  a.theta = dT; a.d = d; a.K = (int32_t)K; a.S = S; a.status = dStat;
  gauss_post_factor_kernel<<<1, 1024, 0, st>>>(a);
  CK(cudaGetLastError());
  if (S > 0) {
    const int64_t tot = (int64_t)S * d;
    gauss_post_sample_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(a);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(theta, dT, (size_t)tot * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  int32_t stat = 0;
  CK(cudaMemcpyAsync(&stat, dStat, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (mup) CK(cudaMemcpyAsync(mup, dM, (size_t)d * sizeof(double), cudaMemcpyDeviceToHost, st));
  std::vector<double> li;
  if (U) { li.resize((size_t)d * d); CK(cudaMemcpyAsync(li.data(), dLi, (size_t)d * d * sizeof(double), cudaMemcpyDeviceToHost, st)); }
  CK(cudaStreamSynchronize(st));
  if (stat) return fail(BCG_ERR_ARG, "Sig0inv + sum(w) Siginv is not positive definite");
  if (U)
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) U[(size_t)i * d + j] = li[(size_t)j * d + i];       // U = (L^-1)^T
  return BCG_OK;
}

// Weighted log-joint, gradient and Hessian of the GLM example models over K points (sampler_kernels.cuh: glm_joint_kernel):
// the reductions of the Laplace sampler.  Z: host K x zld, w: host K, theta: host d; outputs host (any may be null).
extern "C" int bcg_glm_joint(bcg_ctx* ctx, int32_t model, const double* Z, const double* w, int64_t K, int32_t zld, int32_t d,
                             const double* theta, double* value, double* grad, double* hess) {
  RET(use_device(ctx));
  if (!theta || d <= 0 || K < 0 || (K > 0 && (!Z || !w))) return fail(BCG_ERR_ARG, "bad arguments");
  if (model != BCG_MODEL_LR && model != BCG_MODEL_POISSON) return fail(BCG_ERR_ARG, "the Laplace reductions exist for the LR and Poisson models");
  if ((model == BCG_MODEL_POISSON ? d + 1 : d) > zld && K > 0) return fail(BCG_ERR_ARG, "points have too few columns");
  const size_t smem = ((size_t)d + 3 * (size_t)K) * sizeof(double);
  if (smem > 200 * 1024) return fail(BCG_ERR_UNSUPPORTED, "too many points for the single-CTA Laplace reduction (K = %lld)", (long long)K);
  cudaStream_t st = ctx->stream;
  const size_t dd = (size_t)d * d;
  double* base = nullptr;
  RET(ctx_scratch(ctx, 0, ((size_t)K * zld + K + d + 1 + d + dd + 4) * sizeof(double), (void**)&base));
  double* dZ = base; double* dW = dZ + (size_t)K * zld; double* dTh = dW + K; double* dV = dTh + d; double* dG = dV + 1; double* dH = dG + d;
  if (K > 0) {
    CK(cudaMemcpyAsync(dZ, Z, (size_t)K * zld * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dW, w, (size_t)K * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemcpyAsync(dTh, theta, (size_t)d * sizeof(double), cudaMemcpyHostToDevice, st));
  GlmJointArgs a;
  a.Z = dZ; a.w = dW; a.theta = dTh; a.value = dV; a.grad = dG; a.hess = dH; a.K = (int32_t)K; a.zld = zld; a.d = d; a.model = model;
  CK(cudaFuncSetAttribute(glm_joint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  glm_joint_kernel<<<1, 256, smem, st>>>(a);
  CK(cudaGetLastError());
  if (value) CK(cudaMemcpyAsync(value, dV, sizeof(double), cudaMemcpyDeviceToHost, st));
  if (grad) CK(cudaMemcpyAsync(grad, dG, (size_t)d * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (hess) CK(cudaMemcpyAsync(hess, dH, dd * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BCG_OK;
}

// Projection straight from a HOST array, chunked and software-pipelined: while chunk c is being projected on
// `stream`, chunk c+1 is staged (multi-threaded memcpy into pinned memory) and copied on `copy_stream`.  The
// data are not kept on the device (HilbertCoreset projects once).  thetaT: d x S, coff: S or null (host).
static int project_host_pipelined(bcg_ctx* ctx, int kmodel, const double* Z, int64_t n, int32_t zld, int32_t d,
                                  const double* thetaT, const double* coff, int32_t S, bcg_vecs** out) {
  *out = nullptr;
  if (d <= 0 || S <= 0) return fail(BCG_ERR_ARG, "d and S must be positive");
  if ((kmodel == MODEL_POISSON ? d + 1 : d) > zld) return fail(BCG_ERR_ARG, "data has too few columns");
  bcg_vecs* v = nullptr;
  RET(vecs_alloc(ctx, n, S, &v));
  if (n == 0) { *out = v; return BCG_OK; }
  uint16_t* an16 = nullptr;
  auto body = [&]() -> int {
    cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
    const int ld = v->ld;
    const size_t cs_bytes = (size_t)kProjWarps * (S + 1) * sizeof(double);
    const size_t budget = 200 * 1024;
    if (cs_bytes + (size_t)S * sizeof(double) > budget)
      return fail(BCG_ERR_UNSUPPORTED, "projection tile does not fit shared memory for S=%d", S);
    const int ktile = (int)std::min<size_t>(std::min<size_t>(kProjKTile, (size_t)d), (budget - cs_bytes) / ((size_t)S * sizeof(double)));
    const size_t smem = (size_t)ktile * S * sizeof(double) + cs_bytes;
    const int64_t row_bytes = (int64_t)zld * 8;
    const int64_t rows_cap = (int64_t)kPinChunk / row_bytes;             // rows that fit one staging buffer
    if (rows_cap < 1) return fail(BCG_ERR_UNSUPPORTED, "a data row of %d doubles exceeds the staging chunk", zld);
    const int64_t chunk_rows = std::min<int64_t>(n, std::min<int64_t>(rows_cap, std::max<int64_t>(4096, ((int64_t)16 << 20) / row_bytes)));
    const int nchunks = (int)((n + chunk_rows - 1) / chunk_rows);
    ProjectArgs proto;                                                   // what decides the kernel and its grid
    proto.An = v->An; proto.out64 = nullptr; proto.n = chunk_rows; proto.d = d; proto.S = S; proto.ld = ld; proto.model = kmodel;
    proto.sp_tab = ctx->sp_tab; proto.ktile = ktile;
    const int grid = project_grid(ctx, proto);
    // float16 copy for the pre-filter of the persistent greedy kernels, written by the projection itself where the kernel
    // supports it (otherwise the solver makes it with one more pass: vecs_ensure_half)
    if (project_fast_ok(proto) && filter16_shape_ok(ld)) {
      const size_t bytes = (size_t)n * v->ld16 * sizeof(uint16_t);
      if (half_take(ctx, bytes, &an16)) v->An16_bytes = bytes;
    }
    // staging and chunk buffers are context-owned (no cudaMallocHost / cudaMalloc / cudaFree per call)
    const bool direct = is_pinned(Z);                                    // page-locked source: no staging copy
    if (!direct) RET(ensure_pins(ctx));
    double* pin[2] = {reinterpret_cast<double*>(ctx->pin[0]), reinterpret_cast<double*>(ctx->pin[1])};
    double* dev[2] = {nullptr, nullptr};
    DevBuf<double> dT, dC, d_partial;
    DevBuf<unsigned long long> d_zero;
    EventPair copied, kdone;
    const size_t celems = (size_t)chunk_rows * zld;
    for (int i = 0; i < 2; ++i) {
      RET(ctx_scratch(ctx, 6 + i, celems * sizeof(double), (void**)&dev[i]));
      CK(cudaEventCreateWithFlags(&copied.e[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&kdone.e[i], cudaEventDisableTiming));
    }
    CK(dT.alloc((size_t)d * S));
    CK(d_partial.alloc((size_t)nchunks * grid * (S + 1)));
    CK(d_zero.alloc(1));
    CK(cudaMemsetAsync(d_zero, 0, sizeof(unsigned long long), st));
    CK(cudaMemcpyAsync(dT, thetaT, (size_t)d * S * sizeof(double), cudaMemcpyHostToDevice, st));
    if (coff) {
      CK(dC.alloc(S));
      CK(cudaMemcpyAsync(dC, coff, (size_t)S * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    for (int c = 0; c < nchunks; ++c) {
      const int i = c & 1;
      const int64_t r0 = (int64_t)c * chunk_rows;
      const int64_t nr = std::min<int64_t>(chunk_rows, n - r0);
      const double* src = Z + r0 * zld;
      if (c >= 2) {
        CK(cudaStreamWaitEvent(cs, kdone.e[i], 0));                  // device buffer i is free again: ordered on the
        if (!direct) CK(cudaEventSynchronize(copied.e[i]));          // device; the host waits only for ITS staging buffer
      }
      if (!direct) { parallel_memcpy(pin[i], src, (size_t)nr * zld * sizeof(double)); src = pin[i]; }
      CK(cudaMemcpyAsync(dev[i], src, (size_t)nr * zld * sizeof(double), cudaMemcpyHostToDevice, cs));
      CK(cudaEventRecord(copied.e[i], cs));
      CK(cudaStreamWaitEvent(st, copied.e[i], 0));
      ProjectArgs a;
      a.Z = dev[i]; a.rowidx = nullptr; a.theta = dT; a.coff = dC; a.An = v->An + (size_t)r0 * ld; a.norms = v->norms + r0;
      a.out64 = nullptr; a.partial = d_partial.p + (size_t)c * grid * (S + 1); a.zero_rows = d_zero; a.n = nr; a.zld = zld;
      a.d = d; a.S = S; a.ld = ld; a.model = kmodel; a.ktile = ktile; a.sp_tab = ctx->sp_tab;
      a.An16 = an16 ? an16 + (size_t)r0 * v->ld16 : nullptr; a.ld16 = v->ld16;
      RET(dispatch_project(ctx, a, grid, smem));
      CK(cudaEventRecord(kdone.e[i], st));
    }
    RET(finish_colsum(v, d_partial, nchunks * grid, d_zero));
    CK(cudaStreamSynchronize(cs));
    v->An16 = an16;
    an16 = nullptr;
    return BCG_OK;
  };
  const int rc = body();
  if (rc != BCG_OK) {
    if (an16) { cudaStreamSynchronize(ctx->stream); cudaFree(an16); v->An16_bytes = 0; }
    bcg_vecs_destroy(v);
    return rc;
  }
  *out = v;
  return BCG_OK;
}

static int project_host(bcg_ctx* ctx, int model, const double* Z, int64_t n, int32_t zld, int32_t d, const double* theta,
                        int32_t S, const double* Siginv, bcg_vecs** out) {
  RET(use_device(ctx));
  if (!out || !theta || (n > 0 && !Z)) return fail(BCG_ERR_ARG, "null argument");
  std::vector<double> tT, coff;
  int kmodel = 0;
  RET(prepare_model(model, d, theta, S, Siginv, tT, coff, &kmodel));
  return project_host_pipelined(ctx, kmodel, Z, n, zld, d, tT.data(), coff.empty() ? nullptr : coff.data(), S, out);
}

// Gaussian model with A = theta Siginv (host S x d) and coff_s = -0.5 theta_s Siginv theta_s precomputed by the caller
extern "C" int bcg_vecs_project_linear(bcg_ctx* ctx, const double* x, int64_t n, int32_t d, const double* A,
                                       const double* coff, int32_t S, bcg_vecs** out) {
  RET(use_device(ctx));
  if (!out || !A || (n > 0 && !x)) return fail(BCG_ERR_ARG, "null argument");
  std::vector<double> tT = transpose_sd(A, S, d);
  return project_host_pipelined(ctx, MODEL_LINEAR, x, n, d, d, tT.data(), coff, S, out);
}

extern "C" int bcg_vecs_project_lr(bcg_ctx* ctx, const double* Z, int64_t n, int32_t d, const double* theta,
                                   int32_t S, bcg_vecs** out) {
  return project_host(ctx, BCG_MODEL_LR, Z, n, d, d, theta, S, nullptr, out);
}

extern "C" int bcg_vecs_project_gaussian(bcg_ctx* ctx, const double* x, int64_t n, int32_t d, const double* theta,
                                         int32_t S, const double* Siginv, bcg_vecs** out) {
  return project_host(ctx, BCG_MODEL_GAUSSIAN, x, n, d, d, theta, S, Siginv, out);
}

extern "C" int bcg_vecs_project_poisson(bcg_ctx* ctx, const double* Z, int64_t n, int32_t d, const double* theta,
                                        int32_t S, bcg_vecs** out) {
  return project_host(ctx, BCG_MODEL_POISSON, Z, n, d + 1, d, theta, S, nullptr, out);
}

extern "C" int bcg_vecs_shape(bcg_vecs* v, int64_t* n, int32_t* S, int32_t* ld) {
  if (!v) return fail(BCG_ERR_ARG, "null vecs");
  if (n) *n = v->n;
  if (S) *S = v->S;
  if (ld) *ld = v->ld;
  return BCG_OK;
}

extern "C" int bcg_vecs_colsum(bcg_vecs* v, double* out_S) {
  if (!v || !out_S) return fail(BCG_ERR_ARG, "null argument");
  memcpy(out_S, v->colsum.data(), (size_t)v->S * sizeof(double));
  return BCG_OK;
}

extern "C" int bcg_vecs_norm_sum(bcg_vecs* v, double* out) {
  if (!v || !out) return fail(BCG_ERR_ARG, "null argument");
  *out = v->colsum[v->S];
  return BCG_OK;
}

extern "C" int bcg_vecs_zero_rows(bcg_vecs* v, int64_t* count) {
  if (!v || !count) return fail(BCG_ERR_ARG, "null argument");
  *count = (int64_t)v->zero_rows;
  return BCG_OK;
}

extern "C" int bcg_vecs_norms(bcg_vecs* v, int64_t row0, int64_t nrows, double* out) {
  if (!v || !out) return fail(BCG_ERR_ARG, "null argument");
  if (row0 < 0 || nrows < 0 || row0 + nrows > v->n) return fail(BCG_ERR_ARG, "row range out of bounds");
  RET(use_device(v->ctx));
  if (nrows == 0) return BCG_OK;
  CK(cudaMemcpyAsync(out, v->norms + row0, (size_t)nrows * sizeof(double), cudaMemcpyDeviceToHost, v->ctx->stream));
  CK(cudaStreamSynchronize(v->ctx->stream));
  return BCG_OK;
}

extern "C" int bcg_vecs_rows_f64(bcg_vecs* v, int64_t row0, int64_t nrows, double* out) {
  if (!v || !out) return fail(BCG_ERR_ARG, "null argument");
  if (row0 < 0 || nrows < 0 || row0 + nrows > v->n) return fail(BCG_ERR_ARG, "row range out of bounds");
  if (v->n > 0 && !v->An) return fail(BCG_ERR_UNSUPPORTED, "a never-materialised projection has no stored rows");
  RET(use_device(v->ctx));
  const int64_t chunk = std::max<int64_t>(1, (64ll << 20) / ((int64_t)v->S * 8));
  DevBuf<double> tmp;
  CK(tmp.alloc((size_t)std::min(chunk, std::max<int64_t>(nrows, 1)) * v->S));
  for (int64_t r = 0; r < nrows; r += chunk) {
    const int64_t nr = std::min(chunk, nrows - r);
    const int64_t tot = nr * v->S;
    expand_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, v->ctx->stream>>>(v->An, v->norms, row0 + r, nr, v->S,
                                                                               v->ld, tmp);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out + r * v->S, tmp, (size_t)tot * sizeof(double), cudaMemcpyDeviceToHost, v->ctx->stream));
    CK(cudaStreamSynchronize(v->ctx->stream));
  }
  return BCG_OK;
}

extern "C" int bcg_vecs_destroy(bcg_vecs* v) {
  if (!v) return BCG_OK;
  cudaSetDevice(v->ctx->device);
  cudaStreamSynchronize(v->ctx->stream);
  pool_give(&v->ctx->pool_An, &v->ctx->pool_An_bytes, v->An, v->An_bytes);
  pool_give(&v->ctx->pool_norms, &v->ctx->pool_norms_bytes, v->norms, v->norms_bytes);
  pool_give(&v->ctx->pool_An16, &v->ctx->pool_An16_bytes, v->An16, v->An16_bytes);
  if (v->lazy_thetaT) cudaFree(v->lazy_thetaT);
  if (v->lazy_tt) cudaFree(v->lazy_tt);
  if (v->lazy_Siginv) cudaFree(v->lazy_Siginv);
  delete v;
  return BCG_OK;
}

// float16 copy of the unit rows for the pre-filter of the persistent scan (filter_bounds.h): round to nearest, padding zero
__global__ void __launch_bounds__(256) half_copy_kernel(const float* __restrict__ An, int64_t n, int ld, uint16_t* __restrict__ out,
                                                        int ld16) {
  const int gpr = ld16 >> 3;                                  // 8-element groups per row
  const int64_t total = n * gpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / gpr;
    const int e0 = (int)(i - row * gpr) * 8;
    const float* src = An + row * ld + e0;
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; k += 4) {
      if (e0 + k < ld) {                                      // ld is a multiple of 4: whole float4 inside or outside the row
        const float4 t = __ldcs(reinterpret_cast<const float4*>(src + k));
        x[k] = t.x; x[k + 1] = t.y; x[k + 2] = t.z; x[k + 3] = t.w;
      } else {
        x[k] = x[k + 1] = x[k + 2] = x[k + 3] = 0.f;
      }
    }
    uint4 o;
    __half2 h;
    h = __floats2half2_rn(x[0], x[1]); o.x = *reinterpret_cast<unsigned int*>(&h);
    h = __floats2half2_rn(x[2], x[3]); o.y = *reinterpret_cast<unsigned int*>(&h);
    h = __floats2half2_rn(x[4], x[5]); o.z = *reinterpret_cast<unsigned int*>(&h);
    h = __floats2half2_rn(x[6], x[7]); o.w = *reinterpret_cast<unsigned int*>(&h);
    *reinterpret_cast<uint4*>(out + row * ld16 + e0) = o;
  }
}

// make v->An16 (idempotent).  Not an error when the memory is not there: the solver then streams the float32 rows.
static int vecs_ensure_half(bcg_vecs* v) {
  if (v->An16 || !v->An || v->n == 0) return BCG_OK;
  bcg_ctx* ctx = v->ctx;
  const size_t bytes = (size_t)v->n * v->ld16 * sizeof(uint16_t);
  uint16_t* p = nullptr;
  if (!half_take(ctx, bytes, &p)) return BCG_OK;
  v->An16 = p;
  v->An16_bytes = bytes;
  half_copy_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(v->An, v->n, v->ld, v->An16, v->ld16);
  CK(cudaGetLastError());
  return BCG_OK;
}

// ------------------------------------------------------------------------------------------
// solver
// ------------------------------------------------------------------------------------------
static int choose_scan_config(bcg_solver* s) {
  const int ld = s->v->ld;
  ScanConfig& c = s->sc;
  const int nchunk = ld / 4;
  c.lpr = std::min(32, pow2ceil(nchunk));
  c.ch = pow2ceil((nchunk + c.lpr - 1) / c.lpr);
  if (!scan_variant_exists(c.ch, c.lpr))
    return fail(BCG_ERR_UNSUPPORTED, "no scan kernel for S=%d (ch=%d lpr=%d)", s->v->S, c.ch, c.lpr);
  c.r = scan_variant_r(c.ch, c.lpr);
  c.rb = c.r * (32 / c.lpr);
  c.ndir = (s->h.alg == BCG_ALG_GIGA) ? 2 : 1;
  const int row_bytes = ld * 4;
  const int batch_bytes = c.rb * row_bytes;
  const int stage_bytes = env_int("BCG_SCAN_STAGE_BYTES", 8192);
  const int nb = std::max(1, stage_bytes / batch_bytes);
  c.rps = nb * c.rb;
  // float16 pre-filter of the persistent kernels (filter_bounds.h): available for 128 < S <= 512
  c.ch16 = (env_int("BCG_FILTER16", 1) && s->v->An) ? loop_variant_ch16(c.ch, c.lpr) : 0;
  // loop kernel: <= 11 scan warps + 1 control warp.  The float16 pass executes 2.6 x the instructions per byte of the
  // float32 scan and needs the extra warps to stay HBM-bound (measured at N = 1e7, S = 512: 1.63 / 1.51 / 1.455 ms per
  // iteration with 8 / 10 / 11 warps; float32 stream: 2.78 ms)
  c.wpb = std::max(1, std::min(11, env_int("BCG_SCAN_WARPS", c.ch16 ? 11 : 8)));
  c.stages = std::max(1, env_int("BCG_SCAN_STAGES", 2));
  c.evict_first = env_int("BCG_SCAN_EVICT_FIRST", 0);
  const size_t budget = (size_t)(200 * 1024);
  const size_t extra = 64 * sizeof(ScanCand) + 4 * (size_t)s->v->S * sizeof(double) + 64 + 16 * 8 * 12 + 128 +
                       11 * 32 * 8;                                  // loop kernel only (last term: re-scan slots of the float16 pre-filter)
  auto ring = [&](int stages, int rps) { return (size_t)c.wpb * stages * rps * row_bytes + (size_t)c.wpb * stages * 8; };
  while (c.stages > 2 && ring(c.stages, c.rps) + extra > budget) --c.stages;
  while (c.rps > c.rb && ring(c.stages, c.rps) + extra > budget) c.rps -= c.rb;
  while (c.wpb > 1 && ring(c.stages, c.rps) + extra > budget) --c.wpb;
  while (c.stages > 1 && ring(c.stages, c.rps) + extra > budget) --c.stages;
  c.smem = ring(c.stages, c.rps);
  c.loop_smem = c.smem + extra;
  if (c.loop_smem > 227 * 1024) return fail(BCG_ERR_UNSUPPORTED, "scan tile does not fit shared memory (ld=%d)", ld);
  c.grid = s->ctx->sm_count;
  // rows per ring stage in the float16 pass (multiple of its batch: 8 rows when a lane owns one 16-byte group per row, else 4)
  const int fr = c.ch16 == 1 ? 8 : 4;
  c.rps16 = (int)((size_t)c.rps * row_bytes / ((size_t)s->v->ld16 * 2)) / fr * fr;
  if (c.rps16 < fr) c.ch16 = 0;
  CK(scan_set_smem(c));
  s->use_loop = false;
  s->use_omp_loop = false;
  if (env_int("BCG_ENGINE", 2) >= 2 && loop_variant_exists(c.ch, c.lpr)) {
    int coop = 0, nbm = 0;
    CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, s->ctx->device));
    if (s->h.alg != BCG_ALG_OMP) {
      CK(loop_set_smem(c));
      CK(loop_max_blocks_per_sm(c, &nbm));
      s->use_loop = coop && nbm >= 1;
    } else if (env_int("BCG_OMP_LOOP", 1) && c.grid >= 2) {
      CK(omp_loop_set_smem(c));
      CK(omp_loop_max_blocks_per_sm(c, &nbm));
      s->use_omp_loop = coop && nbm >= 1;
    }
  }
  s->filter16 = (s->use_loop || s->use_omp_loop) && c.ch16 > 0;
  if (s->filter16) RET(vecs_ensure_half(s->v));
  return BCG_OK;
}

template <int J>
static int launch_lazy(bcg_solver* s, const LazyArgs& L) {
  lazy_select_kernel<J><<<s->h.n_exact_cands, 256, 0, s->ctx->stream>>>(L);
  CK(cudaGetLastError());
  return BCG_OK;
}

static int launch_scan(bcg_solver* s, cudaEvent_t e0 = nullptr, cudaEvent_t e1 = nullptr) {
  if (s->h.lazy) {
    // never-materialising solver: the selection pass re-evaluates every row from the raw data (lazy_select_kernel.cuh)
    const bcg_vecs* v = s->v;
    LazyArgs L;
    L.a.Z = v->lazy_ds->Z; L.a.n = v->n; L.a.zld = v->lazy_ds->zld; L.a.d = v->lazy_d; L.a.S = v->S; L.a.model = v->lazy_model;
    L.a.kind = s->h.alg; L.a.thetaT = v->lazy_thetaT; L.a.tt = v->lazy_tt; L.a.Siginv = v->lazy_Siginv;
    L.a.dirs = nullptr; L.a.scores = nullptr; L.a.norms = nullptr; L.a.colsum = nullptr;
    L.st = s->d;
    if (e0) CK(cudaEventRecord(e0, s->ctx->stream));
    switch (pow2ceil((v->S + 31) / 32)) {
      case 1: RET(launch_lazy<1>(s, L)); break;
      case 2: RET(launch_lazy<2>(s, L)); break;
      case 4: RET(launch_lazy<4>(s, L)); break;
      case 8: RET(launch_lazy<8>(s, L)); break;
      case 16: RET(launch_lazy<16>(s, L)); break;
      default: RET(launch_lazy<32>(s, L)); break;
    }
    if (e1) CK(cudaEventRecord(e1, s->ctx->stream));
    return BCG_OK;
  }
  ScanArgs a;
  a.g.An = s->v->An;
  a.g.n_rows = s->v->n;
  a.g.ld = s->v->ld;
  a.g.rps = s->sc.rps;
  a.g.stages = s->sc.stages;
  a.g.evict_first = s->sc.evict_first;
  a.dir = s->h.dir32;
  a.cands = s->h.cands;
  a.skip0 = &s->d->halted;
  a.skip1 = &s->d->select_failed;
  a.lost = s->h.cand_lost;
  a.done = &s->d->scan_done;
  a.need_exact = &s->d->need_exact;
  a.force_exact = &s->d->force_exact;
  a.top = &s->d->scan_top;
  a.top_row = &s->d->scan_top_row;
  a.top_cnt = &s->d->scan_cnt;
  if (e0) CK(cudaEventRecord(e0, s->ctx->stream));
  CK(scan_launch(s->sc, a, s->ctx->stream));
  if (e1) CK(cudaEventRecord(e1, s->ctx->stream));
  // exact float64 selection pass: returns at once unless the scan's last CTA found the candidate set ambiguous
  CK(exact_scan_launch(s->h.n_exact_cands, s->d, 0, s->ctx->stream));
  return BCG_OK;
}

static int push_state(bcg_solver* s) {
  CK(cudaMemcpyAsync(s->d, &s->h, sizeof(SolverState), cudaMemcpyHostToDevice, s->ctx->stream));
  return BCG_OK;
}
static int pull_state(bcg_solver* s) {
  CK(cudaMemcpyAsync(&s->h, s->d, sizeof(SolverState), cudaMemcpyDeviceToHost, s->ctx->stream));
  CK(cudaStreamSynchronize(s->ctx->stream));
  return BCG_OK;
}

template <typename T>
static int grow(T** p, size_t old_n, size_t new_n, cudaStream_t st) {
  T* q = nullptr;
  CK(cudaMalloc(&q, new_n * sizeof(T)));
  CK(cudaMemsetAsync(q, 0, new_n * sizeof(T), st));
  if (*p && old_n) CK(cudaMemcpyAsync(q, *p, old_n * sizeof(T), cudaMemcpyDeviceToDevice, st));
  CK(cudaStreamSynchronize(st));
  if (*p) CK(cudaFree(*p));
  *p = q;
  return BCG_OK;
}

// make room for `extra` more stored active rows (host mirror must be current)
static int ensure_capacity(bcg_solver* s, int extra) {
  SolverState& h = s->h;
  const int need = h.nact + extra;
  if (need <= h.cap) return BCG_OK;
  const int ncap = std::max(need, std::max(64, h.cap * 2));
  cudaStream_t st = s->ctx->stream;
  RET(grow(&h.act_idx, (size_t)h.cap, (size_t)ncap, st));
  RET(grow(&h.act_w, (size_t)h.cap, (size_t)ncap, st));
  RET(grow(&h.act_w_new, (size_t)h.cap, (size_t)ncap, st));
  RET(grow(&h.act_norm, (size_t)h.cap, (size_t)ncap, st));
  RET(grow(&h.act_tmp, (size_t)h.cap, (size_t)ncap, st));
  RET(grow(&h.act_rows, (size_t)h.cap * h.ld, (size_t)ncap * h.ld, st));
  h.cap = ncap;
  return BCG_OK;
}

static int solver_init(bcg_solver* s, bcg_ctx* ctx, bcg_vecs* v, int32_t alg, const double* b, double bnorm, double norm_sum,
                       int64_t row_offset, int64_t n_global) {
  const int S = v->S, ld = v->ld;
  SolverState& h = s->h;
  h.alg = alg; h.S = S; h.ld = ld; h.world = 1; h.rank = 0;
  h.n_local = v->n; h.row_offset = row_offset; h.n_global = n_global;
  h.tol = 1e-12; h.bnorm = bnorm; h.nsum = norm_sum;
  h.An = v->An; h.norms = v->norms;
  h.err = bnorm;
  RET(choose_scan_config(s));
  h.n_cands = s->sc.grid * s->sc.wpb;
  h.lazy = (v->n > 0 && !v->An) ? 1 : 0;
  h.n_exact_cands = h.lazy ? ctx->sm_count * 8 : ctx->sm_count;
  h.fused_row = -1;
  if (h.lazy) s->use_loop = s->use_omp_loop = false;     // the persistent kernels stream the resident matrix
  h.check_monotone = 1;
  cudaStream_t st = ctx->stream;
  std::vector<double> bn(S);
  for (int i = 0; i < S; ++i) bn[i] = bnorm == 0. ? 0. : b[i] / bnorm;   // (a non-finite b propagates, as in giga.py:18)
  CK(cudaMalloc(&h.b, S * sizeof(double)));
  CK(cudaMalloc(&h.bn, S * sizeof(double)));
  CK(cudaMalloc(&h.xw, S * sizeof(double)));
  CK(cudaMalloc(&h.xw_new, S * sizeof(double)));
  CK(cudaMalloc(&h.xf, S * sizeof(double)));
  CK(cudaMalloc(&h.dir64, 2 * S * sizeof(double)));
  CK(cudaMalloc(&h.dir32, 2 * ld * sizeof(float)));
  CK(cudaMalloc(&h.wrow, ld * sizeof(float)));
  CK(cudaMalloc(&h.cands, (size_t)h.n_cands * sizeof(ScanCand)));
  CK(cudaMalloc(&s->d_fout, 2 * sizeof(int64_t)));
  CK(cudaMalloc(&s->d, sizeof(SolverState)));
  CK(cudaMalloc(&s->d_ctl, sizeof(LoopCtl)));
  CK(cudaMalloc(&s->d_cta_cands, (size_t)2 * s->sc.grid * sizeof(ScanCand)));
  CK(cudaMalloc(&s->d_cta_lost, (size_t)s->sc.grid * sizeof(float)));
  CK(cudaMalloc(&h.cand_lost, (size_t)h.n_cands * sizeof(float)));
  CK(cudaMalloc(&h.exact_cands, (size_t)h.n_exact_cands * sizeof(ExactCand)));
  CK(cudaMemsetAsync(h.cand_lost, 0xff, (size_t)h.n_cands * sizeof(float), st));
  CK(cudaMemcpyAsync(h.b, b, S * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(h.bn, bn.data(), S * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(h.xw, 0, S * sizeof(double), st));
  CK(cudaMemsetAsync(h.dir32, 0, 2 * ld * sizeof(float), st));
  CK(cudaMemsetAsync(h.dir64, 0, 2 * S * sizeof(double), st));
  CK(cudaMemsetAsync(h.cands, 0xff, (size_t)h.n_cands * sizeof(ScanCand), st));
  CK(cudaStreamSynchronize(st));
  RET(ensure_capacity(s, 64));
  RET(push_state(s));
  CK(cudaEventCreate(&s->ev0));
  CK(cudaEventCreate(&s->ev1));
  CK(cudaStreamSynchronize(st));
  return BCG_OK;
}


extern "C" int bcg_solver_create(bcg_ctx* ctx, bcg_vecs* v, int32_t alg, const double* b, double norm_sum,
                                 int64_t row_offset, int64_t n_global, bcg_solver** out) {
  RET(use_device(ctx));
  if (!out || !v || !b) return fail(BCG_ERR_ARG, "null argument");
  if (alg != BCG_ALG_GIGA && alg != BCG_ALG_FW && alg != BCG_ALG_OMP) return fail(BCG_ERR_ARG, "unknown alg %d", alg);
  *out = nullptr;
  const int S = v->S;
  double bnorm = 0.;
  for (int i = 0; i < S; ++i) bnorm += b[i] * b[i];
  bnorm = sqrt(bnorm);
  if (alg == BCG_ALG_GIGA && bnorm == 0.) return fail(BCG_ERR_ZERO_B, "norm of b must be > 0");
  bcg_solver* s = new bcg_solver();
  memset(&s->h, 0, sizeof(SolverState));
  s->d = nullptr;
  s->d_fout = nullptr;
  s->ev0 = s->ev1 = nullptr;
  s->ctx = ctx;
  s->v = v;
  s->peers_open = false;
  s->use_loop = false;
  s->use_omp_loop = false;
  s->filter16 = false;
  s->d_ctl = nullptr;
  s->d_cta_cands = nullptr;
  s->d_cta_lost = nullptr;
  s->d_claims = nullptr;
  s->claims_cap = 0;
  memset(&s->nw, 0, sizeof(NnlsWork));
  s->d_nw = nullptr;
  s->nw_cap = 0;
  s->trace_on = 0;
  s->d_trace = nullptr;
  s->trace_cap = s->trace_n = 0;
  s->events_cap = 0;
  s->profiling = 0;
  s->build_ms = s->scan_ms = 0.f;
  s->scan_launches = s->step_launches = s->loop_launches = 0;
  for (int i = 0; i < kMaxWorld; ++i) s->peer_ptrs[i] = nullptr;
  const int rc = solver_init(s, ctx, v, alg, b, bnorm, norm_sum, row_offset, n_global);
  if (rc != BCG_OK) { bcg_solver_destroy(s); return rc; }
  *out = s;
  return BCG_OK;
}

extern "C" int bcg_solver_destroy(bcg_solver* s) {
  if (!s) return BCG_OK;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  SolverState& h = s->h;
  // (the mailbox and the peer mappings belong to the context and outlive the solver)
  void* bufs[] = {h.b, h.bn, h.xw, h.xw_new, h.xf, h.dir64, h.dir32, h.wrow, h.cands, h.act_idx, h.act_w,
                  h.act_w_new, h.act_norm, h.act_tmp, h.act_rows, h.events, s->d_fout, s->d, s->d_ctl, s->d_cta_cands, s->d_trace, s->d_claims,
                  s->d_cta_lost, h.cand_lost, h.exact_cands};
  for (void* p : bufs)
    if (p) cudaFree(p);
  if (s->d_nw) cudaMemcpy(&s->nw, s->d_nw, sizeof(NnlsWork), cudaMemcpyDeviceToHost);   // R / R2 may have swapped on the device
  void* nb[] = {s->nw.Q, s->nw.R, s->nw.R2, s->nw.rot, s->nw.rem, s->nw.c, s->nw.z, s->nw.wP, s->nw.h, s->nw.v, s->nw.P, s->nw.Z, s->nw.inP, s->d_nw};
  for (void* p : nb)
    if (p) cudaFree(p);
  for (cudaEvent_t e : s->scan_ev) cudaEventDestroy(e);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  delete s;
  return BCG_OK;
}

static int ensure_ctx_mailbox(bcg_ctx* ctx) {
  const int64_t bytes = 2 * (int64_t)kMaxWorld * mail_slot_bytes_for(1024);
  if (!ctx->mail) {
    CK(cudaMalloc(&ctx->mail, bytes));
    CK(cudaMemset(ctx->mail, 0, bytes));
    ctx->mail_bytes = bytes;
  }
  return BCG_OK;
}

// the mailbox belongs to the context: its handle can be exported before any solver exists, so the host can fold the
// handle exchange into the same all-gather as the row counts
extern "C" int bcg_ctx_comm_handle(bcg_ctx* ctx, void* handle64) {
  if (!ctx || !handle64) return fail(BCG_ERR_ARG, "null argument");
  RET(use_device(ctx));
  RET(ensure_ctx_mailbox(ctx));
  cudaIpcMemHandle_t hnd;
  CK(cudaIpcGetMemHandle(&hnd, ctx->mail));
  memcpy(handle64, &hnd, 64);
  return BCG_OK;
}

static int ensure_mailbox(bcg_solver* s, int world) {
  (void)world;
  bcg_ctx* ctx = s->ctx;
  const int64_t bytes = 2 * (int64_t)kMaxWorld * mail_slot_bytes_for(1024);
  if (mail_slot_bytes_for(s->h.ld) > mail_slot_bytes_for(1024)) return fail(BCG_ERR_UNSUPPORTED, "row too long for the mailbox");
  if (!ctx->mail) {
    CK(cudaMalloc(&ctx->mail, bytes));
    CK(cudaMemset(ctx->mail, 0, bytes));
    ctx->mail_bytes = bytes;
  }
  return BCG_OK;
}

extern "C" int bcg_solver_comm_handle(bcg_solver* s, void* handle64) {
  if (!s || !handle64) return fail(BCG_ERR_ARG, "null argument");
  RET(use_device(s->ctx));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle must be 64 bytes");
  RET(ensure_mailbox(s, kMaxWorld));
  cudaIpcMemHandle_t hnd;
  CK(cudaIpcGetMemHandle(&hnd, s->ctx->mail));
  memcpy(handle64, &hnd, 64);
  return BCG_OK;
}

extern "C" int bcg_solver_comm_connect(bcg_solver* s, int32_t world, int32_t rank, const void* handles64) {
  if (!s || !handles64) return fail(BCG_ERR_ARG, "null argument");
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
    return fail(BCG_ERR_ARG, "bad world/rank %d/%d (max world %d)", world, rank, kMaxWorld);
  RET(use_device(s->ctx));
  RET(ensure_mailbox(s, kMaxWorld));
  SolverState& h = s->h;
  bcg_ctx* ctx = s->ctx;
  for (int p = 0; p < world; ++p) {
    if (p == rank) {
      s->peer_ptrs[p] = ctx->mail;
    } else {
      std::array<unsigned char, 64> key;
      memcpy(key.data(), (const char*)handles64 + 64 * p, 64);
      void* ptr = nullptr;
      for (auto& pc : ctx->peer_cache)
        if (pc.first == key) { ptr = pc.second; break; }
      if (!ptr) {
        cudaIpcMemHandle_t hnd;
        memcpy(&hnd, key.data(), 64);
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
          return fail(BCG_ERR_COMM, "cudaIpcOpenMemHandle(rank %d) failed: %s", p, cudaGetErrorString(e));
        ctx->peer_cache.emplace_back(key, ptr);
      }
      s->peer_ptrs[p] = ptr;
    }
    h.mail_peer[p] = (unsigned char*)s->peer_ptrs[p];
  }
  s->peers_open = true;
  h.world = world;
  h.rank = rank;
  h.mail_local = ctx->mail;
  h.mail_slot_bytes = mail_slot_bytes_for(h.ld);
  h.seq = (++ctx->mail_epoch) << 32;
  RET(push_state(s));
  CK(cudaStreamSynchronize(s->ctx->stream));
  return BCG_OK;
}

// NNLS work space for the current active-set capacity.  The factorisation buffers (Q, T = R^-1 and its double) are sized
// by min(cap, S + 1) -- the passive set holds independent columns, nP <= S -- so once the capacity exceeds S they keep
// their shape: growing the active set then only extends the small per-slot arrays and the warm start SURVIVES the growth.
static int ensure_nnls(bcg_solver* s) {
  const int cap = s->h.cap, S = s->h.S;
  if (s->d_nw && s->nw_cap == cap) return BCG_OK;
  NnlsWork& w = s->nw;
  cudaStream_t st = s->ctx->stream;
  if (s->d_nw) CK(cudaMemcpy(&w, s->d_nw, sizeof(NnlsWork), cudaMemcpyDeviceToHost));      // R / R2 may have swapped
  const int old_cap = s->d_nw ? s->nw_cap : 0;
  const int tld = std::min(cap, S + 1);
  const bool keep = s->d_nw && w.Q && w.tld == tld;
  if (!keep) {
    void* old[] = {w.Q, w.R, w.R2};
    for (void* p : old)
      if (p) CK(cudaFree(p));
    w.Q = w.R = w.R2 = nullptr;
    CK(cudaMalloc(&w.Q, (size_t)tld * S * sizeof(double)));
    CK(cudaMalloc(&w.R, (size_t)tld * tld * sizeof(double)));
    CK(cudaMalloc(&w.R2, (size_t)tld * tld * sizeof(double)));
    w.valid = 0;
    w.nP = w.nZ = 0;
  }
  // per-slot / per-position arrays: extended with their contents (zero beyond the old capacity)
  RET(grow(&w.rot, (size_t)3 * old_cap, (size_t)3 * cap, st));
  RET(grow(&w.rem, (size_t)old_cap, (size_t)cap, st));
  RET(grow(&w.c, (size_t)old_cap, (size_t)cap, st));
  RET(grow(&w.z, (size_t)2 * old_cap, (size_t)2 * cap, st));
  RET(grow(&w.wP, (size_t)old_cap, (size_t)cap, st));
  RET(grow(&w.h, (size_t)old_cap, (size_t)cap, st));
  RET(grow(&w.P, (size_t)old_cap, (size_t)cap, st));
  RET(grow(&w.Z, (size_t)old_cap, (size_t)cap, st));
  RET(grow(&w.inP, (size_t)old_cap, (size_t)cap, st));
  if (!w.v) CK(cudaMalloc(&w.v, (size_t)S * sizeof(double)));
  w.cap = cap;
  w.tld = tld;
  w.downdate = env_int("BCG_NNLS_DOWNDATE", 1);
  if (!s->d_nw) CK(cudaMalloc(&s->d_nw, sizeof(NnlsWork)));
  CK(cudaMemcpyAsync(s->d_nw, &w, sizeof(NnlsWork), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  s->nw_cap = cap;
  return BCG_OK;
}

static int invalidate_nnls(bcg_solver* s) {
  if (!s->d_nw) return BCG_OK;
  const int32_t zero = 0;
  CK(cudaMemcpyAsync(&s->d_nw->valid, &zero, sizeof(int32_t), cudaMemcpyHostToDevice, s->ctx->stream));
  return BCG_OK;
}

// NNLS re-solve over the stored rows with positive weight, on the device (snnls.py:82-97 with from_scratch = 1;
// orthopursuit.py:39-41 warm-started with from_scratch = 0).  Refreshes A w and error().
extern "C" int bcg_solver_nnls(bcg_solver* s, int32_t from_scratch) {
  if (!s) return fail(BCG_ERR_ARG, "null solver");
  RET(use_device(s->ctx));
  RET(ensure_nnls(s));
  RET(push_state(s));
  nnls_kernel<<<1, kStepThreads, 0, s->ctx->stream>>>(s->d, s->d_nw, from_scratch ? 1 : 0, env_int("BCG_OMP_WIDE", 1));
  CK(cudaGetLastError());
  RET(pull_state(s));
  return BCG_OK;
}

static int run_persistent(bcg_solver* s, int32_t itrs, bool omp) {
  cudaStream_t st = s->ctx->stream;
  SolverState& h = s->h;
// The whole build call as ONE persistent cooperative kernel (loop_kernel.cuh / omp_loop_kernel.cuh).  Rare exception: when the control
  // warp finds the float32 candidate set ambiguous (an unpublished score inside the near-tie window) it stops before
  // that iteration; the selection is redone exactly in float64 (exact_scan_kernel) and the loop is relaunched for the
  // remaining iterations with that result.
  LoopArgs la;
  la.st = s->d;
  la.ctl = s->d_ctl;
  la.cta_cands = s->d_cta_cands;
  la.cta_lost = s->d_cta_lost;
  la.g.An = s->v->An;
  la.g.n_rows = s->v->n;
  la.g.ld = s->v->ld;
  la.g.rps = s->sc.rps;
  la.g.stages = s->sc.stages;
  la.g.evict_first = s->sc.evict_first;
  const bool f16 = s->filter16 && s->v->An16 != nullptr;
  la.g.An16 = f16 ? s->v->An16 : nullptr;
  la.g.ld16 = s->v->ld16;
  la.g.rps16 = s->sc.rps16;
  la.g.eps16 = filter_eps_unit(s->v->S);
  la.wpb = s->sc.wpb;
  if (s->claims_cap < itrs) {
    if (s->d_claims) CK(cudaFree(s->d_claims));
    s->d_claims = nullptr;
    CK(cudaMalloc(&s->d_claims, 2 * (size_t)itrs * sizeof(unsigned int)));   // claim counters | filter bounds
    s->claims_cap = itrs;
  }
  la.claims = s->d_claims;
  la.filt_L = reinterpret_cast<int*>(s->d_claims + s->claims_cap);
  la.static_frac = (float)env_int("BCG_STATIC_PCT", 100) / 100.f;   // measured: a larger dynamic share only costs (atomics); the grid is HBM-bound either way
  la.trace = nullptr;
  s->trace_n = 0;
  if (s->trace_on) {
    if (s->trace_cap < itrs) {
      if (s->d_trace) CK(cudaFree(s->d_trace));
      s->d_trace = nullptr;
      CK(cudaMalloc(&s->d_trace, (size_t)itrs * 8 * sizeof(unsigned long long)));
      s->trace_cap = itrs;
    }
    CK(cudaMemsetAsync(s->d_trace, 0, (size_t)itrs * 8 * sizeof(unsigned long long), st));
    s->trace_n = itrs;
  }
  CK(cudaEventRecord(s->ev0, st));
  int done = 0;
  s->loop_launches = 0;
  s->scan_launches = 0;
  la.cont = 0;
  la.use_pre = 0;
  for (;;) {
    la.itrs = itrs - done;
    la.trace = s->trace_on ? s->d_trace + (size_t)done * 8 : nullptr;
    CK(cudaMemsetAsync(s->d_claims, 0, (size_t)la.itrs * sizeof(unsigned int), st));
    CK(cudaMemsetAsync(la.filt_L, 0x80, (size_t)la.itrs * sizeof(int), st));
    CK(cudaMemsetAsync(s->d_ctl, 0, sizeof(LoopCtl), st));
    if (omp) CK(omp_loop_launch(s->sc, la, s->d_nw, env_int("BCG_OMP_WIDE", 1), st));
    else CK(loop_launch(s->sc, la, st));
    CK(cudaEventRecord(s->ev1, st));                                 // (re-recorded per launch: the last one counts)
    s->loop_launches += 1;
    RET(pull_state(s));
    if (h.filt_overflow) {                                           // more near-maximal row groups than re-scan slots (e.g. many
      s->filter16 = false;                                           // duplicated rows): this solver goes back to the float32 stream
      la.g.An16 = nullptr;
    }
    if (!h.need_exact || h.halted || h.comm_error) break;
    done += h.iters_done;
    if (done >= itrs) break;                                         // (cannot happen: the stop precedes an iteration)
    step_kernel<<<1, kStepThreads, 0, st>>>(s->d, 0, 1, 0);          // float64 direction of the pending iteration
    CK(exact_scan_launch(h.n_exact_cands, s->d, 1, st));
    CK(cudaGetLastError());
    s->scan_launches += 1;
    la.cont = 1;
    la.use_pre = 1;
  }
  CK(cudaEventElapsedTime(&s->build_ms, s->ev0, s->ev1));
  s->step_launches = s->scan_launches;
  s->scan_ms = 0.f;
  return BCG_OK;
}

extern "C" int bcg_solver_build(bcg_solver* s, int32_t itrs, double tol, bcg_iter_event* events, int32_t* n_events) {
  if (!s) return fail(BCG_ERR_ARG, "null solver");
  RET(use_device(s->ctx));
  if (n_events) *n_events = 0;
  if (itrs <= 0 || s->h.halted || (s->v->n == 0 && s->h.world == 1)) return BCG_OK;
  cudaStream_t st = s->ctx->stream;
  SolverState& h = s->h;
  RET(ensure_capacity(s, itrs + 1));
  if (s->events_cap < itrs) {
    if (h.events) CK(cudaFree(h.events));
    h.events = nullptr;
    s->events_cap = 0;
    const int want = std::max(itrs, 256);
    CK(cudaMalloc(&h.events, (size_t)want * sizeof(bcg_iter_event)));
    s->events_cap = want;
  }
  CK(cudaMemsetAsync(h.events, 0, (size_t)itrs * sizeof(bcg_iter_event), st));
  h.n_events = 0;
  h.tol = tol;
  h.comm_error = 0;
  RET(push_state(s));
  const bool loop = s->use_loop && !s->profiling;
  if (h.alg == BCG_ALG_OMP && s->use_omp_loop && !s->profiling && !env_int("BCG_OMP_TRACE", 0)) {
    // OrthoPursuit as one persistent kernel: CTA 0 runs the selection / NNLS logic, the other CTAs scan (omp_loop_kernel.cuh)
    RET(ensure_nnls(s));
    h.n_cands = 2 * (s->sc.grid - 1);
    RET(push_state(s));
    RET(run_persistent(s, itrs, true));
  } else if (h.alg == BCG_ALG_OMP) {
    // OrthoPursuit, launch per iteration (profiling, tracing, S > 512, never-materialising solver): selection scan +
    // on-device NNLS, no host round trip inside the loop
    RET(ensure_nnls(s));
    h.n_cands = s->sc.grid * s->sc.wpb;
    RET(push_state(s));
    const int wide = env_int("BCG_OMP_WIDE", 1);
    DevBuf<unsigned long long> d_omp_trace;
    struct TraceGuard {                       // the state must not keep a pointer to the trace buffer once it is freed
      bcg_solver* s;
      ~TraceGuard() {
        if (s->h.omp_trace) {
          s->h.omp_trace = nullptr;
          cudaMemcpy(&s->d->omp_trace, &s->h.omp_trace, sizeof(void*), cudaMemcpyHostToDevice);
        }
      }
    } trace_guard{s};
    if (env_int("BCG_OMP_TRACE", 0)) {
      CK(d_omp_trace.alloc((size_t)itrs * 16));
      CK(cudaMemsetAsync(d_omp_trace, 0, (size_t)itrs * 16 * sizeof(unsigned long long), st));
      h.omp_trace = d_omp_trace;
      RET(push_state(s));
    }
    if (s->profiling) {
      while ((int)s->scan_ev.size() < 2 * itrs) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        s->scan_ev.push_back(e);
      }
    }
    CK(cudaEventRecord(s->ev0, st));
    step_kernel<<<1, kStepThreads, 0, st>>>(s->d, 0, 1, 1);           // reset the per-call retry flag; first residual direction
    for (int i = 0; i < itrs; ++i) {
      if (s->profiling) RET(launch_scan(s, s->scan_ev[2 * i], s->scan_ev[2 * i + 1]));
      else RET(launch_scan(s));
      omp_iteration_kernel<<<1, kStepThreads, 0, st>>>(s->d, s->d_nw, wide, (i + 1 < itrs) ? 1 : 0);
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(s->ev1, st));
    RET(pull_state(s));
    CK(cudaEventElapsedTime(&s->build_ms, s->ev0, s->ev1));
    if (d_omp_trace.p) {
      // diagnostics: mean time between the phase marks of omp_iteration (see omp_mark), first / second half of the call
      std::vector<unsigned long long> tr((size_t)itrs * 16);
      CK(cudaMemcpy(tr.data(), d_omp_trace, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      static const char* names[12] = {"start", "count+save", "pick_local", "neg-dir", "select-end", "nnls-setup", "resid+dual",
                                      "qr-append", "solve", "resid2", "writeback+err", "event+prep"};
      for (int half = 0; half < 2; ++half) {
        double acc[12] = {0}, sub[4] = {0}; int cnt = 0;
        for (int i = half * itrs / 2; i < (half + 1) * itrs / 2; ++i) {
          const unsigned long long* t = &tr[(size_t)i * 16];
          bool ok = true;
          for (int k = 0; k < 12; ++k) ok = ok && t[k] != 0;
          if (!ok || t[11] - t[0] > 250000) continue;         // skip incomplete (failed / step-back) iterations
          for (int k = 1; k < 12; ++k) acc[k] += (double)(t[k] - t[k - 1]);
          sub[0] += (double)(t[12] - t[6]); sub[1] += (double)(t[13] - t[12]); sub[2] += (double)(t[14] - t[13]);
          sub[3] += (double)(t[7] - t[14]);
          ++cnt;
        }
        fprintf(stderr, "[bcg] omp trace, iterations %d..%d (%d clean):", half * itrs / 2, (half + 1) * itrs / 2, cnt);
        for (int k = 1; k < 12; ++k) fprintf(stderr, " %s %.1f", names[k], cnt ? acc[k] / cnt / 1e3 : 0.);
        fprintf(stderr, " us | append: pass0 %.1f pass1 %.1f sums+store %.1f Tcol %.1f us\n", cnt ? sub[0] / cnt / 1e3 : 0.,
                cnt ? sub[1] / cnt / 1e3 : 0., cnt ? sub[2] / cnt / 1e3 : 0., cnt ? sub[3] / cnt / 1e3 : 0.);
      }
    }
    s->scan_launches = itrs;
    s->step_launches = itrs + 1;
    s->loop_launches = 0;
    s->scan_ms = 0.f;
    if (s->profiling) {
      for (int i = 0; i < itrs; ++i) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, s->scan_ev[2 * i], s->scan_ev[2 * i + 1]));
        s->scan_ms += ms;
      }
    }
  } else if (loop) {
    RET(run_persistent(s, itrs, false));
  } else {
    if (s->profiling) {
      while ((int)s->scan_ev.size() < 2 * itrs) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        s->scan_ev.push_back(e);
      }
    }
    CK(cudaEventRecord(s->ev0, st));
    step_kernel<<<1, kStepThreads, 0, st>>>(s->d, 0, 1, 1);
    for (int i = 0; i < itrs; ++i) {
      if (s->profiling) RET(launch_scan(s, s->scan_ev[2 * i], s->scan_ev[2 * i + 1]));
      else RET(launch_scan(s));
      step_kernel<<<1, kStepThreads, 0, st>>>(s->d, 1, (i + 1 < itrs) ? 1 : 0, 0);
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(s->ev1, st));
    RET(pull_state(s));
    CK(cudaEventElapsedTime(&s->build_ms, s->ev0, s->ev1));
    s->scan_launches = itrs;
    s->step_launches = itrs + 1;
    s->loop_launches = 0;
    s->scan_ms = 0.f;
    if (s->profiling) {
      for (int i = 0; i < itrs; ++i) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, s->scan_ev[2 * i], s->scan_ev[2 * i + 1]));
        s->scan_ms += ms;
      }
    }
  }
  const int ne = h.n_events;
  if (events && ne > 0) CK(cudaMemcpy(events, h.events, (size_t)ne * sizeof(bcg_iter_event), cudaMemcpyDeviceToHost));
  if (n_events) *n_events = ne;
  if (h.comm_error == 2) return fail(BCG_ERR_STATE, "no comparable score in the scan (non-finite matrix entries?)");
  if (h.comm_error) return fail(BCG_ERR_COMM, "peer-memory candidate exchange timed out (a rank is missing)");
  return BCG_OK;
}

extern "C" int bcg_solver_omp_select(bcg_solver* s, int64_t* f) {
  if (!s || !f) return fail(BCG_ERR_ARG, "null argument");
  RET(use_device(s->ctx));
  if (s->h.alg != BCG_ALG_OMP) return fail(BCG_ERR_STATE, "solver was not created with BCG_ALG_OMP");
  cudaStream_t st = s->ctx->stream;
  RET(ensure_capacity(s, 1));
  s->h.comm_error = 0;
  RET(push_state(s));
  step_kernel<<<1, kStepThreads, 0, st>>>(s->d, 0, 1, 0);
  RET(launch_scan(s));
  omp_select_kernel<<<1, kStepThreads, 0, st>>>(s->d, s->d_fout);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(f, s->d_fout, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  RET(pull_state(s));
  if (s->h.comm_error) return fail(BCG_ERR_COMM, "peer-memory candidate exchange timed out (a rank is missing)");
  return BCG_OK;
}

// argmax over the local rows of <a_n / ||a_n||, dir> (sparsevi.py:51,56-57: corrs = vecs.dot(resid)/norms);
// no solver state is changed.  *score = the float64 inner product of the winning unit row with dir.
extern "C" int bcg_solver_probe_argmax(bcg_solver* s, const double* dir, int64_t* f, double* score) {
  if (!s || !dir || !f || !score) return fail(BCG_ERR_ARG, "null argument");
  RET(use_device(s->ctx));
  if (s->h.alg == BCG_ALG_GIGA) return fail(BCG_ERR_STATE, "probe needs a one-direction solver (FW / OMP)");
  if (s->v->n == 0) { *f = -1; *score = 0.; return BCG_OK; }
  const int S = s->h.S, ld = s->h.ld;
  double nrm = 0.;
  for (int i = 0; i < S; ++i) nrm += dir[i] * dir[i];
  nrm = sqrt(nrm);
  const double inv = nrm > 0. ? 1. / nrm : 1.;
  std::vector<double> d64(S);
  std::vector<float> d32(ld, 0.f);
  for (int i = 0; i < S; ++i) { d64[i] = dir[i] * inv; d32[i] = (float)d64[i]; }
  cudaStream_t st = s->ctx->stream;
  const int keep_sel = s->h.select_failed, keep_halt = s->h.halted;
  s->h.select_failed = 0;
  s->h.halted = 0;
  RET(push_state(s));
  CK(cudaMemcpyAsync(s->h.dir64, d64.data(), S * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(s->h.dir32, d32.data(), ld * sizeof(float), cudaMemcpyHostToDevice, st));
  RET(launch_scan(s));
  double* d_score = reinterpret_cast<double*>(s->d_fout + 1);
  probe_kernel<<<1, kStepThreads, 0, st>>>(s->d, s->d_fout, d_score);
  CK(cudaGetLastError());
  int64_t hf = -1;
  double hs = 0.;
  CK(cudaMemcpyAsync(&hf, s->d_fout, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&hs, d_score, sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  s->h.select_failed = keep_sel;
  s->h.halted = keep_halt;
  RET(push_state(s));
  CK(cudaStreamSynchronize(st));
  *f = hf;
  *score = hs * nrm;
  return BCG_OK;
}

extern "C" int bcg_solver_error(bcg_solver* s, double* err) {
  if (!s || !err) return fail(BCG_ERR_ARG, "null argument");
  *err = s->h.err;
  return BCG_OK;
}

extern "C" int bcg_solver_halted(bcg_solver* s, int32_t* reached) {
  if (!s || !reached) return fail(BCG_ERR_ARG, "null argument");
  *reached = s->h.halted;
  return BCG_OK;
}

extern "C" int bcg_solver_active(bcg_solver* s, int64_t cap, int64_t* idx, double* w, int64_t* k) {
  if (!s || !k) return fail(BCG_ERR_ARG, "null argument");
  RET(use_device(s->ctx));
  const int64_t n = s->h.nact;
  *k = n;
  if (n == 0 || (!idx && !w)) return BCG_OK;
  if (cap < n) return fail(BCG_ERR_ARG, "capacity %lld < %lld stored rows", (long long)cap, (long long)n);
  cudaStream_t st = s->ctx->stream;
  if (idx) CK(cudaMemcpyAsync(idx, s->h.act_idx, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  if (w) CK(cudaMemcpyAsync(w, s->h.act_w, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BCG_OK;
}

extern "C" int bcg_solver_size(bcg_solver* s, int64_t* n_positive, int64_t* n_stored) {
  if (!s) return fail(BCG_ERR_ARG, "null solver");
  const int64_t n = s->h.nact;
  if (n_stored) *n_stored = n;
  if (n_positive) {
    std::vector<double> w((size_t)std::max<int64_t>(n, 1));
    int64_t k = 0;
    RET(bcg_solver_active(s, n, nullptr, w.data(), &k));
    int64_t c = 0;
    for (int64_t i = 0; i < n; ++i) c += w[i] > 0.;
    *n_positive = c;
  }
  return BCG_OK;
}

extern "C" int bcg_solver_active_rows(bcg_solver* s, int64_t first, int64_t count, double* out) {
  if (!s || !out) return fail(BCG_ERR_ARG, "null argument");
  if (first < 0 || count < 0 || first + count > s->h.nact) return fail(BCG_ERR_ARG, "active row range out of bounds");
  RET(use_device(s->ctx));
  if (count == 0) return BCG_OK;
  cudaStream_t st = s->ctx->stream;
  const int64_t tot = count * s->h.S;
  DevBuf<double> tmp;
  CK(tmp.alloc((size_t)tot));
  expand_active_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(s->h.act_rows, s->h.act_norm, first, count, s->h.S,
                                                                   s->h.ld, tmp);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, tmp, (size_t)tot * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BCG_OK;
}

extern "C" int bcg_solver_set_weights(bcg_solver* s, const double* w, int64_t k) {
  if (!s || (k > 0 && !w)) return fail(BCG_ERR_ARG, "null argument");
  if (k != s->h.nact) return fail(BCG_ERR_ARG, "expected %d weights, got %lld", s->h.nact, (long long)k);
  RET(use_device(s->ctx));
  cudaStream_t st = s->ctx->stream;
  if (k > 0) CK(cudaMemcpyAsync(s->h.act_w, w, (size_t)k * sizeof(double), cudaMemcpyHostToDevice, st));
  s->h.kkt_valid = 0;
  RET(push_state(s));
  RET(invalidate_nnls(s));
  refresh_kernel<<<1, kStepThreads, 0, st>>>(s->d);
  CK(cudaGetLastError());
  RET(pull_state(s));
  return BCG_OK;
}

// Replace the stored active set: rows `idx` (GLOBAL indices owned by this rank) with weights `w`; their unit rows and
// norms are gathered on the device, A w and the error are recomputed.  This is the sparse form of assigning the dense
// weight vector `self.w = ...` (what the sampling solvers of snnls/sampling.py:33-35 do after every draw).
__global__ void gather_active_kernel(SolverState* st, int k) {
  const int ld = st->ld;
  for (int r = blockIdx.x; r < k; r += gridDim.x) {
    const int64_t lrow = st->act_idx[r] - st->row_offset;
    const float* src = st->An + (size_t)lrow * ld;
    for (int c = threadIdx.x; c < ld; c += blockDim.x) st->act_rows[(size_t)r * ld + c] = src[c];
    if (threadIdx.x == 0) st->act_norm[r] = st->norms[lrow];
  }
}

extern "C" int bcg_solver_set_active(bcg_solver* s, const int64_t* idx, const double* w, int64_t k) {
  if (!s || k < 0 || (k > 0 && (!idx || !w))) return fail(BCG_ERR_ARG, "bad arguments");
  RET(use_device(s->ctx));
  SolverState& h = s->h;
  for (int64_t i = 0; i < k; ++i)
    if (idx[i] < h.row_offset || idx[i] >= h.row_offset + h.n_local)
      return fail(BCG_ERR_ARG, "row %lld is not owned by this rank", (long long)idx[i]);
  cudaStream_t st = s->ctx->stream;
  if (k > h.cap) { h.nact = 0; RET(ensure_capacity(s, (int)k)); }
  if (k > 0) {
    CK(cudaMemcpyAsync(h.act_idx, idx, (size_t)k * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h.act_w, w, (size_t)k * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  h.nact = (int32_t)k;
  h.kkt_valid = 0;
  RET(push_state(s));
  if (k > 0) {
    gather_active_kernel<<<(unsigned)std::min<int64_t>(k, 1024), 128, 0, st>>>(s->d, (int)k);
    CK(cudaGetLastError());
  }
  RET(invalidate_nnls(s));
  refresh_kernel<<<1, kStepThreads, 0, st>>>(s->d);
  CK(cudaGetLastError());
  RET(pull_state(s));
  return BCG_OK;
}

extern "C" int bcg_solver_reset(bcg_solver* s) {
  if (!s) return fail(BCG_ERR_ARG, "null solver");
  RET(use_device(s->ctx));
  SolverState& h = s->h;
  h.nact = 0;
  h.halted = 0;
  h.retried = 0;
  h.select_failed = 0;
  h.kkt_valid = 0;
  h.scan_cnt = 0;
  h.err = h.bnorm;
  RET(invalidate_nnls(s));
  CK(cudaMemsetAsync(h.xw, 0, h.S * sizeof(double), s->ctx->stream));
  RET(push_state(s));
  CK(cudaStreamSynchronize(s->ctx->stream));
  return BCG_OK;
}

extern "C" int bcg_solver_timing(bcg_solver* s, float* build_ms, float* scan_ms, int32_t* scan_launches,
                                 int32_t* step_launches) {
  if (!s) return fail(BCG_ERR_ARG, "null solver");
  if (build_ms) *build_ms = s->build_ms;
  if (scan_ms) *scan_ms = s->scan_ms;
  if (scan_launches) *scan_launches = s->scan_launches;
  if (step_launches) *step_launches = s->step_launches + s->loop_launches;
  return BCG_OK;
}

extern "C" int bcg_solver_set_check_monotone(bcg_solver* s, int32_t check) {
  if (!s) return fail(BCG_ERR_ARG, "null solver");
  RET(use_device(s->ctx));
  s->h.check_monotone = check ? 1 : 0;
  RET(push_state(s));
  CK(cudaStreamSynchronize(s->ctx->stream));
  return BCG_OK;
}

extern "C" int bcg_solver_set_force_exact(bcg_solver* s, int32_t on) {
  if (!s) return fail(BCG_ERR_ARG, "null solver");
  RET(use_device(s->ctx));
  s->h.force_exact = on ? 1 : 0;
  RET(push_state(s));
  CK(cudaStreamSynchronize(s->ctx->stream));
  return BCG_OK;
}

extern "C" int bcg_solver_exact_count(bcg_solver* s, int64_t* n_exact) {
  if (!s || !n_exact) return fail(BCG_ERR_ARG, "null argument");
  *n_exact = s->h.n_exact;
  return BCG_OK;
}

// float16 pre-filter of the persistent kernels: switch it off / on for this solver (on only where it is available)
extern "C" int bcg_solver_set_filter16(bcg_solver* s, int32_t on) {
  if (!s) return fail(BCG_ERR_ARG, "null solver");
  RET(use_device(s->ctx));
  if (!on) { s->filter16 = false; return BCG_OK; }
  if ((s->use_loop || s->use_omp_loop) && s->sc.ch16 > 0) {
    RET(vecs_ensure_half(s->v));
    s->filter16 = s->v->An16 != nullptr;
    s->h.filt_overflow = 0;
  }
  return BCG_OK;
}

// enabled: the next persistent launch streams the float16 copy; rows_rescanned: rows re-scanned in float32 so far
extern "C" int bcg_solver_filter16_stats(bcg_solver* s, int32_t* enabled, int64_t* rows_rescanned) {
  if (!s) return fail(BCG_ERR_ARG, "null solver");
  if (enabled) *enabled = (s->filter16 && s->v->An16) ? 1 : 0;
  if (rows_rescanned) *rows_rescanned = (int64_t)s->h.filt_rows;
  return BCG_OK;
}

extern "C" int bcg_solver_set_profiling(bcg_solver* s, int32_t per_kernel_events) {
  if (!s) return fail(BCG_ERR_ARG, "null solver");
  s->profiling = per_kernel_events ? 1 : 0;
  return BCG_OK;
}

extern "C" int bcg_solver_set_trace(bcg_solver* s, int32_t enable) {
  if (!s) return fail(BCG_ERR_ARG, "null solver");
  s->trace_on = enable ? 1 : 0;
  return BCG_OK;
}

extern "C" int bcg_solver_get_trace(bcg_solver* s, int32_t cap_iters, uint64_t* out, int32_t* n_iters) {
  if (!s || !n_iters) return fail(BCG_ERR_ARG, "null argument");
  RET(use_device(s->ctx));
  *n_iters = s->trace_n;
  if (!out || s->trace_n == 0) return BCG_OK;
  if (cap_iters < s->trace_n) return fail(BCG_ERR_ARG, "trace capacity too small");
  CK(cudaMemcpy(out, s->d_trace, (size_t)s->trace_n * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return BCG_OK;
}
