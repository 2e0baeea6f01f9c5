/*
 * bcg.h -- C ABI of the B200-native coreset-construction engine (libbcg_b200.so).
 *
 * Drop-in boundary for ONE hot path of trevorcampbell/bayesian-coresets: the projection of N
 * datapoints into an N x S matrix of centred log-likelihood vectors and the greedy sparse-NNLS
 * loop (GIGA / Frank-Wolfe / OrthoPursuit) that runs over that matrix.  Plain pointers and
 * sizes only; every call returns an int status (0 = BCG_OK) and bcg_last_error() returns a
 * human-readable message for the last failure on the calling thread.  Host pointers are
 * borrowed for the duration of a call only.  Not re-entrant per handle.
 *
 * Reference interfaces replaced (paths relative to the reference repository root):
 *   bcg_vecs_from_host_f64     bayesiancoresets/snnls/giga.py:10-13 (column norms, An = A/norms)
 *                              + coreset/hilbert.py:24 (b = vecs.sum(axis=0))
 *   bcg_vecs_project_lr        bayesiancoresets/projector.py:19-21 with
 *                              examples/common/model_lr.py:25-32 as the log-likelihood
 *   bcg_vecs_project_gaussian  projector.py:19-21 with examples/common/model_gaussian.py:4-10
 *   bcg_vecs_project_poisson   projector.py:19-21 with examples/common/model_poiss.py:25-38
 *   bcg_vecs_colsum / _rows    hilbert.py:24 / the ndarray returned by Projector.project
 *   bcg_glm_joint              examples/common/model_lr.py:38-80 / model_poiss.py:44-93 (log_joint, grad, hess: Laplace sampler reductions)
 *   bcg_sampler_gaussian_post  examples/common/model_gaussian.py:23-30 + examples/gaussian/main.py:107-113 (sampler feeding Projector.update)
 *   bcg_pseudo_grad            projector.py:23-28 + model_lr.py:50-57 / model_poiss.py:58-67 / model_gaussian.py:12-15, bpsvi.py:53
 *   bcg_dataset_project_lazy   the same projection without storing it (norms, b only); rows re-evaluated per selection pass
 *   bcg_dataset_audit          projector.py:19-21 + giga.py:31-38 / frankwolfe.py:17 (independent float64 re-evaluation)
 *   bcg_solver_create          snnls/snnls.py:9-16 + giga.py:8-18 / frankwolfe.py:7-13 /
 *                              orthopursuit.py:9-15
 *   bcg_solver_build           snnls/snnls.py:31-79 driving giga.py:20-64 / frankwolfe.py:15-40
 *   bcg_solver_omp_select      orthopursuit.py:17-35 (+ the w[f] = 1 of :38)
 *   bcg_solver_nnls            orthopursuit.py:39-41, snnls/snnls.py:82-97 (scipy.optimize.nnls)
 *   bcg_solver_error           snnls/snnls.py:28-29
 *   bcg_solver_active / size   snnls/snnls.py:22-26 (sparse form of w)
 *   bcg_solver_set_weights     the write-back of snnls.py:88 / orthopursuit.py:41
 *   bcg_solver_reset           snnls/snnls.py:18-20
 *
 * Storage: the N x S matrix is kept as unit-norm float32 rows (row stride `ld` floats, a
 * multiple of 4, zero padded) plus one float64 norm per row; all S-vector and scalar state of
 * the solvers is float64.
 */
#ifndef BCG_H
#define BCG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BCG_ABI_VERSION 1

/* status codes */
#define BCG_OK             0
#define BCG_ERR_CUDA       1   /* a CUDA runtime call or kernel failed */
#define BCG_ERR_ARG        2   /* invalid argument */
#define BCG_ERR_NO_DEVICE  3   /* no usable CUDA device: there is NO CPU fallback */
#define BCG_ERR_ZERO_B     4   /* ||b|| == 0 (giga.py:16-17) */
#define BCG_ERR_STATE      5   /* call not valid in the current solver state */
#define BCG_ERR_COMM       6   /* peer-memory exchange failed or timed out */
#define BCG_ERR_UNSUPPORTED 7  /* shape outside the supported range (e.g. S > 1024) */

/* models (bcg_dataset_project) */
#define BCG_MODEL_LR       0   /* examples/common/model_lr.py:25-32        z_n = y_n x_n */
#define BCG_MODEL_GAUSSIAN 1   /* examples/common/model_gaussian.py:4-10   known covariance */
#define BCG_MODEL_POISSON  2   /* examples/common/model_poiss.py:25-38     z_n = [x_n, y_n] */

/* algorithms (bcg_solver_create) */
#define BCG_ALG_GIGA 0
#define BCG_ALG_FW   1
#define BCG_ALG_OMP  2

/* per-iteration event codes (bcg_iter_event.code) */
#define BCG_IT_OK            0
#define BCG_IT_FAIL_CDIR     1  /* giga.py:28-29     aux0 = cdirnrm */
#define BCG_IT_FAIL_GEODESIC 2  /* giga.py:50-51     aux0 = gA, aux1 = gB */
#define BCG_IT_FAIL_GAMMA    3  /* frankwolfe.py:33  aux0 = gammanum, aux1 = gammadenom */
#define BCG_IT_FAIL_MONOTONE 4  /* snnls.py:58-61    aux0 = new error, aux1 = previous error */

typedef struct bcg_iter_event {
  int32_t code;      /* BCG_IT_* */
  int32_t nact;      /* number of stored active rows after the iteration */
  int64_t f;         /* selected GLOBAL row index (-1 when selection itself failed) */
  double  error;     /* ||A w - b||_2 after the iteration (unchanged state on failure) */
  double  aux0;
  double  aux1;
} bcg_iter_event;

typedef struct bcg_ctx    bcg_ctx;     /* one CUDA device + stream */
typedef struct bcg_dataset bcg_dataset; /* device-resident N_local x d data (float64) */
typedef struct bcg_vecs   bcg_vecs;    /* device-resident N_local x S projection */
typedef struct bcg_solver bcg_solver;  /* greedy sparse-NNLS state over a bcg_vecs */

/* ---- library / context ------------------------------------------------------------------ */
int         bcg_abi_version(void);
const char* bcg_last_error(void);
int  bcg_device_count(int* count);
int  bcg_ctx_create(int device, bcg_ctx** out);
int  bcg_ctx_destroy(bcg_ctx* ctx);
int  bcg_ctx_info(bcg_ctx* ctx, char* name, int name_cap, int* sm_count, int* cc_major, int* cc_minor,
                  int64_t* total_mem_bytes);
int  bcg_ctx_synchronize(bcg_ctx* ctx);
/* release the context's one-slot cache of matrix buffers (a destroyed bcg_vecs leaves its buffers there for the next one) */
int  bcg_ctx_trim(bcg_ctx* ctx);
int  bcg_ctx_mem_info(bcg_ctx* ctx, int64_t* free_bytes, int64_t* total_bytes);
/* Page-locked host memory for the caller's input arrays.  Every host -> device upload of this library checks
 * whether its source is page-locked (allocated here, or registered by the caller with the CUDA runtime): such a
 * source is copied by DMA straight from the caller's array, a pageable one is first staged through the library's
 * pinned buffers with a multi-threaded memcpy. */
int  bcg_host_alloc(int64_t bytes, void** out);
int  bcg_host_free(void* p);
/* write `bytes` of device scratch (evicts L2 between timed iterations) */
int  bcg_ctx_flush_l2(bcg_ctx* ctx, int64_t bytes);

/* ---- projection matrix ------------------------------------------------------------------ */
/* rows: host, row-major, n x S float64 with row stride ld_host (elements) */
int  bcg_vecs_from_host_f64(bcg_ctx* ctx, const double* rows, int64_t n, int32_t S, int64_t ld_host,
                            bcg_vecs** out);
/* The three bcg_vecs_project_* calls take HOST arrays and pipeline the upload with the projection (chunks are
 * staged and copied while the previous chunk is being projected); the data are not kept on the device.
 * Z: host n x d float64 (z_n = y_n x_n); theta: host S x d float64 */
int  bcg_vecs_project_lr(bcg_ctx* ctx, const double* Z, int64_t n, int32_t d, const double* theta, int32_t S,
                         bcg_vecs** out);
/* x: host n x d; theta: host S x d; Siginv: host d x d (symmetric) */
int  bcg_vecs_project_gaussian(bcg_ctx* ctx, const double* x, int64_t n, int32_t d, const double* theta,
                               int32_t S, const double* Siginv, bcg_vecs** out);
/* Z: host n x (d+1) = [x, y]; theta: host S x d */
int  bcg_vecs_project_poisson(bcg_ctx* ctx, const double* Z, int64_t n, int32_t d, const double* theta,
                              int32_t S, bcg_vecs** out);
/* Device-resident dataset: upload Z (n x zld float64) once, project it many times (SparseVI / BatchPSVI
 * re-project with new samples at every optimisation step: sparsevi.py:23-42, bpsvi.py:24-40).
 * bcg_dataset_project evaluates model `model` for the S samples theta (host S x d) and row-centres it
 * (projector.py:19-21).  Optional outputs: out_vecs = the resident unit-row matrix; rows64 = host n x S
 * float64 centred rows (what Projector.project returns; for small n); colsum = host S column sums.  With
 * only colsum requested the N x S matrix is never written (the project(data).sum(axis=0) of
 * sparsevi.py:71-72 / bpsvi.py:49-51). */
int  bcg_dataset_create(bcg_ctx* ctx, const double* Z, int64_t n, int32_t zld, bcg_dataset** out);
int  bcg_dataset_destroy(bcg_dataset* ds);
/* rowidx (host, nsel entries) or NULL: project only the listed data rows, in that order, gathered on the
 * device -- the `data[sub_idcs]` of the subsampled tangent spaces (sparsevi.py:33-35, bpsvi.py:33-35,
 * hilbert.py:16-17) without re-uploading the rows */
int  bcg_dataset_project(bcg_dataset* ds, const int64_t* rowidx, int64_t nsel, int32_t model, int32_t d,
                         const double* theta, int32_t S, const double* Siginv, bcg_vecs** out_vecs,
                         double* rows64, double* colsum);
/* Independent float64 AUDIT of one selection pass (csrc/audit_kernel.cuh): every data row's centred log-likelihood
 * vector is recomputed from Z and theta in float64 with libdevice math (projector.py:19-21 + model_*.log_likelihood),
 * normalised (giga.py:10-13) and scored against the direction(s) of one greedy iteration -- kind = BCG_ALG_GIGA:
 * dirs = [cdir | xw] (2 x S), score = masked s0 / sqrt(1 - s1^2) (giga.py:31-38); otherwise dirs = [residual] (S),
 * score = <a_n / ||a_n||, residual> (frankwolfe.py:17, orthopursuit.py:19, sparsevi.py:51).  Shares no code and no
 * storage with the production scan.  All outputs are host arrays and optional: scores (n, needs dirs; -inf for zero
 * rows), norms (n), colsum (S; hilbert.py:24).  Also the first half of the never-materialising select. */
int  bcg_dataset_audit(bcg_dataset* ds, int32_t model, int32_t d, const double* theta, int32_t S, const double* Siginv,
                       int32_t kind, const double* dirs, double* scores, double* norms, double* colsum);
/* The NEVER-MATERIALISING projection: only the row norms, b = the column sums and the zero-row count are
 * computed; the result has no N x S matrix (bcg_vecs_rows_f64 fails on it).  A solver created over it re-evaluates every
 * row from the dataset in float64 at each selection pass (csrc/lazy_select_kernel.cuh): 8 N d_in bytes resident instead of
 * 4 N S, at ~10x the time per iteration.  `ds` must outlive the result and every solver created over it. */
int  bcg_dataset_project_lazy(bcg_dataset* ds, int32_t model, int32_t d, const double* theta, int32_t S,
                              const double* Siginv, bcg_vecs** out);
/* Datapoint gradients of the log-likelihood for the K pseudo-points of BatchPSVI, on the device (csrc/pseudo_grad_kernel.cuh):
 * glls (host K x S x dz, dz = d, or d + 1 for Poisson) = grad_z log-likelihood (model_lr.py:50-57, model_poiss.py:58-67 with the
 * documented repair, model_gaussian.py:12-15) centred over the last axis (projector.py:26), and / or its contraction
 * ugrad (host K x dz) = -(1/S) sum_s w_k resid_s glls[k,s,:] (bpsvi.py:53) without forming the (K, S, dz) array.
 * pts: host K x zld; theta: host S x d. */
int  bcg_pseudo_grad(bcg_ctx* ctx, int32_t model, const double* pts, int64_t K, int32_t zld, int32_t d, const double* theta,
                     int32_t S, const double* Siginv, const double* w, const double* resid, double* glls, double* ugrad);
/* Weighted Gaussian posterior SAMPLER on the device (csrc/sampler_kernels.cuh): theta = mup + E U^T with
 * (mup, U) = weighted_post(th0, Sig0inv, Siginv, pts, w) of examples/common/model_gaussian.py:23-30 -- the sampler_w of
 * examples/gaussian/main.py:107-113 that SparseVI / BatchPSVI call before every projection.  E: host S x d standard normals
 * drawn by the CALLER (the reference's np.random.randn call, so seeded runs consume the global RNG identically).
 * Outputs (host): theta S x d; optional mup (d) and U (d x d upper, Sigp = U U^T). */
int  bcg_sampler_gaussian_post(bcg_ctx* ctx, int32_t d, const double* th0, const double* Sig0inv, const double* Siginv,
                               const double* pts, const double* w, int64_t K, const double* E, int32_t S, double* theta,
                               double* mup, double* U);
/* Reductions of the LAPLACE sampler (examples/logistic_poisson_regression/main.py:16-41) over the K coreset points, on the
 * device: value = log_joint(Z, theta, w), grad = grad_th_log_joint, hess = hess_th_log_joint of examples/common/model_lr.py:
 * 25-80 (BCG_MODEL_LR) or model_poiss.py:32-93 (BCG_MODEL_POISSON), standard normal prior.  Z: host K x zld; outputs host,
 * any may be null. */
int  bcg_glm_joint(bcg_ctx* ctx, int32_t model, const double* Z, const double* w, int64_t K, int32_t zld, int32_t d,
                   const double* theta, double* value, double* grad, double* hess);
/* ll_ns = x_n . A_s + coff_s : the Gaussian model with A = theta Siginv (host S x d) and
 * coff_s = -0.5 theta_s Siginv theta_s precomputed by the caller */
int  bcg_dataset_project_linear(bcg_dataset* ds, const int64_t* rowidx, int64_t nsel, int32_t d, const double* A,
                                const double* coff, int32_t S, bcg_vecs** out_vecs, double* rows64, double* colsum);
/* host-array Gaussian projection with A = theta Siginv and coff precomputed (see bcg_dataset_project_linear) */
int  bcg_vecs_project_linear(bcg_ctx* ctx, const double* x, int64_t n, int32_t d, const double* A, const double* coff,
                             int32_t S, bcg_vecs** out);
int  bcg_vecs_shape(bcg_vecs* v, int64_t* n, int32_t* S, int32_t* ld);
int  bcg_vecs_colsum(bcg_vecs* v, double* out_S);          /* sum over local rows of the centred vectors */
int  bcg_vecs_norm_sum(bcg_vecs* v, double* out);          /* sum of local row norms */
int  bcg_vecs_zero_rows(bcg_vecs* v, int64_t* count);      /* rows whose norm is exactly 0 */
int  bcg_vecs_norms(bcg_vecs* v, int64_t row0, int64_t nrows, double* out);
int  bcg_vecs_rows_f64(bcg_vecs* v, int64_t row0, int64_t nrows, double* out); /* norm * unit row */
int  bcg_vecs_destroy(bcg_vecs* v);

/* ---- greedy solver ---------------------------------------------------------------------- */
/* b: host S float64 (GLOBAL target); norm_sum: GLOBAL sum of row norms (used by FW);
 * row_offset: global index of local row 0; n_global: total rows over all ranks */
int  bcg_solver_create(bcg_ctx* ctx, bcg_vecs* v, int32_t alg, const double* b, double norm_sum,
                       int64_t row_offset, int64_t n_global, bcg_solver** out);
int  bcg_solver_destroy(bcg_solver* s);
/* N-sharding over GPUs of one node: every rank exports a 64-byte handle of its mailbox, the
 * host exchanges them (any transport), then every rank connects to all `world` handles.  The mailbox and the
 * mapped peer mailboxes belong to the context and are reused by all its solvers (mapping a peer costs ~100 ms once).
 * Ranks must connect their solvers in the same order, and must be synchronised (any host barrier) between
 * bcg_solver_comm_connect and the first bcg_solver_build, and between build calls of DIFFERENT solvers. */
int  bcg_solver_comm_handle(bcg_solver* s, void* handle64);
int  bcg_ctx_comm_handle(bcg_ctx* ctx, void* handle64);       /* the same handle, available before a solver exists */
int  bcg_solver_comm_connect(bcg_solver* s, int32_t world, int32_t rank, const void* handles64);
/* run up to `itrs` greedy iterations entirely on the device (GIGA, FW).  events: host array of
 * `itrs` entries; n_events receives how many iterations were attempted. */
int  bcg_solver_build(bcg_solver* s, int32_t itrs, double tol, bcg_iter_event* events, int32_t* n_events);
/* OMP selection step: residual scan + negative direction over the active set; the selected row
 * joins the active set with weight 1 (orthopursuit.py:17-38).  *f = selected global index. */
int  bcg_solver_omp_select(bcg_solver* s, int64_t* f);
/* argmax over the local rows of <a_n/||a_n||, dir> and its float64 value; no state change.  Needs a
 * one-direction solver (BCG_ALG_FW / BCG_ALG_OMP).  Replaces sparsevi.py:51,56-57. */
int  bcg_solver_probe_argmax(bcg_solver* s, const double* dir, int64_t* f, double* score);
/* NNLS re-solve over the stored rows with positive weight, on the device, float64 Lawson-Hanson
 * (snnls.py:82-97 optimize() with from_scratch = 1; orthopursuit.py:39-41 with a warm start otherwise).
 * For BCG_ALG_OMP bcg_solver_build runs selection + this solve per iteration without host round trips. */
int  bcg_solver_nnls(bcg_solver* s, int32_t from_scratch);
int  bcg_solver_error(bcg_solver* s, double* err);
int  bcg_solver_size(bcg_solver* s, int64_t* n_positive, int64_t* n_stored);
int  bcg_solver_halted(bcg_solver* s, int32_t* reached_numeric_limit);
/* stored active set in selection order: global indices, weights (may contain zeros) */
int  bcg_solver_active(bcg_solver* s, int64_t cap, int64_t* idx, double* w, int64_t* k);
/* unnormalised float64 active rows, k x S, same order as bcg_solver_active */
int  bcg_solver_active_rows(bcg_solver* s, int64_t first, int64_t count, double* out);
/* overwrite the weights of the stored active rows (k must equal n_stored); recomputes A w and error */
int  bcg_solver_set_weights(bcg_solver* s, const double* w, int64_t k);
/* replace the stored active set by rows idx (global indices owned by this rank) with weights w: rows / norms are gathered
 * on the device, A w and error() recomputed -- the sparse form of `self.w = ...` (snnls/sampling.py:33-35) */
int  bcg_solver_set_active(bcg_solver* s, const int64_t* idx, const double* w, int64_t k);
int  bcg_solver_reset(bcg_solver* s);
/* snnls.py:9 `check_error_monotone` (default 1): with 0 the monotone-error test of snnls.py:56-61 is skipped -- and, as in
 * the reference, the retry flag is then never cleared by a successful step */
int  bcg_solver_set_check_monotone(bcg_solver* s, int32_t check);
/* Exactness of the selection.  The float32 scan publishes a bounded candidate set and every scan warp reports the best
 * score it did not publish; whenever such a score, or more than 8 published ones, lie inside the near-tie window of the
 * float32 maximum, the selection is redone by an exact float64 pass over all local rows (lowest index on ties).
 * exact_count: selections resolved that way so far.  set_force_exact(1): every selection takes the exact pass (tests). */
int  bcg_solver_exact_count(bcg_solver* s, int64_t* n_exact);
int  bcg_solver_set_force_exact(bcg_solver* s, int32_t on);
/* float16 pre-filter of the persistent greedy kernels (row lengths 129..512).  The loop streams a float16 copy of the unit
 * rows (2 N S bytes per iteration instead of 4 N S), bounds every row's float32 score from it with a rigorous error bound,
 * and re-scans in float32 only the row groups whose bound reaches the float32 maximum's near-tie window -- the selection
 * (giga.py:38, frankwolfe.py:17, orthopursuit.py:19) is bit-identical to the plain float32 scan.  On by default where the
 * copy fits in device memory (env BCG_FILTER16=0 disables it); filter16_stats: whether the next build uses it, and the
 * number of rows re-scanned in float32 so far. */
int  bcg_solver_set_filter16(bcg_solver* s, int32_t on);
int  bcg_solver_filter16_stats(bcg_solver* s, int32_t* enabled, int64_t* rows_rescanned);
/* device time of the last bcg_solver_build: total, and summed over the scan kernel launches */
int  bcg_solver_timing(bcg_solver* s, float* build_ms, float* scan_ms, int32_t* scan_launches,
                       int32_t* step_launches);
int  bcg_solver_set_profiling(bcg_solver* s, int32_t per_kernel_events);
/* diagnostics of the persistent loop kernel: 8 device timestamps (globaltimer ns) per iteration of the
 * last bcg_solver_build -- grid arrived, next direction published, scan started, scan finished (CTA 0),
 * then 4 control-warp milestones (candidates reduced, row fetched, line search done, committed) */
int  bcg_solver_set_trace(bcg_solver* s, int32_t enable);
int  bcg_solver_get_trace(bcg_solver* s, int32_t cap_iters, uint64_t* out, int32_t* n_iters);

/* ---- process group for N-sharding (torch-free) ------------------------------------------------ */
/* One process per GPU of one node.  Used for bootstrap (row counts, the 64-byte mailbox handles of
 * bcg_solver_comm_handle) and for the one-off / per-pass reductions of S-vectors (b = vecs.sum(axis=0) of hilbert.py:24
 * over all shards, sum ||a_n|| of frankwolfe.py:21, the column sums of sparsevi.py:48 / bpsvi.py:51); the per-iteration
 * candidate exchange of the greedy loop is fused into the kernels over NVLink peer memory and never comes here.
 * TCP on addr:port (rank 0 listens, the others connect; addr = "127.0.0.1" under torchrun), star topology.
 * All-reduces gather the contributions and reduce them in rank order on every rank: bit-identical results. */
typedef struct bcg_comm bcg_comm;
int  bcg_comm_create(const char* addr, int32_t port, int32_t rank, int32_t world, int32_t timeout_ms, bcg_comm** out);
int  bcg_comm_destroy(bcg_comm* c);
int  bcg_comm_rank(bcg_comm* c, int32_t* rank, int32_t* world);
/* recv: world * bytes, contribution of rank r at offset r * bytes; bytes = 0 is a barrier */
int  bcg_comm_allgather(bcg_comm* c, const void* send, int64_t bytes, void* recv);
int  bcg_comm_allreduce_f64(bcg_comm* c, double* data, int64_t n, int32_t op /* 0 sum, 1 max */);
int  bcg_comm_barrier(bcg_comm* c);

#ifdef __cplusplus
}
#endif
#endif /* BCG_H */
