/* b200als.h -- C ABI of libb200als.so, the Blackwell-native WRMF / ALS half-iteration engine.
 *
 * This is the drop-in boundary for rsparse's matrix-factorization hot path.  Every entry point
 * is `extern "C"`, takes plain pointers and sizes (no R, Rcpp, Armadillo or torch types) and
 * returns an integer status (0 = ok); `b200als_last_error()` gives the message.  The library's
 * only compute is hand-written sm_100a CUDA; there is no CPU fallback -- without a CUDA device
 * every compute call fails with B200ALS_ECUDA.
 *
 * Citations are relative to the reference tree (dselivanov/rsparse @ 54f7e6a):
 *   - `.Call` targets being replaced: src/RcppExports.cpp:329-416, registered :456-490,
 *     R stubs R/RcppExports.R:88-102, callers R/model_WRMF.R:493-495, :514
 *   - argument meaning:              src/wrmf_implicit.cpp:5-31, src/wrmf_explicit.cpp:5-27
 *   - sparse input layout:           inst/include/mapped_csc.hpp:8-29, src/utils.cpp:69-78
 *   - solver codes:                  inst/include/wrmf.hpp:16-18
 *
 * Dense layout: a factor matrix is `rank x n` column-major (R / Armadillo layout), i.e. n
 * contiguous rows of `rank` values.  Sparse layout: CSC whose COLUMNS are the rows being solved
 * for (R passes the CSC of users x items for the item half-iteration and the row-major copy
 * re-labelled as CSC for the user half-iteration, R/model_WRMF.R:184-189); 32-bit 0-based
 * indices exactly as in R's `@i` / `@p` slots; values either double (R's `@x`) or float.
 */
#ifndef B200ALS_H
#define B200ALS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ALS_VERSION 100

/* status codes */
#define B200ALS_OK 0
#define B200ALS_EINVAL 1       /* bad argument (null pointer, negative size, unknown code)       */
#define B200ALS_ECUDA 2        /* CUDA runtime error / no device                                 */
#define B200ALS_ENCCL 3        /* NCCL error                                                     */
#define B200ALS_ENOTSPD 4      /* a per-row system was not positive definite (Cholesky pivot<=0)  */
#define B200ALS_EUNSUPPORTED 5 /* valid in the reference but not implemented by this engine yet   */

/* inst/include/wrmf.hpp:16-18 */
#define B200ALS_CHOLESKY 0
#define B200ALS_CONJUGATE_GRADIENT 1
#define B200ALS_NNLS 2

#define B200ALS_IMPLICIT 0
#define B200ALS_EXPLICIT 1

/* which factor matrix a session call refers to */
#define B200ALS_ITEMS 0
#define B200ALS_USERS 1

/* Zero-copy view of a dgCMatrix, field for field the reference's MappedCSC<double>
 * (inst/include/mapped_csc.hpp:8-29) plus an optional float copy of the values. */
typedef struct b200als_csc {
  int32_t n_rows;        /* length of the dimension `idx` indexes (= rows of the fixed factor matrix X) */
  int32_t n_cols;        /* number of columns = number of rows of Y being solved for                   */
  int64_t nnz;
  const int32_t* ptr;    /* [n_cols + 1]  R slot @p                                                    */
  const int32_t* idx;    /* [nnz]         R slot @i, ascending within a column                          */
  const double* val_f64; /* [nnz]         R slot @x, or NULL                                           */
  const float* val_f32;  /* [nnz]         optional float values (used when val_f64 is NULL)            */
} b200als_csc;

const char* b200als_last_error(void);
int b200als_version(void);
/* Number of visible CUDA devices (0 and B200ALS_ECUDA when there is none). */
int b200als_device_count(int* count);
int b200als_set_device(int device);

/* Host programs without their own CUDA binding (the R shim, bench.py): page-locked host buffers, a
 * CUDA-event timer on the engine's stream (start: device sync + record; stop: record + sync, returns
 * elapsed device milliseconds) and the number of kernels this library has launched so far. */
int b200als_host_alloc(size_t bytes, void** out);
int b200als_host_free(void* p);
int b200als_timer_start(void);
int b200als_timer_stop(float* ms);
unsigned long long b200als_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * 1. Stateless half-iteration calls -- same argument lists as the reference's Rcpp exports
 *    (src/wrmf_implicit.cpp:5-31, src/wrmf_explicit.cpp:5-27) with the S4/SEXP objects flattened:
 *    `X` is rank x m_csc->n_rows (read-only), `Y` is rank x m_csc->n_cols (updated IN PLACE, as the
 *    reference does, R/model_WRMF.R:492), `XtX` is rank x rank and already contains +lambda*I
 *    (R/model_WRMF.R:476-484); pass XtX = NULL to have the engine compute XX^T + lambda*I on the GPU.
 *    All pointers are HOST pointers; the call copies in, computes on the current device, copies Y
 *    back and retains nothing.  `*loss` receives the reference's return value (loss / nnz).
 *    `n_threads` is accepted for signature compatibility and ignored.
 *    solver = B200ALS_NNLS runs c_nnls (inst/include/nnls.hpp:10-48) on the GPU.
 *    Bias terms, with the reference's conventions:
 *      with_biases: X and Y carry `rank` = R's private$rank rows (user rank + 2, R/model_WRMF.R:162-166),
 *        is_x_bias_last_row:  X = [1, ..., x_bias]   Y = [y_bias, ..., 1]
 *        otherwise:           X = [x_bias, ..., 1]   Y = [1, ..., y_bias]       (wrmf_implicit.hpp:96-101)
 *        the solved system has rank-1 unknowns and XtX is (rank-1) x (rank-1), built from X without its bias row
 *        (R/model_WRMF.R:474-486); the row of ones in Y is left untouched.
 *      global_bias (implicit only): values below sqrt(epsilon) count as 0 (wrmf_implicit.hpp:108-109);
 *        without with_biases `global_bias_base` ([rank], host) is rewritten when initialize_bias_base is set
 *        (:111-112) and read otherwise.
 *      One deliberate difference: implicit + CONJUGATE_GRADIENT + with_biases raises a dimension error in the
 *      reference (`init` loses a row twice, :191 and :199); here it is solved with `init` dropped once.
 * ---------------------------------------------------------------------------------------------- */
int b200als_als_implicit_float(const b200als_csc* m_csc, int rank, const float* X, float* Y,
                               const float* XtX, double lambda, int n_threads, unsigned solver,
                               unsigned cg_steps, int with_biases, int is_x_bias_last_row,
                               double global_bias, float* global_bias_base,
                               int initialize_bias_base, double* loss);
int b200als_als_implicit_double(const b200als_csc* m_csc, int rank, const double* X, double* Y,
                                const double* XtX, double lambda, int n_threads, unsigned solver,
                                unsigned cg_steps, int with_biases, int is_x_bias_last_row,
                                double global_bias, double* global_bias_base,
                                int initialize_bias_base, double* loss);
int b200als_als_explicit_float(const b200als_csc* m_csc, int rank, const float* X, float* Y,
                               const float* cnt_X, double lambda, unsigned n_threads,
                               unsigned solver, unsigned cg_steps, int dynamic_lambda,
                               int with_biases, int is_x_bias_last_row, double* loss);
int b200als_als_explicit_double(const b200als_csc* m_csc, int rank, const double* X, double* Y,
                                const double* cnt_X, double lambda, unsigned n_threads,
                                unsigned solver, unsigned cg_steps, int dynamic_lambda,
                                int with_biases, int is_x_bias_last_row, double* loss);

/* initialize_biases<T> (inst/include/wrmf_utils.hpp:170-183; Rcpp exports src/wrmf_init.cpp:6-34; called by
 * R/model_WRMF.R:260-289): five alternating sweeps that seed the user and item biases.  (csc_ptr [n_item+1],
 * csc_idx, csc_val) is the user x item matrix by item column (@p, @i, @x of c_ui), (csr_ptr [n_user+1], csr_idx,
 * csr_val) the same entries by user (c_iu).  user_bias [n_user] / item_bias [n_item] are in/out host vectors.  For
 * explicit feedback with calculate_global_bias both value arrays are shifted by the mean rating IN PLACE, as the
 * reference does (:48-51).  *global_bias receives the return value (0 unless calculate_global_bias). */
int b200als_initialize_biases_float(int32_t n_user, int32_t n_item, int64_t nnz, const int32_t* csc_ptr,
                                    const int32_t* csc_idx, double* csc_val, const int32_t* csr_ptr,
                                    const int32_t* csr_idx, double* csr_val, float* user_bias, float* item_bias,
                                    double lambda, int dynamic_lambda, int non_negative, int calculate_global_bias,
                                    int is_explicit_feedback, double* global_bias);
int b200als_initialize_biases_double(int32_t n_user, int32_t n_item, int64_t nnz, const int32_t* csc_ptr,
                                     const int32_t* csc_idx, double* csc_val, const int32_t* csr_ptr,
                                     const int32_t* csr_idx, double* csr_val, double* user_bias, double* item_bias,
                                     double lambda, int dynamic_lambda, int non_negative, int calculate_global_bias,
                                     int is_explicit_feedback, double* global_bias);

/* XtX = tcrossprod(X) + lambda*I  (R/model_WRMF.R:474-486, :347-353); X is rank x n host memory. */
int b200als_gram_float(const float* X, int rank, int64_t n, double lambda, float* XtX);

/* top-k recommendation: `top_product` (src/matrix_top_product.cpp:20-102, called through
 * find_top_product R/utils.R:31-59 by MatrixFactorizationRecommender$predict).  R conventions are kept at the
 * boundary: `exclude` and the returned indices are 1-based, missing entries are NA (INT_MIN / R's NA_real_),
 * outputs are n_user x top_k column-major, `glob_mean` is added to every score.  Scores are accumulated in double
 * like the reference (which converts float factors to double, R/utils.R:35-36).  user_emb is rank x n_user
 * (what transform_ produces before R's t()); not_recommend is a CSR over users with ascending 0-based column
 * indices (the @p / @j slots of a dgRMatrix) or NULL.  Limits of this engine: rank <= 256, top_k <= 128. */
int b200als_top_product(const float* user_emb, int64_t n_user, const float* item_emb, int32_t n_item, int rank,
                        int top_k, const int32_t* not_recommend_ptr, const int32_t* not_recommend_idx,
                        const int32_t* exclude, int n_exclude, double glob_mean, int32_t* idx_out,
                        double* scores_out);

/* ------------------------------------------------------------------------------------------------
 * 2. Session API -- the device-resident form of WRMF$fit_transform (R/model_WRMF.R:173-360).
 *    The sparse matrix is uploaded once in both orientations, both factor matrices stay in HBM,
 *    and the ALS loop (item half, user half, convergence test of R/model_WRMF.R:318-338) runs
 *    without touching the host.  fp32 only (north star).
 * ---------------------------------------------------------------------------------------------- */
typedef struct b200als_session b200als_session;

typedef struct b200als_options {
  int feedback;        /* B200ALS_IMPLICIT / B200ALS_EXPLICIT                                     */
  int solver;          /* wrmf.hpp:16-18                                                         */
  int cg_steps;        /* R default 3                                                            */
  int dynamic_lambda;  /* explicit feedback only (wrmf_explicit.hpp:78)                          */
  double lambda;
  int kernel;          /* kernel choice: 0 = auto.  1 = generic streaming kernels (CG and Cholesky; the reference's arithmetic,
                          full XtX).  CG only: 2 = register-resident kernel with the full XtX (rank 128), 3 = eigenbasis of
                          XtX even for small inputs (implicit; any rank % 4 == 0 up to 256), 10 = as 3 but without the
                          register-resident kernel (every row goes to the shared-memory tile kernel or, if too long, the
                          streaming kernel).  Cholesky only (rank 64 / 128): 0 = row-per-thread panel kernel with the per-row
                          Gram on tcgen05 (rank 128) / warp per system (rank 64); 4 = row-per-thread panel kernel with the
                          fp32 FFMA2 Gram at both ranks                                                                  */
  int reserved[7];     /* reserved[0]: tile staging of the register-resident CG kernel -- 0 default, 1 cp.async.bulk
                          (TMA engine), 2 cp.async (LDGSTS); reserved[1]: CTAs/SM the kernel is built for (CG resident
                          kernel: 0 default, 3, 4; rank-128 FFMA2-Gram Cholesky: 0 default (= 3), 2, 3); reserved[2]: arithmetic
                          of XtX -- 0 default (environment B200ALS_GRAM, else 3xTF32 on tcgen05 at rank 128 / 256), 1 = bf16
                          operands on tcgen05 (fp32 accumulate), 2 = fp32 FMA, 3 = 3xTF32 explicitly; the rest must be 0    */
} b200als_options;

void b200als_default_options(b200als_options* o);

/* c_ui: CSC of users x items (columns = items, idx = users)  -> item half-iteration
 * c_iu: CSC of items x users (columns = users, idx = items)  -> user half-iteration
 * Either may be NULL if that half-iteration is never run.  Host pointers; copied to the device.
 * With a communicator (section 3) each rank passes only ITS block of columns for each
 * orientation plus the global offset of that block, and holds full copies of both factor matrices. */
int b200als_create(b200als_session** out, const b200als_csc* c_ui, const b200als_csc* c_iu,
                   int32_t n_user, int32_t n_item, int rank, const b200als_options* opts);
int b200als_destroy(b200als_session* s);

/* Format ingest on the device (the host-side MatrixExtra::as.csr.matrix / t_shallow of
 * R/model_WRMF.R:184-189): a session created with ONE orientation builds the other one in HBM with a stable
 * radix sort, ascending indices inside every column.  Single GPU only.
 * b200als_get_orientation copies an orientation back (ptr[n_cols+1], idx[nnz], val[nnz]; any may be NULL). */
int b200als_build_missing_orientation(b200als_session* s);
int b200als_get_orientation(b200als_session* s, int which, int32_t* ptr, int32_t* idx, float* val, int64_t* nnz_out);

/* Copy a full factor matrix (rank x n, host memory) to / from the session. */
int b200als_set_factors(b200als_session* s, int which, const float* host);
int b200als_get_factors(b200als_session* s, int which, float* host);
/* Initialise like R/model_WRMF.R:203-244: users ~ N(0,1)/100, items zero for CG / N(0,1)/100 else. */
int b200als_init_factors(b200als_session* s, uint64_t seed);
/* Fill one factor matrix with N(0,1) * scale * (1+f)^-decay (f = feature index) on the device.  decay > 0
 * gives a trained-like, ill-conditioned XtX so that the fixed-step CG really takes all its steps (i.i.d.
 * factors make XtX ~ c*I and CG exits after one step through the rsnew < CG_TOL test). */
int b200als_randomize_factors(b200als_session* s, int which, uint64_t seed, float scale, float decay);

/* One half-iteration solving for `which` (B200ALS_ITEMS / B200ALS_USERS); solver_override < 0
 * keeps the session's solver, otherwise uses the given code (the avoid_cg path of
 * R/model_WRMF.R:112,:412-452 passes B200ALS_CHOLESKY).  `*loss` may be NULL. */
int b200als_half_iteration(b200als_session* s, int which, int solver_override, double* loss);

/* The loop of R/model_WRMF.R:318-338: up to n_iter (items, users) pairs, stopping when
 * loss_prev / loss - 1 < convergence_tol.  loss_trace (may be NULL) receives 2 values per
 * iteration run; *n_iter_done the count of iterations run. */
int b200als_fit(b200als_session* s, int n_iter, double convergence_tol, double* loss_trace,
                int* n_iter_done);

/* transform_ (R/model_WRMF.R:412-452): solve the user half with Y = 0 and CG replaced by
 * Cholesky against the session's item factors; writes rank x n_user to `host_out` without
 * disturbing the session's user factors. */
int b200als_transform(b200als_session* s, float* host_out, double* loss);

/* Device time (ms, CUDA events on the engine's stream) of the last half-iteration, split into
 * gram (XtX), prepare (basis change) and solve; any pointer may be NULL. */
int b200als_last_timing(b200als_session* s, float* gram_ms, float* prep_ms, float* solve_ms,
                        float* comm_ms);

/* ------------------------------------------------------------------------------------------------
 * 3. Multi-GPU: one process per GPU.  Rank 0 creates an id, the host program distributes the
 *    128 bytes (torch.distributed / MPI / a file), every rank calls b200als_comm_init.  A session
 *    created afterwards shards the solved rows by contiguous blocks and all-gathers the updated
 *    factor slices over NCCL once per half-iteration.
 * ---------------------------------------------------------------------------------------------- */
#define B200ALS_UNIQUE_ID_BYTES 128
int b200als_comm_unique_id(void* id_out /* 128 bytes */);
int b200als_comm_init(const void* id, int rank, int world_size);
int b200als_comm_destroy(void);
int b200als_comm_info(int* rank, int* world_size);
/* How the last sharded half-iteration exchanged the solved rows: 0 = no exchange yet / single GPU, 1 = peer-memory
 * pushes (CUDA IPC mappings of the peers' factor matrices, copy engines over NVLink), 2 = grouped NCCL broadcasts
 * (the fallback when the peers cannot be mapped; B200ALS_EXCHANGE=nccl|p2p in the environment forces either). */
int b200als_exchange_mode(b200als_session* s, int* mode);
/* Tell a session which global column range [begin, end) of each orientation this rank owns. */
int b200als_set_shard(b200als_session* s, int which, int32_t begin, int32_t end);

/* ------------------------------------------------------------------------------------------------
 * 4. Synthetic workload generator used by bench.py and the tests: every
 *    row draws exactly nnz_per_row distinct ascending column ids (counter-based hash, seed) and a
 *    positive confidence (implicit) or a rating in 1..5 (explicit).  Fills a session's c_iu
 *    orientation directly in HBM, or host arrays when the *_host pointers are given.
 * ---------------------------------------------------------------------------------------------- */
int b200als_synth_csr_host(int32_t n_rows, int32_t n_cols, int32_t nnz_per_row, uint64_t seed,
                           int explicit_values, int64_t row_offset, int32_t* ptr, int32_t* idx,
                           float* val_f32, double* val_f64);
int b200als_create_synthetic(b200als_session** out, int32_t n_user_local, int64_t user_offset,
                             int32_t n_user_global, int32_t n_item, int32_t nnz_per_row,
                             uint64_t seed, int rank, const b200als_options* opts);
/* Bias terms inside the session (R/model_WRMF.R:260-297; wrmf_implicit.hpp:105-157, wrmf_explicit.hpp:57-64): with
 * with_user_item_bias the session must have been created with rank = R's private$rank = rank + 2 and the factor matrices
 * carry the reference's layouts (users [1, ..., user_bias], items [item_bias, ..., 1]); global_bias is the value
 * initialize_biases returned (implicit feedback; explicit feedback subtracts it from the values beforehand, as R does).
 * Every half-iteration, fit and transform of the session then runs the bias-aware kernels on the device. Single GPU. */
int b200als_set_bias(b200als_session* s, int with_user_item_bias, double global_bias);

/* Which kernel took how many rows in the last CG half-iteration of orientation `which` (bench.py reports it):
 * counts[0] register-resident kernel; [1..4] shared-memory tile kernel -- 4 warps x 4 CTAs/SM double- / single-buffered,
 * 8 warps x 2 CTAs/SM single-buffered, 16 warps x 1 CTA/SM double-buffered; [5..7] the same kernel on thread-block clusters
 * of 2 / 4 / 8 CTAs (opt-in); [8] long rows (rank 128: per-row Gram on tcgen05, als_cg_gram_kernel; otherwise the
 * streaming kernel); [9] empty rows.  caps[0..8] = longest row each class takes.  nnz_local = entries of the local block. */
int b200als_row_plan(b200als_session* s, int which, int32_t counts[10], int32_t caps[9], int64_t* nnz_local);

/* Skewed synthetic data for the robustness points of bench.py: col_dist 0 = one id per equal-width stratum
 * (as above), 1 = Zipf(1.0)-like popularity (ids log-uniform over [0, n_item), made distinct and ascending per row);
 * len_dist 0 = exactly nnz_per_row entries per row, 1 = log-normal row lengths (sigma = 1) with mean nnz_per_row.
 * A row's content depends only on its global id, so shards of different world sizes hold the same matrix. */
int b200als_create_synthetic_ex(b200als_session** out, int32_t n_user_local, int64_t user_offset,
                                int32_t n_user_global, int32_t n_item, int32_t nnz_per_row,
                                uint64_t seed, int rank, const b200als_options* opts, int col_dist, int len_dist);

#ifdef __cplusplus
}
#endif
#endif /* B200ALS_H */
