// engine_stateless.inl -- part of engine.cu (included there; not a standalone translation unit).
// ------------------------------------------------------------------------------------------------------
// 1. stateless calls
// ------------------------------------------------------------------------------------------------------
// Bias arguments of the reference entry points (src/wrmf_implicit.cpp:5-31, src/wrmf_explicit.cpp:5-27).
template <typename T>
struct BiasArgs {
  int with_biases = 0, is_x_bias_last_row = 0;
  double global_bias = 0.0;
  T* global_bias_base = nullptr;   // host, [rank - with_biases], in/out
  int initialize_bias_base = 0;
};

// `rank` = rows of X and Y as the caller holds them (R's private$rank: rank + 2 with biases, model_WRMF.R:162-166).
// With biases the reference solves a (rank-1)-sized system on row-dropped views (drop_row, wrmf_utils.hpp:3-10):
//   is_x_bias_last_row:  X = [1, ..., x_bias]   Y = [y_bias, ..., 1]     X_nnz = X rows 0..rank-2, x_biases = last row
//   otherwise:           X = [x_bias, ..., 1]   Y = [1, ..., y_bias]     X_nnz = X rows 1..rank-1, x_biases = first row
// Here the views are materialised once on the device as compact matrices Xc (n_src x k), xb (n_src), Yc (n_tgt x k),
// k = rank - 1, the generic kernels run on those, and the solved rows are scattered back into Y.
//
// half_on_device: everything between the uploads and the downloads of a half-iteration with (or without) bias terms, on
// DEVICE buffers -- shared by the stateless calls (host pointers around it) and the device-resident session.
//   dX: n_src x rank (fixed), dY: n_tgt x rank (solved in place), dG: XtX of the (rank - with_biases)-wide view or nullptr
//   (computed here), d_gbb: device copy of global_bias_base ([ks], in/out; only without user/item biases).
// On return `o` carries the regulariser window of the loss (reg_ld / reg_lo / reg_hi) for finish_loss.
template <typename T>
struct BiasScratch {
  DevBuf Xc, Yc, Xb, Rhs, G;
};
template <typename T>
static int half_on_device(Ctx& c, CscDev<T>& D, int rank, const T* dX, T* dY, long long n_src, long long n_tgt, const T* dG,
                          HalfOpts& o, int with_biases, int is_x_bias_last_row, double global_bias, T* d_gbb,
                          int initialize_bias_base, BiasScratch<T>& W) {
  const bool implicit = (o.feedback == B200ALS_IMPLICIT);
  const bool wb = with_biases != 0, is_last = is_x_bias_last_row != 0;
  double gbias = implicit ? global_bias : 0.0;
  if (gbias < std::sqrt((double)std::numeric_limits<T>::epsilon())) gbias = 0.0;          // wrmf_implicit.hpp:108-109
  if (wb && rank < 2) return fail(B200ALS_EINVAL, "with_biases needs at least 2 rows in X / Y");
  if (!wb && gbias != 0.0 && !d_gbb) return fail(B200ALS_EINVAL, "global_bias needs global_bias_base");
  const int ks = wb ? rank - 1 : rank;           // size of the solved system
  const int xo = (wb && !is_last) ? 1 : 0;       // X_nnz = drop_row(X_nnz, is_x_bias_last_row)          (:190 / :88)
  const int xbcol = is_last ? rank - 1 : 0;      // x_biases                                             (:115-119)
  const int io = (wb && is_last) ? 1 : 0;        // init = drop_row(init, !is_x_bias_last_row), sic      (:191 / :90)
  const int oo = (wb && !is_last) ? 1 : 0;       // Y.head(rank-1) / Y.tail(rank-1)                      (:240-252)
  const T* Xs = dX;   // what the kernels gather from
  T* Ys = dY;         // what they solve in place
  const int cp_grid = c.sm_count * 8;
  if (wb) {
    CU(W.Xc.ensure(sizeof(T) * (size_t)ks * (size_t)std::max<long long>(1, n_src)));
    CU(W.Yc.ensure(sizeof(T) * (size_t)ks * (size_t)std::max<long long>(1, n_tgt)));
    CU(W.Xb.ensure(sizeof(T) * (size_t)std::max<long long>(1, n_src)));
    pack_cols_kernel<T><<<cp_grid, 256, 0, c.stream>>>(dX, rank, xo, ks, n_src, W.Xc.template as<T>());
    LAUNCHED(); CU(cudaGetLastError());
    pack_cols_kernel<T><<<cp_grid, 256, 0, c.stream>>>(dX, rank, xbcol, 1, n_src, W.Xb.template as<T>());
    LAUNCHED(); CU(cudaGetLastError());
    pack_cols_kernel<T><<<cp_grid, 256, 0, c.stream>>>(dY, rank, io, ks, n_tgt, W.Yc.template as<T>());
    LAUNCHED(); CU(cudaGetLastError());
    Xs = W.Xc.template as<T>();
    Ys = W.Yc.template as<T>();
    o.with_biases = 1;
    o.xbias = W.Xb.p;
    o.reg_ld = rank;                     // every learned row of X: all but the row of ones (:286-302 / :148-172)
    o.reg_lo = is_last ? 1 : 0;
    o.reg_hi = is_last ? rank : rank - 1;
  }
  o.gbias = gbias;
  const T* G = nullptr;
  if (implicit) {
    if (dG) {
      G = dG;
    } else {
      CU(W.G.ensure(sizeof(T) * (size_t)ks * ks));
      TRY(run_gram<T>(c, Xs, ks, n_src, o.lambda, W.G.template as<T>(), nullptr));   // R/model_WRMF.R:474-486
      G = W.G.template as<T>();
    }
    if (wb || gbias != 0.0) {
      // rhs_init = -X_nnz-view * (x_biases + global_bias) (:143-154) ; global_bias_base = sum(X, 1) * (-global_bias) (:111-112)
      CU(W.Rhs.ensure(sizeof(T) * (size_t)ks));
      const bool compute = wb || initialize_bias_base;
      if (compute) {
        if (ks > 256) return fail(B200ALS_EUNSUPPORTED, "bias terms: rank > 256 is not supported");
        const int cs_grid = c.sm_count * 4;
        CU(c.reg_partials.ensure(sizeof(double) * (size_t)cs_grid * ks));
        weighted_colsum_kernel<T><<<cs_grid, 256, 0, c.stream>>>(Xs, ks, n_src, wb ? W.Xb.template as<T>() : nullptr,
                                                               wb ? (T)gbias : T(1), c.reg_partials.f64());
        LAUNCHED(); CU(cudaGetLastError());
        finish_colsum_kernel<T><<<(ks + 127) / 128, 128, 0, c.stream>>>(c.reg_partials.f64(), cs_grid, ks, wb ? -1.0 : -gbias,
                                                                       W.Rhs.template as<T>());
        LAUNCHED(); CU(cudaGetLastError());
        if (!wb) CU(cudaMemcpyAsync(d_gbb, W.Rhs.p, sizeof(T) * (size_t)ks, cudaMemcpyDeviceToDevice, c.stream));
      } else {
        CU(cudaMemcpyAsync(W.Rhs.p, d_gbb, sizeof(T) * (size_t)ks, cudaMemcpyDeviceToDevice, c.stream));
      }
      o.rhs_init = W.Rhs.p;
    }
  }
  TRY(solve_rows<T>(c, D, Xs, Ys, G, nullptr, ks, o));
  if (wb) {
    unpack_cols_kernel<T><<<cp_grid, 256, 0, c.stream>>>(W.Yc.template as<T>(), ks, n_tgt, dY, rank, oo);
    LAUNCHED(); CU(cudaGetLastError());
  }
  return B200ALS_OK;
}

template <typename T>
static int stateless_half(const b200als_csc* A, int rank, const T* X, T* Y, const T* XtX, const T* cnt_X, HalfOpts o,
                          double* loss, const BiasArgs<T>& ba = BiasArgs<T>()) {
  Ctx& c = ctx();
  TRY(c.init());
  if (!A || !X || !Y) return fail(B200ALS_EINVAL, "null argument");
  if (rank <= 0) return fail(B200ALS_EINVAL, "rank must be positive");
  const bool implicit = (o.feedback == B200ALS_IMPLICIT);
  const bool wb = ba.with_biases != 0;
  const int ks = wb ? rank - 1 : rank;
  CscDev<T> D;
  TRY(upload_csc<T>(A, D, c.stream));
  const size_t k = (size_t)rank;
  const size_t n_src = (size_t)A->n_rows, n_tgt = (size_t)A->n_cols;
  DevBuf dX, dY, dG, dCnt, dGbb;
  BiasScratch<T> W;
  CU(dX.ensure(sizeof(T) * k * n_src));
  CU(dY.ensure(sizeof(T) * k * n_tgt));
  CU(cudaMemcpyAsync(dX.p, X, sizeof(T) * k * n_src, cudaMemcpyHostToDevice, c.stream));
  CU(cudaMemcpyAsync(dY.p, Y, sizeof(T) * k * n_tgt, cudaMemcpyHostToDevice, c.stream));
  const T* G = nullptr;
  if (implicit && XtX) {
    CU(dG.ensure(sizeof(T) * (size_t)ks * ks));
    CU(cudaMemcpyAsync(dG.p, XtX, sizeof(T) * (size_t)ks * ks, cudaMemcpyHostToDevice, c.stream));
    G = dG.template as<T>();
  }
  T* d_gbb = nullptr;
  if (implicit && !wb && ba.global_bias_base) {
    CU(dGbb.ensure(sizeof(T) * (size_t)ks));
    CU(cudaMemcpyAsync(dGbb.p, ba.global_bias_base, sizeof(T) * (size_t)ks, cudaMemcpyHostToDevice, c.stream));
    d_gbb = dGbb.template as<T>();
  }
  const T* dcnt = nullptr;
  if (o.feedback == B200ALS_EXPLICIT && o.dynamic_lambda && o.lambda > 0) {
    if (!cnt_X) return fail(B200ALS_EINVAL, "explicit feedback with dynamic_lambda needs cnt_X");
    CU(dCnt.ensure(sizeof(T) * n_src));
    CU(cudaMemcpyAsync(dCnt.p, cnt_X, sizeof(T) * n_src, cudaMemcpyHostToDevice, c.stream));
    dcnt = dCnt.template as<T>();
  }
  TRY(half_on_device<T>(c, D, rank, dX.template as<T>(), dY.template as<T>(), (long long)n_src, (long long)n_tgt, G, o,
                        ba.with_biases, ba.is_x_bias_last_row, ba.global_bias, d_gbb, ba.initialize_bias_base, W));
  if (d_gbb && o.gbias != 0.0 && ba.initialize_bias_base)
    CU(cudaMemcpyAsync(ba.global_bias_base, d_gbb, sizeof(T) * (size_t)ks, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaMemcpyAsync(Y, dY.p, sizeof(T) * k * n_tgt, cudaMemcpyDeviceToHost, c.stream));
  TRY(finish_loss<T>(c, dX.template as<T>(), rank, A->n_rows, dcnt, o, A->nnz, 0.0, false, loss));
  return B200ALS_OK;
}

// ------------------------------------------------------------------------------------------------------
// 1b. pipelined stateless call: fp32, CG, rank 128, large inputs.  The solved rows are cut into blocks of
//     <= 512k rows / 64M non-zeros; block c+1 and c+2 travel host->device (copy engine) while block c is
//     classified, rotated, solved and rotated back on the compute stream and block c-1 returns device->host.
//     Device buffers are cached in the context between calls; no data is retained.
// ------------------------------------------------------------------------------------------------------
__global__ void diag_matrix_kernel(const float* __restrict__ d, float* __restrict__ G, int k) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < k * k) G[e] = ((e / k) == (e % k)) ? d[e / k] : 0.f;
}
struct PipeBuf {
  DevBuf ptr, idx, val64, val32, Y, short_list, long_list, counts;
  cudaEvent_t h2d_done = nullptr, compute_done = nullptr, d2h_done = nullptr;
  bool used = false;
};
struct PipeCtx {
  static constexpr int NB = 3;
  cudaStream_t h2d = nullptr, d2h = nullptr;
  PipeBuf buf[NB];
  DevBuf X, G, G64, Vt, Q, Qt, Q64, diag, Gdiag, cnt;
  int init() {
    if (h2d) return B200ALS_OK;
    CU(cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking));
    for (auto& b : buf) {
      CU(cudaEventCreateWithFlags(&b.h2d_done, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&b.compute_done, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&b.d2h_done, cudaEventDisableTiming));
    }
    return B200ALS_OK;
  }
};
static PipeCtx& pipe_ctx() {
  static thread_local PipeCtx p;
  return p;
}
static int rotate_matrix(Ctx& c, float* M, long long n, const float* R, int k);

static int stateless_pipelined(const b200als_csc* A, const float* X, float* Y, const float* XtX, const float* cnt_X,
                               const HalfOpts& o, double* loss) {
  Ctx& c = ctx();
  PipeCtx& pc = pipe_ctx();
  NvtxRange nv("b200als/stateless_pipelined");
  TRY(pc.init());
  const int k = kResK;
  const bool implicit = (o.feedback == B200ALS_IMPLICIT);
  const int32_t* hp = A->ptr;
  // ---- block boundaries from the host row pointers ----
  int64_t kMaxRows = 512 * 1024;
  const int64_t kMaxNnz = 64ll * 1024 * 1024;
  if (const char* er = getenv("B200ALS_PIPELINE_ROWS")) kMaxRows = std::max<int64_t>(1, atoll(er));  // tests: force many blocks
  std::vector<int32_t> cuts{0};
  int64_t max_rows = 0, max_nnz = 0;
  while (cuts.back() < A->n_cols) {
    const int32_t b = cuts.back();
    int32_t e = (int32_t)std::min<int64_t>(A->n_cols, (int64_t)b + kMaxRows);
    while (e > b + 1 && (int64_t)hp[e] - hp[b] > kMaxNnz) e = b + std::max(1, (e - b) / 2);
    cuts.push_back(e);
    max_rows = std::max<int64_t>(max_rows, e - b);
    max_nnz = std::max<int64_t>(max_nnz, (int64_t)hp[e] - hp[b]);
  }
  const int n_chunks = (int)cuts.size() - 1;
  for (auto& b : pc.buf) {
    CU(b.ptr.ensure(sizeof(int32_t) * (size_t)(max_rows + 1)));
    CU(b.idx.ensure(sizeof(int32_t) * (size_t)max_nnz));
    if (A->val_f64) CU(b.val64.ensure(sizeof(double) * (size_t)max_nnz));
    CU(b.val32.ensure(sizeof(float) * (size_t)max_nnz));
    CU(b.Y.ensure(sizeof(float) * (size_t)max_rows * k));
    CU(b.short_list.ensure(sizeof(int32_t) * (size_t)max_rows));
    CU(b.long_list.ensure(sizeof(int32_t) * (size_t)max_rows));
    CU(b.counts.ensure(4 * sizeof(int)));
    b.used = false;
  }
  // ---- fixed matrix, Gram, eigenbasis (compute stream) ----
  const size_t xbytes = sizeof(float) * (size_t)k * (size_t)A->n_rows;
  CU(pc.X.ensure(xbytes));
  // One process per GPU, every rank calling with its own block of rows and the SAME fixed matrix: with
  // B200ALS_STATELESS_SHARE_FIXED=1 only rank 0 uploads it and the others receive it over NVLink (ncclBroadcast) --
  // on a host whose memory system, not PCIe, bounds N concurrent uploads that removes (N - 1) x the matrix from the host path.
  // Opt-in: the caller asserts that all ranks pass identical X (the engine cannot check it).
  const char* esh = getenv("B200ALS_STATELESS_SHARE_FIXED");
  if (g_comm.world > 1 && g_comm.comm && esh && esh[0] == '1') {
    if (g_comm.rank == 0) CU(cudaMemcpyAsync(pc.X.p, X, xbytes, cudaMemcpyHostToDevice, c.stream));
    NC(g_nccl.Broadcast(pc.X.p, pc.X.p, (size_t)k * (size_t)A->n_rows, ncclFloat, 0, g_comm.comm, c.stream));
  } else {
    CU(cudaMemcpyAsync(pc.X.p, X, xbytes, cudaMemcpyHostToDevice, c.stream));
  }
  const float* diag = nullptr;
  const float* Glong = nullptr;
  if (implicit) {
    CU(pc.G.ensure(sizeof(float) * k * k));
    CU(pc.G64.ensure(sizeof(double) * k * k));
    CU(pc.Vt.ensure(sizeof(double) * k * k));
    CU(pc.Q64.ensure(sizeof(double) * k * k));
    CU(pc.Q.ensure(sizeof(float) * k * k));
    CU(pc.Qt.ensure(sizeof(float) * k * k));
    CU(pc.diag.ensure(sizeof(float) * k));
    CU(pc.Gdiag.ensure(sizeof(float) * k * k));
    if (XtX) {
      CU(cudaMemcpyAsync(pc.G.p, XtX, sizeof(float) * k * k, cudaMemcpyHostToDevice, c.stream));
      convert_kernel<float, double><<<(k * k + 255) / 256, 256, 0, c.stream>>>(pc.G.f32(), pc.G64.f64(), k * k);
      LAUNCHED(); CU(cudaGetLastError());
    } else {
      TRY(run_gram<float>(c, pc.X.f32(), k, A->n_rows, o.lambda, pc.G.f32(), pc.G64.f64()));
    }
    const size_t jsm = sizeof(double) * (size_t)k * (k + 1);
    CU(cudaFuncSetAttribute(jacobi_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jsm));
    jacobi_eig_kernel<<<1, kJacobiThreads, jsm, c.stream>>>(pc.G64.f64(), pc.Vt.f64(), k, pc.Q.f32(), pc.diag.f32(),
                                                            pc.Q64.f64(), 30, 1);
    LAUNCHED(); CU(cudaGetLastError());
    convert_kk_kernel<<<(k * k + 255) / 256, 256, 0, c.stream>>>(pc.Q64.f64(), pc.Qt.f32(), k, 1);
    LAUNCHED(); CU(cudaGetLastError());
    diag_matrix_kernel<<<(k * k + 255) / 256, 256, 0, c.stream>>>(pc.diag.f32(), pc.Gdiag.f32(), k);
    LAUNCHED(); CU(cudaGetLastError());
    TRY(rotate_matrix(c, pc.X.f32(), A->n_rows, pc.Q.f32(), k));
    diag = pc.diag.f32();
    Glong = pc.Gdiag.f32();
  }
  const float* dcnt = nullptr;
  if (!implicit && o.dynamic_lambda && o.lambda > 0) {
    if (!cnt_X) return fail(B200ALS_EINVAL, "explicit feedback with dynamic_lambda needs cnt_X");
    CU(pc.cnt.ensure(sizeof(float) * (size_t)A->n_rows));
    CU(cudaMemcpyAsync(pc.cnt.p, cnt_X, sizeof(float) * (size_t)A->n_rows, cudaMemcpyHostToDevice, c.stream));
    dcnt = pc.cnt.f32();
  }
  CU(cudaMemsetAsync(c.loss_acc.p, 0, sizeof(double), c.stream));
  CU(cudaMemsetAsync(c.status.p, 0, sizeof(int), c.stream));
  const int res_grid = c.sm_count * 3;
  const int gen_grid = c.sm_count * 4;
  CU(c.loss_partials.ensure(sizeof(double) * (size_t)c.sm_count * 8));
  const size_t res_smem = sizeof(ResidentSmem);
  CU(cudaFuncSetAttribute(als_cg_resident_kernel<false, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)res_smem));
  // ---- the pipeline ----
  for (int ci = 0; ci < n_chunks; ci++) {
    PipeBuf& b = pc.buf[ci % PipeCtx::NB];
    const int32_t r0 = cuts[ci], r1 = cuts[ci + 1], nr = r1 - r0;
    const int64_t e0 = hp[r0], ne = (int64_t)hp[r1] - e0;
    // host -> device
    if (b.used) CU(cudaStreamWaitEvent(pc.h2d, b.d2h_done, 0));
    CU(cudaMemcpyAsync(b.ptr.p, hp + r0, sizeof(int32_t) * (size_t)(nr + 1), cudaMemcpyHostToDevice, pc.h2d));
    if (ne) {
      CU(cudaMemcpyAsync(b.idx.p, A->idx + e0, sizeof(int32_t) * (size_t)ne, cudaMemcpyHostToDevice, pc.h2d));
      if (A->val_f64) CU(cudaMemcpyAsync(b.val64.p, A->val_f64 + e0, sizeof(double) * (size_t)ne, cudaMemcpyHostToDevice, pc.h2d));
      else CU(cudaMemcpyAsync(b.val32.p, A->val_f32 + e0, sizeof(float) * (size_t)ne, cudaMemcpyHostToDevice, pc.h2d));
    }
    CU(cudaMemcpyAsync(b.Y.p, Y + (size_t)r0 * k, sizeof(float) * (size_t)nr * k, cudaMemcpyHostToDevice, pc.h2d));
    CU(cudaEventRecord(b.h2d_done, pc.h2d));
    // compute
    CU(cudaStreamWaitEvent(c.stream, b.h2d_done, 0));
    if (A->val_f64 && ne) {
      convert_kernel<double, float><<<(unsigned)((ne + 255) / 256), 256, 0, c.stream>>>(b.val64.f64(), b.val32.f32(), ne);
      LAUNCHED(); CU(cudaGetLastError());
    }
    CU(cudaMemsetAsync(b.counts.p, 0, 4 * sizeof(int), c.stream));
    classify_rows_kernel<<<(nr + 255) / 256, 256, 0, c.stream>>>(b.ptr.i32(), nr, kResMaxN, b.short_list.i32(), b.long_list.i32(), b.counts.i32());
    LAUNCHED(); CU(cudaGetLastError());
    zero_empty_rows_kernel<float><<<(unsigned)(((long long)nr * k + 255) / 256), 256, 0, c.stream>>>(b.ptr.i32(), nr, k, b.Y.f32());
    LAUNCHED(); CU(cudaGetLastError());
    if (implicit) TRY(rotate_matrix(c, b.Y.f32(), nr, pc.Q.f32(), k));
    ResidentParams R;
    R.ptr = b.ptr.i32(); R.idx = b.idx.i32(); R.val = b.val32.f32();
    R.X = pc.X.f32(); R.Y = b.Y.f32(); R.diag = diag; R.G = nullptr;
    R.feedback = o.feedback; R.cg_steps = o.cg_steps; R.dynamic_lambda = o.dynamic_lambda; R.lambda = (float)o.lambda;
    R.row_list = b.short_list.i32(); R.n_list = 0; R.n_list_dev = b.counts.i32(); R.ptr_base = (int)e0; R.row_begin = 0;
    R.loss_partials = c.loss_partials.f64();
    als_cg_resident_kernel<false, 1, 3><<<res_grid, kResThreads, res_smem, c.stream>>>(R);
    LAUNCHED(); CU(cudaGetLastError());
    sum_partials_kernel<<<1, 32, 0, c.stream>>>(c.loss_partials.f64(), res_grid, c.loss_acc.f64(), 1);
    LAUNCHED(); CU(cudaGetLastError());
    {  // rows longer than the register tile: streaming kernel on the same (rotated) data
      SolveParams<float> P{};
      P.one_minus_g = 1.f;
      P.ptr = b.ptr.i32(); P.idx = b.idx.i32(); P.val = b.val32.f32(); P.X = pc.X.f32(); P.Y = b.Y.f32();
      P.G = implicit ? Glong : nullptr; P.k = k; P.n_targets = nr; P.feedback = o.feedback; P.cg_steps = o.cg_steps;
      P.dynamic_lambda = o.dynamic_lambda; P.solver = 0; P.lambda = o.lambda; P.row_list = b.long_list.i32(); P.n_list = 0;
      P.n_list_dev = b.counts.i32() + 1; P.ptr_base = (int)e0; P.row_begin = 0; P.ticket = c.ticket.u64();
      P.loss_partials = c.loss_partials.f64(); P.status = c.status.i32();
      CU(cudaMemsetAsync(c.ticket.p, 0, sizeof(unsigned long long), c.stream));
      als_cg_generic_kernel<float, 4><<<gen_grid, 256, 0, c.stream>>>(P);
      LAUNCHED(); CU(cudaGetLastError());
      sum_partials_kernel<<<1, 32, 0, c.stream>>>(c.loss_partials.f64(), gen_grid, c.loss_acc.f64(), 1);
      LAUNCHED(); CU(cudaGetLastError());
    }
    if (implicit) TRY(rotate_matrix(c, b.Y.f32(), nr, pc.Qt.f32(), k));
    CU(cudaEventRecord(b.compute_done, c.stream));
    // device -> host
    CU(cudaStreamWaitEvent(pc.d2h, b.compute_done, 0));
    CU(cudaMemcpyAsync(Y + (size_t)r0 * k, b.Y.p, sizeof(float) * (size_t)nr * k, cudaMemcpyDeviceToHost, pc.d2h));
    CU(cudaEventRecord(b.d2h_done, pc.d2h));
    b.used = true;
  }
  TRY(finish_loss<float>(c, pc.X.f32(), k, A->n_rows, dcnt, o, A->nnz, 0.0, false, loss));
  CU(cudaStreamSynchronize(pc.d2h));
  CU(cudaStreamSynchronize(pc.h2d));
  return B200ALS_OK;
}

// large fp32 CG problems at rank 128 take the pipelined path (B200ALS_PIPELINE=0 disables, =1 forces)
static bool use_pipelined(const b200als_csc* m, int rank, const float* X, const float* Y, const HalfOpts& o) {
  if (!m || !X || !Y || !m->ptr || rank != kResK || o.solver != B200ALS_CONJUGATE_GRADIENT) return false;
  if (ctx().init() != B200ALS_OK) return false;
  const char* env = getenv("B200ALS_PIPELINE");
  if (env && env[0] == '0') return false;
  if (env && env[0] == '1') return m->n_cols > 0;
  return m->n_cols >= 200000;
}

// bias terms run on the generic kernels of the plain (non-pipelined) call
extern "C" int b200als_als_implicit_float(const b200als_csc* m, int rank, const float* X, float* Y, const float* XtX,
                                          double lambda, int, unsigned solver, unsigned cg_steps, int with_biases,
                                          int is_x_bias_last_row, double global_bias, float* global_bias_base,
                                          int initialize_bias_base, double* loss) {
  HalfOpts o{B200ALS_IMPLICIT, (int)solver, (int)cg_steps, 0, 0, lambda};
  const bool biased = with_biases || global_bias >= std::sqrt((double)std::numeric_limits<float>::epsilon());
  if (!biased && use_pipelined(m, rank, X, Y, o)) return stateless_pipelined(m, X, Y, XtX, nullptr, o, loss);
  BiasArgs<float> ba{with_biases, is_x_bias_last_row, global_bias, global_bias_base, initialize_bias_base};
  return stateless_half<float>(m, rank, X, Y, XtX, nullptr, o, loss, ba);
}
extern "C" int b200als_als_implicit_double(const b200als_csc* m, int rank, const double* X, double* Y, const double* XtX,
                                           double lambda, int, unsigned solver, unsigned cg_steps, int with_biases,
                                           int is_x_bias_last_row, double global_bias, double* global_bias_base,
                                           int initialize_bias_base, double* loss) {
  HalfOpts o{B200ALS_IMPLICIT, (int)solver, (int)cg_steps, 0, 0, lambda};
  BiasArgs<double> ba{with_biases, is_x_bias_last_row, global_bias, global_bias_base, initialize_bias_base};
  return stateless_half<double>(m, rank, X, Y, XtX, nullptr, o, loss, ba);
}
extern "C" int b200als_als_explicit_float(const b200als_csc* m, int rank, const float* X, float* Y, const float* cnt_X,
                                          double lambda, unsigned, unsigned solver, unsigned cg_steps, int dynamic_lambda,
                                          int with_biases, int is_x_bias_last_row, double* loss) {
  HalfOpts o{B200ALS_EXPLICIT, (int)solver, (int)cg_steps, dynamic_lambda != 0, 0, lambda};
  if (!with_biases && use_pipelined(m, rank, X, Y, o)) return stateless_pipelined(m, X, Y, nullptr, cnt_X, o, loss);
  BiasArgs<float> ba{with_biases, is_x_bias_last_row, 0.0, nullptr, 0};
  return stateless_half<float>(m, rank, X, Y, nullptr, cnt_X, o, loss, ba);
}
extern "C" int b200als_als_explicit_double(const b200als_csc* m, int rank, const double* X, double* Y, const double* cnt_X,
                                           double lambda, unsigned, unsigned solver, unsigned cg_steps, int dynamic_lambda,
                                           int with_biases, int is_x_bias_last_row, double* loss) {
  HalfOpts o{B200ALS_EXPLICIT, (int)solver, (int)cg_steps, dynamic_lambda != 0, 0, lambda};
  BiasArgs<double> ba{with_biases, is_x_bias_last_row, 0.0, nullptr, 0};
  return stateless_half<double>(m, rank, X, Y, nullptr, cnt_X, o, loss, ba);
}

// initialize_biases<T> (wrmf_utils.hpp:170-183; src/wrmf_init.cpp:6-34) -- see bias_init.cuh
template <typename T>
static int initialize_biases_impl(int32_t n_user, int32_t n_item, int64_t nnz, const int32_t* csc_ptr, const int32_t* csc_idx,
                                  double* csc_val, const int32_t* csr_ptr, const int32_t* csr_idx, double* csr_val,
                                  T* user_bias, T* item_bias, double lambda, int dynamic_lambda, int non_negative,
                                  int calculate_global_bias, int is_explicit, double* global_bias) {
  Ctx& c = ctx();
  TRY(c.init());
  if (!csc_ptr || !csr_ptr || !user_bias || !item_bias || n_user < 0 || n_item < 0 || nnz < 0)
    return fail(B200ALS_EINVAL, "bad argument");
  if (nnz > 0 && (!csc_idx || !csc_val || !csr_idx || !csr_val)) return fail(B200ALS_EINVAL, "null matrix slots");
  DevBuf cp, ci, cv, rp, ri, rv, ub, ib, part, scal, um, ua, im, ia;
  const size_t e = (size_t)std::max<int64_t>(1, nnz);
  CU(cp.ensure(sizeof(int32_t) * ((size_t)n_item + 1)));
  CU(rp.ensure(sizeof(int32_t) * ((size_t)n_user + 1)));
  CU(ci.ensure(sizeof(int32_t) * e)); CU(ri.ensure(sizeof(int32_t) * e));
  CU(cv.ensure(sizeof(double) * e)); CU(rv.ensure(sizeof(double) * e));
  CU(ub.ensure(sizeof(T) * (size_t)std::max(1, n_user)));
  CU(ib.ensure(sizeof(T) * (size_t)std::max(1, n_item)));
  const int grid = c.sm_count * 4;
  CU(part.ensure(sizeof(double) * (size_t)grid));
  CU(scal.ensure(sizeof(double) * 4));
  cudaStream_t st = c.stream;
  CU(cudaMemcpyAsync(cp.p, csc_ptr, sizeof(int32_t) * ((size_t)n_item + 1), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(rp.p, csr_ptr, sizeof(int32_t) * ((size_t)n_user + 1), cudaMemcpyHostToDevice, st));
  if (nnz) {
    CU(cudaMemcpyAsync(ci.p, csc_idx, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ri.p, csr_idx, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(cv.p, csc_val, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(rv.p, csr_val, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
  }
  if (n_user) CU(cudaMemcpyAsync(ub.p, user_bias, sizeof(T) * (size_t)n_user, cudaMemcpyHostToDevice, st));
  if (n_item) CU(cudaMemcpyAsync(ib.p, item_bias, sizeof(T) * (size_t)n_item, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(scal.p, 0, sizeof(double) * 4, st));
  double* d_scal = scal.f64();   // [0] sum of values, [1] sum(user_bias), [2] sum(item_bias)
  auto device_sum = [&](auto* v, long long n, double* out) -> int {
    using V = std::remove_pointer_t<decltype(v)>;
    sum_to_partials_kernel<std::remove_const_t<V>><<<grid, 256, 0, st>>>(v, n, part.f64());
    LAUNCHED(); CU(cudaGetLastError());
    sum_partials_kernel<<<1, 32, 0, st>>>(part.f64(), grid, out, 0);
    LAUNCHED(); CU(cudaGetLastError());
    return B200ALS_OK;
  };
  double g = 0.0;
  const unsigned gi = (unsigned)std::max(1, (n_item + 127) / 128), gu = (unsigned)std::max(1, (n_user + 127) / 128);
  if (calculate_global_bias && nnz > 0) {
    TRY(device_sum((const double*)cv.f64(), (long long)nnz, d_scal));
    double s = 0.0;
    CU(cudaMemcpyAsync(&s, d_scal, sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (is_explicit) {
      g = s / (double)nnz;                                                     // mean rating (wrmf_utils.hpp:40-43)
      shift_values_kernel<<<grid, 256, 0, st>>>(cv.f64(), rv.f64(), (long long)nnz, d_scal, 1.0 / (double)nnz);
      LAUNCHED(); CU(cudaGetLastError());
    } else {
      g = s / (s + (double)n_item * (double)n_user - (double)nnz);             // (:91-94)
    }
  }
  if (!is_explicit && non_negative) g = std::fmax(0.0, g);                       // (:95)
  if (is_explicit) {
    for (int iter = 0; iter < 5; iter++) {                                      // (:54-80)
      if (n_item) {
        bias_sweep_explicit_kernel<T><<<gi, 128, 0, st>>>(cp.i32(), ci.i32(), cv.f64(), n_item, ub.template as<T>(),
                                                          ib.template as<T>(), (T)lambda, dynamic_lambda, non_negative);
        LAUNCHED(); CU(cudaGetLastError());
      }
      if (n_user) {
        bias_sweep_explicit_kernel<T><<<gu, 128, 0, st>>>(rp.i32(), ri.i32(), rv.f64(), n_user, ib.template as<T>(),
                                                          ub.template as<T>(), (T)lambda, dynamic_lambda, non_negative);
        LAUNCHED(); CU(cudaGetLastError());
      }
    }
    if (calculate_global_bias && nnz > 0) {   // the reference shifts the caller's values in place (:48-51)
      CU(cudaMemcpyAsync(csc_val, cv.p, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(csr_val, rv.p, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost, st));
    }
  } else {
    CU(um.ensure(sizeof(double) * (size_t)std::max(1, n_user))); CU(ua.ensure(sizeof(double) * (size_t)std::max(1, n_user)));
    CU(im.ensure(sizeof(double) * (size_t)std::max(1, n_item))); CU(ia.ensure(sizeof(double) * (size_t)std::max(1, n_item)));
    const double lam_t = (double)(T)lambda;   // `T lambda` in the reference's signature
    if (n_user) {
      bias_means_implicit_kernel<<<gu, 128, 0, st>>>(rp.i32(), rv.f64(), n_user, n_item, lam_t, um.f64(), ua.f64());
      LAUNCHED(); CU(cudaGetLastError());
    }
    if (n_item) {
      bias_means_implicit_kernel<<<gi, 128, 0, st>>>(cp.i32(), cv.f64(), n_item, n_user, lam_t, im.f64(), ia.f64());
      LAUNCHED(); CU(cudaGetLastError());
    }
    for (int iter = 0; iter < 5; iter++) {                                      // (:130-162)
      if (iter > 0 && n_user) TRY(device_sum((const T*)ub.template as<T>(), (long long)n_user, d_scal + 1));
      if (n_item) {
        bias_sweep_implicit_kernel<T><<<gi, 128, 0, st>>>(cp.i32(), ci.i32(), cv.f64(), n_item, n_user, ub.template as<T>(),
                                                          (iter > 0 && n_user) ? d_scal + 1 : nullptr, im.f64(), ia.f64(), g,
                                                          non_negative, ib.template as<T>());
        LAUNCHED(); CU(cudaGetLastError());
        TRY(device_sum((const T*)ib.template as<T>(), (long long)n_item, d_scal + 2));
      }
      if (n_user) {
        bias_sweep_implicit_kernel<T><<<gu, 128, 0, st>>>(rp.i32(), ri.i32(), rv.f64(), n_user, n_item, ib.template as<T>(),
                                                          n_item ? d_scal + 2 : nullptr, um.f64(), ua.f64(), g, non_negative,
                                                          ub.template as<T>());
        LAUNCHED(); CU(cudaGetLastError());
      }
    }
  }
  if (n_user) CU(cudaMemcpyAsync(user_bias, ub.p, sizeof(T) * (size_t)n_user, cudaMemcpyDeviceToHost, st));
  if (n_item) CU(cudaMemcpyAsync(item_bias, ib.p, sizeof(T) * (size_t)n_item, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (global_bias) *global_bias = g;
  return B200ALS_OK;
}
extern "C" int b200als_initialize_biases_float(int32_t n_user, int32_t n_item, int64_t nnz, const int32_t* csc_ptr,
                                               const int32_t* csc_idx, double* csc_val, const int32_t* csr_ptr,
                                               const int32_t* csr_idx, double* csr_val, float* user_bias, float* item_bias,
                                               double lambda, int dynamic_lambda, int non_negative, int calculate_global_bias,
                                               int is_explicit_feedback, double* global_bias) {
  return initialize_biases_impl<float>(n_user, n_item, nnz, csc_ptr, csc_idx, csc_val, csr_ptr, csr_idx, csr_val, user_bias,
                                       item_bias, lambda, dynamic_lambda, non_negative, calculate_global_bias,
                                       is_explicit_feedback, global_bias);
}
extern "C" int b200als_initialize_biases_double(int32_t n_user, int32_t n_item, int64_t nnz, const int32_t* csc_ptr,
                                                const int32_t* csc_idx, double* csc_val, const int32_t* csr_ptr,
                                                const int32_t* csr_idx, double* csr_val, double* user_bias, double* item_bias,
                                                double lambda, int dynamic_lambda, int non_negative, int calculate_global_bias,
                                                int is_explicit_feedback, double* global_bias) {
  return initialize_biases_impl<double>(n_user, n_item, nnz, csc_ptr, csc_idx, csc_val, csr_ptr, csr_idx, csr_val, user_bias,
                                        item_bias, lambda, dynamic_lambda, non_negative, calculate_global_bias,
                                        is_explicit_feedback, global_bias);
}

extern "C" int b200als_gram_float(const float* X, int rank, int64_t n, double lambda, float* XtX) {
  Ctx& c = ctx();
  TRY(c.init());
  if (!X || !XtX || rank <= 0 || n < 0) return fail(B200ALS_EINVAL, "bad argument");
  DevBuf dX, dG;
  CU(dX.ensure(sizeof(float) * (size_t)rank * (size_t)n));
  CU(dG.ensure(sizeof(float) * (size_t)rank * rank));
  CU(cudaMemcpyAsync(dX.p, X, sizeof(float) * (size_t)rank * (size_t)n, cudaMemcpyHostToDevice, c.stream));
  TRY(run_gram<float>(c, dX.f32(), rank, n, lambda, dG.f32(), nullptr));
  CU(cudaMemcpyAsync(XtX, dG.p, sizeof(float) * (size_t)rank * rank, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  return B200ALS_OK;
}

// pinned host memory + device timers for host programs without a CUDA binding (bench.py, the R shim)
extern "C" int b200als_host_alloc(size_t bytes, void** out) {
  TRY(ctx().init());
  if (!out) return fail(B200ALS_EINVAL, "null out");
  CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return B200ALS_OK;
}
extern "C" int b200als_host_free(void* p) {
  if (p) CU(cudaFreeHost(p));
  return B200ALS_OK;
}
static cudaEvent_t g_t0 = nullptr, g_t1 = nullptr;
extern "C" int b200als_timer_start(void) {
  Ctx& c = ctx();
  TRY(c.init());
  if (!g_t0) { CU(cudaEventCreate(&g_t0)); CU(cudaEventCreate(&g_t1)); }
  CU(cudaDeviceSynchronize());
  CU(cudaEventRecord(g_t0, c.stream));
  return B200ALS_OK;
}
extern "C" int b200als_timer_stop(float* ms) {
  Ctx& c = ctx();
  if (!g_t0 || !ms) return fail(B200ALS_EINVAL, "timer not started");
  CU(cudaEventRecord(g_t1, c.stream));
  CU(cudaEventSynchronize(g_t1));
  CU(cudaDeviceSynchronize());
  CU(cudaEventElapsedTime(ms, g_t0, g_t1));
  return B200ALS_OK;
}
