// engine_solve.inl -- part of engine.cu (included there; not a standalone translation unit).
// ------------------------------------------------------------------------------------------------------
// half-iteration dispatch
// ------------------------------------------------------------------------------------------------------
struct HalfOpts {
  int feedback, solver, cg_steps, dynamic_lambda, kernel;
  double lambda;
  int stage = 0;  // tile staging of the resident kernel: 0 default, 1 cp.async.bulk (UBLKCP), 2 cp.async (LDGSTS)
  int ctas = 0;   // resident CTAs per SM the kernel is compiled for: 0 default, 3 or 4
  int row_begin = 0, row_count = -1;  // solve only rows [row_begin, row_begin + row_count) of the block (-1: all)
  bool reset_loss = true;             // zero the loss accumulator first (false: add to it)
  // bias terms, all on compact matrices (see stateless_half): device pointers of the element type being solved
  int with_biases = 0;
  double gbias = 0.0;                 // global_bias after the sqrt(eps) cut-off (wrmf_implicit.hpp:108-109)
  const void* xbias = nullptr;        // [n_src]
  const void* rhs_init = nullptr;     // [k]
  int reg_ld = 0, reg_lo = 0, reg_hi = 0;  // loss regulariser over columns [lo, hi) of the n_src x ld matrix (0: whole matrix)
};
constexpr int kDefaultCtas = 3;
constexpr int kDefaultStage = 1;  // LDGSTS: measured 6 % faster than the UBLKCP variant on C3 (profiles/)

template <typename T>
static int classify_rows(Ctx& c, CscDev<T>& A) {
  if (A.n_short >= 0) return B200ALS_OK;
  CU(A.short_list.ensure(sizeof(int32_t) * (size_t)std::max(1, A.n_cols)));
  CU(A.long_list.ensure(sizeof(int32_t) * (size_t)std::max(1, A.n_cols)));
  DevBuf counts;
  CU(counts.ensure(3 * sizeof(int)));
  CU(cudaMemsetAsync(counts.p, 0, 3 * sizeof(int), c.stream));
  if (A.n_cols > 0) {
    classify_rows_kernel<<<(A.n_cols + 255) / 256, 256, 0, c.stream>>>(A.ptr.i32(), A.n_cols, kResMaxN,
                                                                      A.short_list.i32(), A.long_list.i32(),
                                                                      counts.i32());
    LAUNCHED(); CU(cudaGetLastError());
  }
  int h[3];
  CU(cudaMemcpyAsync(h, counts.p, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  A.n_short = h[0];
  A.n_long = h[1];
  A.n_empty = h[2];
  A.all_short = (h[0] == A.n_cols);
  return B200ALS_OK;
}

template <typename T, int KPL>
static int launch_cg_generic(Ctx& c, const SolveParams<T>& P, int n_work, int* grid_out) {
  const int grid = (int)std::min<long long>((long long)c.sm_count * 4, std::max(1, (n_work + 7) / 8));
  als_cg_generic_kernel<T, KPL><<<grid, 256, 0, c.stream>>>(P);
  LAUNCHED(); CU(cudaGetLastError());
  *grid_out = grid;
  return B200ALS_OK;
}

// Runs one half-iteration on device data.  `diag`/`rotated`: the caller has put X and Y in the eigenbasis of
// G (implicit CG, rank 128, resident kernel) and passes the eigenvalues.  Accumulates the loss numerator
// (sum over solved rows) into c.loss_acc[0].
template <typename T>
static int solve_rows(Ctx& c, CscDev<T>& A, const T* X, T* Y, const T* G, const float* diag, int k, const HalfOpts& o) {
  if (o.solver != B200ALS_CHOLESKY && o.solver != B200ALS_CONJUGATE_GRADIENT && o.solver != B200ALS_NNLS)
    return fail(B200ALS_EINVAL, "unknown solver code");
  if (o.feedback == B200ALS_IMPLICIT && !G && !diag) return fail(B200ALS_EINVAL, "implicit feedback needs XtX");
  if (o.reset_loss) {
    CU(cudaMemsetAsync(c.loss_acc.p, 0, sizeof(double), c.stream));
    CU(cudaMemsetAsync(c.status.p, 0, sizeof(int), c.stream));
  }
  if (A.n_cols == 0) return B200ALS_OK;
  const bool sub_range = (o.row_count >= 0);
  const int n_rows_here = sub_range ? o.row_count : A.n_cols;
  if (n_rows_here == 0) return B200ALS_OK;
  SolveParams<T> P{};
  const bool biased = o.with_biases || o.gbias != 0.0;
  P.xbias = static_cast<const T*>(o.xbias);
  P.rhs_init = (o.feedback == B200ALS_IMPLICIT) ? static_cast<const T*>(o.rhs_init) : nullptr;
  P.gbias = (T)o.gbias;
  P.one_minus_g = (T)(1 - o.gbias);
  P.solve_empty = (o.feedback == B200ALS_IMPLICIT && biased) ? 1 : 0;
  P.ptr = A.ptr.i32();
  P.idx = A.idx.i32();
  P.val = A.val.template as<T>();
  P.X = X;
  P.Y = Y;
  P.G = (o.feedback == B200ALS_IMPLICIT) ? G : nullptr;
  P.k = k;
  P.n_targets = n_rows_here;
  P.row_begin = sub_range ? o.row_begin : 0;
  P.feedback = o.feedback;
  P.cg_steps = o.cg_steps;
  P.dynamic_lambda = o.dynamic_lambda;
  P.solver = o.solver;
  P.lambda = o.lambda;
  P.row_list = nullptr;
  P.n_list = 0;
  P.n_list_dev = nullptr;
  P.ptr_base = 0;
  P.ticket = c.ticket.u64();
  P.status = c.status.i32();
  const int max_grid = c.sm_count * 8;
  CU(c.loss_partials.ensure(sizeof(double) * (size_t)max_grid));
  P.loss_partials = c.loss_partials.f64();

  auto run_generic_cg = [&](const int32_t* list, int n_list) -> int {
    P.row_list = list;
    P.n_list = n_list;
    const int n_work = list ? n_list : n_rows_here;
    if (n_work == 0) return B200ALS_OK;
    CU(cudaMemsetAsync(c.ticket.p, 0, sizeof(unsigned long long), c.stream));
    int grid = 0;
    if (k <= 32) TRY((launch_cg_generic<T, 1>(c, P, n_work, &grid)));
    else if (k <= 64) TRY((launch_cg_generic<T, 2>(c, P, n_work, &grid)));
    else if (k <= 128) TRY((launch_cg_generic<T, 4>(c, P, n_work, &grid)));
    else if (k <= 256) TRY((launch_cg_generic<T, 8>(c, P, n_work, &grid)));
    else return fail(B200ALS_EUNSUPPORTED, "rank > 256 is not supported");
    sum_partials_kernel<<<1, 32, 0, c.stream>>>(P.loss_partials, grid, c.loss_acc.f64(), 1);
    LAUNCHED(); CU(cudaGetLastError());
    return B200ALS_OK;
  };

  if (o.solver == B200ALS_CHOLESKY || o.solver == B200ALS_NNLS) {
    auto run_generic_chol = [&](const int32_t* list, int n_list) -> int {
      P.row_list = list;
      P.n_list = n_list;
      const int n_work = list ? n_list : n_rows_here;
      if (n_work == 0) return B200ALS_OK;
      const size_t smem = chol_generic_smem_bytes<T>(k, o.solver);
      if (smem > c.smem_optin)
        return fail(B200ALS_EUNSUPPORTED, "cholesky / nnls: rank too large for the shared-memory factorisation (needs " +
                                             std::to_string(smem) + " B)");
      CU(cudaFuncSetAttribute(als_chol_generic_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (227 * 1024) / (smem + 1024)));
      const int grid = std::min(c.sm_count * per_sm, std::max(1, n_work));
      CU(cudaMemsetAsync(c.ticket.p, 0, sizeof(unsigned long long), c.stream));
      als_chol_generic_kernel<T><<<grid, 256, smem, c.stream>>>(P);
      LAUNCHED(); CU(cudaGetLastError());
      sum_partials_kernel<<<1, 32, 0, c.stream>>>(P.loss_partials, grid, c.loss_acc.f64(), 1);
      LAUNCHED(); CU(cudaGetLastError());
      return B200ALS_OK;
    };
    bool tiled = false;
    if constexpr (sizeof(T) == 4)
      tiled = (o.solver == B200ALS_CHOLESKY) && (k == 64 || k == 128) && o.kernel != 1 && !sub_range && !biased;
    if (!tiled) return run_generic_chol(nullptr, 0);
    if constexpr (sizeof(T) == 4) {
      // rows with 1..80 non-zeros: row-per-thread (or tile) kernel; longer rows: generic kernel; empty rows: zero
      TRY(classify_rows(c, A));
      if (A.n_empty > 0) {
        zero_empty_rows_kernel<T><<<(unsigned)(((long long)A.n_cols * k + 255) / 256), 256, 0, c.stream>>>(P.ptr, A.n_cols, k, Y);
        LAUNCHED(); CU(cudaGetLastError());
      }
      if (A.n_short > 0) {
        P.row_list = A.all_short ? nullptr : A.short_list.i32();
        P.n_list = A.n_short;
        // default (and kernel = 4): row-per-thread panel kernel (als_chol_rows.cuh), measured 2.0x (rank 64) / 1.6x
        // (rank 128) faster than its predecessor, the 16 x 16 register-block kernel, which stays selectable as kernel = 5
        const bool rows_kernel = (o.kernel != 5);
        // persistent CTAs: exactly as many as are co-resident (registers AND shared memory), else a second wave
        int per_sm = 1, grid = 1;
        auto launch = [&](auto kern, int threads, size_t smem) -> cudaError_t {
          cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return e;
          e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
          if (e != cudaSuccess) return e;
          grid = std::min(c.sm_count * std::max(1, per_sm), A.n_short);
          kern<<<grid, threads, smem, c.stream>>>(P);
          return cudaSuccess;
        };
        if (rows_kernel) {
          if (k == 64 && o.kernel == 9) CU(launch(als_chol_warp64_kernel, 32, sizeof(CholRowsSmem<64>)));   // warp per system (experimental, not yet run on a GPU)
          else if (k == 64) CU(launch(als_chol_rows_kernel<64, 8>, 64, sizeof(CholRowsSmem<64>)));
          else if (o.kernel == 6) CU(launch(als_chol_rows_kernel<128, 3, 1>, 128, sizeof(CholRowsSmem<128>)));   // tcgen05 Gram, single-buffered (experimental)
          else if (o.kernel == 8) CU(launch(als_chol_rows_split_kernel, kSplitThreads, sizeof(CholRowsSmem<128>)));   // split rows (experimental, not yet run on a GPU)
          else if (o.kernel == 7) CU(launch(als_chol_rows_kernel<128, 3, 2>, 128, sizeof(CholRowsSmem<128>)));   // tcgen05 Gram, pipelined (experimental, not yet run on a GPU)
          else if (o.ctas == 2) CU(launch(als_chol_rows_kernel<128, 2>, 128, sizeof(CholRowsSmem<128>)));
          else CU(launch(als_chol_rows_kernel<128, 3>, 128, sizeof(CholRowsSmem<128>)));   // measured: 134.5 vs 171.2 ms / 1 M rows
        } else {
          if (k == 64) CU(launch(als_chol_tile_kernel<64>, kCholThreads, sizeof(CholTileSmem<64>)));
          else CU(launch(als_chol_tile_kernel<128>, kCholThreads, sizeof(CholTileSmem<128>)));
        }
        LAUNCHED(); CU(cudaGetLastError());
        sum_partials_kernel<<<1, 32, 0, c.stream>>>(P.loss_partials, grid, c.loss_acc.f64(), 1);
        LAUNCHED(); CU(cudaGetLastError());
      }
      if (A.n_long > 0) TRY(run_generic_chol(A.long_list.i32(), A.n_long));
    }
    return B200ALS_OK;
  }

  // ---- conjugate gradient ----
  bool resident = false;
  if constexpr (sizeof(T) == 4) {
    resident = (k == kResK) && (o.kernel != 1) && (o.feedback == B200ALS_EXPLICIT || G || diag) && !biased;
    if (o.kernel == 2 && !resident) return fail(B200ALS_EUNSUPPORTED, "resident kernel requires rank 128 fp32");
  }
  if (!resident) {
    if (diag && !G) return fail(B200ALS_EINVAL, "generic CG needs the full XtX");
    return run_generic_cg(nullptr, 0);
  }
  if constexpr (sizeof(T) == 4) {
    TRY(classify_rows(c, A));
    if (sub_range && !A.all_short) return fail(B200ALS_EINVAL, "row sub-ranges need a block without empty or long rows");
    if (A.n_empty > 0) {
      zero_empty_rows_kernel<T><<<(unsigned)(((long long)A.n_cols * k + 255) / 256), 256, 0, c.stream>>>(P.ptr, A.n_cols, k, Y);
      LAUNCHED(); CU(cudaGetLastError());
    }
    if (A.n_short > 0) {
      ResidentParams R;
      R.ptr = P.ptr;
      R.idx = P.idx;
      R.val = (const float*)P.val;
      R.X = (const float*)X;
      R.Y = (float*)Y;
      R.diag = diag;
      R.G = (const float*)G;
      R.feedback = o.feedback;
      R.cg_steps = o.cg_steps;
      R.dynamic_lambda = o.dynamic_lambda;
      R.lambda = (float)o.lambda;
      R.row_list = A.all_short ? nullptr : A.short_list.i32();
      R.n_list = sub_range ? n_rows_here : A.n_short;
      R.n_list_dev = nullptr;
      R.ptr_base = 0;
      R.row_begin = sub_range ? o.row_begin : 0;
      R.loss_partials = P.loss_partials;
      const int ctas = (o.ctas == 3 || o.ctas == 4) ? o.ctas : kDefaultCtas;
      const int grid = std::min(c.sm_count * ctas, R.n_list);
      const size_t smem = sizeof(ResidentSmem);
      const bool full_g = (o.feedback == B200ALS_IMPLICIT) && !diag;
      auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, kResThreads, smem, c.stream>>>(R);
        return cudaSuccess;
      };
      const int stage = (o.stage == 1) ? 0 : (o.stage == 2 ? 1 : kDefaultStage);
      if (full_g) {
        if (ctas == 4) CU(stage == 0 ? launch(als_cg_resident_kernel<true, 0, 4>) : launch(als_cg_resident_kernel<true, 1, 4>));
        else CU(stage == 0 ? launch(als_cg_resident_kernel<true, 0, 3>) : launch(als_cg_resident_kernel<true, 1, 3>));
      } else {
        if (ctas == 4) CU(stage == 0 ? launch(als_cg_resident_kernel<false, 0, 4>) : launch(als_cg_resident_kernel<false, 1, 4>));
        else CU(stage == 0 ? launch(als_cg_resident_kernel<false, 0, 3>) : launch(als_cg_resident_kernel<false, 1, 3>));
      }
      LAUNCHED(); CU(cudaGetLastError());
      sum_partials_kernel<<<1, 32, 0, c.stream>>>(P.loss_partials, grid, c.loss_acc.f64(), 1);
      LAUNCHED(); CU(cudaGetLastError());
    }
    if (A.n_long > 0) {
      if (diag && !G) return fail(B200ALS_EINVAL, "rows longer than 80 need the full XtX for the streaming kernel");
      TRY(run_generic_cg(A.long_list.i32(), A.n_long));
    }
  }
  return B200ALS_OK;
}

// loss = (sum_rows + lambda * regulariser) / nnz, rounded through T like the reference's return type
// (wrmf_implicit.hpp:286-304, wrmf_explicit.hpp:147-173)
template <typename T>
static int finish_loss(Ctx& c, const T* X, int k, long long n_src, const T* cnt_X, const HalfOpts& o, int64_t nnz,
                       double rows_sum, bool rows_sum_given, double* loss_out) {
  double reg = 0.0;
  if (o.lambda > 0) {
    const bool weighted = (o.feedback == B200ALS_EXPLICIT) && o.dynamic_lambda;
    const int grid = c.sm_count * 2;
    CU(c.reg_partials.ensure(sizeof(double) * (size_t)grid));
    if (o.reg_ld > 0)
      sqnorm_cols_kernel<T><<<grid, 256, 0, c.stream>>>(X, o.reg_ld, o.reg_lo, o.reg_hi, n_src, weighted ? cnt_X : nullptr,
                                                       c.reg_partials.f64());
    else
      sqnorm_kernel<T><<<grid, 256, 0, c.stream>>>(X, k, n_src, weighted ? cnt_X : nullptr, c.reg_partials.f64());
    LAUNCHED(); CU(cudaGetLastError());
    sum_partials_kernel<<<1, 32, 0, c.stream>>>(c.reg_partials.f64(), grid, c.loss_acc.f64() + 1, 0);
    LAUNCHED(); CU(cudaGetLastError());
  }
  double h[2] = {0, 0};
  int st = 0;
  CU(cudaMemcpyAsync(h, c.loss_acc.p, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
  CU(cudaMemcpyAsync(&st, c.status.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  if (st != 0) return fail(B200ALS_ENOTSPD, "a per-row system was not positive definite (Cholesky pivot <= 0)");
  if (o.lambda > 0) reg = h[1];
  const double rows = rows_sum_given ? rows_sum : h[0];
  if (loss_out) *loss_out = (double)(T)((rows + o.lambda * reg) / (double)nnz);
  return B200ALS_OK;
}
