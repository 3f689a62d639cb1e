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

// rows with 1..80 entries / longer rows / empty rows for the Cholesky kernels: stable compactions (ascending row ids),
// so the launch order and with it the loss summation order is the same on every run
struct LenInRange {
  const int32_t* ptr;
  int lo, hi;
  __host__ __device__ bool operator()(const int& r) const {
    const int n = ptr[r + 1] - ptr[r];
    return n >= lo && n <= hi;
  }
};
template <typename T>
static int classify_rows(Ctx& c, CscDev<T>& A) {
  if (A.n_short >= 0) return B200ALS_OK;
  CU(A.short_list.ensure(sizeof(int32_t) * (size_t)std::max(1, A.n_cols)));
  CU(A.long_list.ensure(sizeof(int32_t) * (size_t)std::max(1, A.n_cols)));
  DevBuf counts, temp;
  CU(counts.ensure(2 * sizeof(int)));
  CU(cudaMemsetAsync(counts.p, 0, 2 * sizeof(int), c.stream));
  if (A.n_cols > 0) {
    cub::CountingInputIterator<int> rows(0);
    size_t temp_bytes = 0;
    CU(cub::DeviceSelect::If(nullptr, temp_bytes, rows, (int32_t*)nullptr, (int*)nullptr, A.n_cols, LenInRange{nullptr, 0, 0}, c.stream));
    CU(temp.ensure(temp_bytes));
    CU(cub::DeviceSelect::If(temp.p, temp_bytes, rows, A.short_list.i32(), counts.i32(), A.n_cols,
                             LenInRange{A.ptr.i32(), 1, kCholMaxN}, c.stream));
    LAUNCHED();
    CU(cub::DeviceSelect::If(temp.p, temp_bytes, rows, A.long_list.i32(), counts.i32() + 1, A.n_cols,
                             LenInRange{A.ptr.i32(), kCholMaxN + 1, std::numeric_limits<int>::max()}, c.stream));
    LAUNCHED();
  }
  int h[2];
  CU(cudaMemcpyAsync(h, counts.p, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  A.n_short = h[0];
  A.n_long = h[1];
  A.n_empty = A.n_cols - h[0] - h[1];
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

// ---- row-length classes of the CG path ---------------------------------------------------------------------
// Rows are binned by their number of entries: [resident kernel: 1..80 at rank 128] / three tile-kernel classes sized so
// that 4, 2 or 1 CTAs of als_cg_tile_kernel share an SM / longer rows (streaming kernel) / empty rows (zeroed).
// Lists come from stable compactions (cub::DeviceSelect::If over a counting iterator): ascending row ids, so the
// launch order, hence the loss summation order, is the same on every run.
// rank 128: rows beyond the single-CTA tile classes are solved by als_cg_gram_kernel (per-row Gram on tcgen05 + CG on the
// explicit 128 x 128 system) instead of clusters / the streaming kernel.  B200ALS_GRAM_ROWS=0 switches it off,
// B200ALS_GRAM_ROWS_MIN=n lowers the row length from which it takes over (A/B runs).
static bool gram_rows_enabled(int k) {
  if (k != kTcK) return false;
  const char* e = getenv("B200ALS_GRAM_ROWS");
  return !(e && e[0] == '0');
}
// The Gram-rows kernel forms X_nnz diag(c - 1) X_nnz' as Z Z' with z_j = sqrt(c_j - 1) x_j: implicit confidences must all be
// >= 1 (checked once per uploaded matrix); explicit feedback has unit weights.
template <typename T>
static int gram_rows_for(Ctx& c, CscDev<T>& A, int k, int feedback, bool* ok) {
  *ok = false;
  if constexpr (sizeof(T) != 4) return B200ALS_OK;
  if (!gram_rows_enabled(k)) return B200ALS_OK;
  if (feedback == B200ALS_EXPLICIT) { *ok = true; return B200ALS_OK; }
  if (A.all_ge1 < 0) {
    DevBuf flag;
    CU(flag.ensure(sizeof(int)));
    CU(cudaMemsetAsync(flag.p, 0, sizeof(int), c.stream));
    if (A.nnz > 0) {
      any_below_one_kernel<<<c.sm_count * 4, 256, 0, c.stream>>>((const float*)A.val.p, (long long)A.nnz, flag.i32());
      LAUNCHED(); CU(cudaGetLastError());
    }
    int h = 0;
    CU(cudaMemcpyAsync(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    CU(cudaStreamSynchronize(c.stream));
    A.all_ge1 = h ? 0 : 1;
  }
  *ok = (A.all_ge1 == 1);
  return B200ALS_OK;
}
static int tile_kpad(int k) { return k <= 16 ? 16 : k <= 32 ? 32 : k <= 64 ? 64 : k <= 128 ? 128 : 256; }
static int tile_cap_for(int kpad, int warps, size_t budget, bool full_g, int nbuf, int cluster) {
  int cap = 0;
  for (int cnd = 4; cnd <= 8192; cnd += 4) {
    const TileCgLayout L{kpad, cnd, warps, full_g ? 1 : 0, nbuf, cluster > 1 ? 1 : 0};
    if (L.bytes() > budget) break;
    cap = cnd;
  }
  return cap;
}
template <typename T>
static int plan_rows(Ctx& c, CscDev<T>& A, int k, bool resident_ok, bool full_g, bool gram_rows) {
  // the plan depends on the rank, the mode and the A/B switches of the environment: cached until any of them changes
  auto env_int = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
  const int sig[8] = {k, resident_ok ? 1 : 0, full_g ? 1 : 0, env_int("B200ALS_TILE_WARPS_S", 4) * 100 + env_int("B200ALS_TILE_WARPS_M", 0), env_int("B200ALS_TILE_CLUSTER", -1),
                      env_int("B200ALS_TILE_CLUSTER_MIN", -1), gram_rows ? 1 : 0, env_int("B200ALS_GRAM_ROWS_MIN", -1)};
  if (A.plan_key == 0 && std::memcmp(sig, A.plan_sig, sizeof(sig)) == 0) return B200ALS_OK;
  using RC = typename CscDev<T>::RowClass;
  const int kpad = tile_kpad(k);
  int warpsL = 16;   // one CTA per SM: 16 warps (two-stage cross-warp sum); B200ALS_TILE_WARPS_L = 4 | 8 | 16 for A/B runs
  if (const char* e = getenv("B200ALS_TILE_WARPS_L")) { const int v = atoi(e); if (v == 4 || v == 8 || v == 16) warpsL = v; }
  int warpsS = 4, ctasS = 4;   // B200ALS_TILE_WARPS_S = 2 | 4 (A/B): 2 warps per CTA => up to 8 CTAs per SM in the small classes
  if (const char* e = getenv("B200ALS_TILE_WARPS_S")) { if (atoi(e) == 2) { warpsS = 2; ctasS = 8; } }
  const size_t sm_bytes = c.smem_optin + 1024;   // per-SM shared memory (the opt-in per-block limit + the 1 KB reserve)
  // {warps per CTA, CTAs per SM, CTAs per cluster, tile buffers}.  Measured (profiles/r2/tile_ab.txt): every warp of a
  // CTA repeats the CG vector algebra, so FEW warps per CTA and MANY CTAs per SM win -- rank 128, rows of 80: 4 warps x 4
  // CTAs with ONE tile buffer (the other CTAs cover a row's load) 17.4 ms per 1 M rows, 8 warps x 2 CTAs double-buffered
  // 24.4 ms; where both fit, the double-buffered form of the same shape is 4 % faster.  Hence, by increasing row length:
  // 4 warps x 4 CTAs double-buffered, the same single-buffered, 8 (rank 256: 4) warps x 2 CTAs single-buffered, 16 warps x
  // 1 CTA double-buffered (one CTA per SM needs the warps).
  // Clusters: off unless B200ALS_TILE_CLUSTER >= 2 -- on the heavy-tailed robustness point the streaming kernel beat them
  // (32.8 vs 40.6 ms); on 1 M uniform rows of 800 entries they won by 8 % (387 vs 421 ms).
  // warps of the 2-CTA/SM single-buffered class: 8, except at rank 256 where only 4 leave room for rows of 100 entries
  // (measured: C5 slice 360 ms vs 405 ms for the 16-warp 1-CTA class; rank 128 ragged rows 25.3 ms with 8 vs 26.9 ms with 4).
  // B200ALS_TILE_WARPS_M = 4 | 8 overrides.
  int warpsM = (kpad == 256) ? 4 : 8;
  if (const char* e = getenv("B200ALS_TILE_WARPS_M")) { const int v = atoi(e); if (v == 4 || v == 8) warpsM = v; }
  const int shape[7][4] = {{warpsS, ctasS, 1, 2}, {warpsS, ctasS, 1, 1}, {warpsM, 2, 1, 1}, {warpsL, 1, 1, 2}, {16, 1, 2, 2}, {16, 1, 4, 2}, {16, 1, 8, 2}};
  int max_cluster = 1;   // B200ALS_TILE_CLUSTER = 1 | 2 | 4 | 8: largest cluster used (1: longer rows go to the long-row kernels)
  if (const char* e = getenv("B200ALS_TILE_CLUSTER")) max_cluster = std::max(1, atoi(e));
  if (gram_rows) max_cluster = 1;   // rank 128: long rows go to the tensor-core Gram kernel, not to clusters
  int min_cluster = 2;   // B200ALS_TILE_CLUSTER_MIN = 2 | 4 | 8: smallest cluster used (rows between one CTA's capacity and
                         // min_cluster / 2 times it go to the streaming kernel)
  if (const char* e = getenv("B200ALS_TILE_CLUSTER_MIN")) min_cluster = std::max(2, atoi(e));
  int lo = 1;
  RC& R = A.cls[CscDev<T>::kClsResident];
  R.lo = 1; R.hi = resident_ok ? kResMaxN : 0; R.cap = 0; R.warps = 0;
  if (resident_ok) lo = kResMaxN + 1;
  for (int t = 0; t < CscDev<T>::kNumTile; t++) {
    RC& C = A.cls[CscDev<T>::kClsTile0 + t];
    C.warps = shape[t][0];
    C.cluster = shape[t][2];
    C.nbuf = shape[t][3];
    C.cap = tile_cap_for(kpad, C.warps, sm_bytes / shape[t][1] - 1024, full_g, C.nbuf, C.cluster);
    C.lo = lo;
    // a cluster of CL CTAs holds CL slabs of ceil(n / CL) <= cap entries each
    // a cluster size below B200ALS_TILE_CLUSTER_MIN keeps its range of row lengths but hands it to the streaming kernel
    C.stream = (C.cluster > 1 && C.cluster < min_cluster && C.cluster <= max_cluster);
    const bool enabled = (C.cluster == 1) || (C.cluster <= max_cluster);
    C.hi = enabled ? std::max(lo - 1, C.cap * C.cluster) : lo - 1;
    lo = C.hi + 1;
  }
  if (gram_rows) {
    if (const char* e = getenv("B200ALS_GRAM_ROWS_MIN")) {
      const int v = std::max(R.hi + 1, atoi(e));
      for (int t = 0; t < CscDev<T>::kNumTile; t++) {
        RC& C = A.cls[CscDev<T>::kClsTile0 + t];
        if (C.hi < C.lo) continue;
        if (C.lo >= v) C.hi = C.lo - 1;          // class switched off
        else if (C.hi >= v) C.hi = v - 1;
      }
      lo = std::min(lo, v);
    }
  }
  RC& Lg = A.cls[CscDev<T>::kClsLong];
  Lg.lo = lo; Lg.hi = std::numeric_limits<int>::max(); Lg.cap = 0; Lg.warps = 0;
  DevBuf counts, temp;
  CU(counts.ensure(sizeof(int) * CscDev<T>::kNumCls));
  CU(cudaMemsetAsync(counts.p, 0, sizeof(int) * CscDev<T>::kNumCls, c.stream));
  if (A.n_cols > 0) {
    cub::CountingInputIterator<int> rows(0);
    size_t temp_bytes = 0;
    CU(cub::DeviceSelect::If(nullptr, temp_bytes, rows, (int32_t*)nullptr, (int*)nullptr, A.n_cols, LenInRange{nullptr, 0, 0}, c.stream));
    CU(temp.ensure(temp_bytes));
    for (int q = 0; q < CscDev<T>::kNumCls; q++) {
      RC& C = A.cls[q];
      if (C.hi < C.lo) continue;
      CU(C.list.ensure(sizeof(int32_t) * (size_t)A.n_cols));
      CU(cub::DeviceSelect::If(temp.p, temp_bytes, rows, C.list.i32(), counts.i32() + q, A.n_cols,
                               LenInRange{A.ptr.i32(), C.lo, C.hi}, c.stream));
      LAUNCHED();
    }
  }
  int h[CscDev<T>::kNumCls];
  CU(cudaMemcpyAsync(h, counts.p, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  int total = 0;
  A.plan_single = -1;
  for (int q = 0; q < CscDev<T>::kNumCls; q++) {
    A.cls[q].count = (A.cls[q].hi < A.cls[q].lo) ? 0 : h[q];
    total += A.cls[q].count;
    if (A.cls[q].count == A.n_cols && A.n_cols > 0) A.plan_single = q;
  }
  A.plan_empty = A.n_cols - total;
  std::memcpy(A.plan_sig, sig, sizeof(sig));
  A.plan_key = 0;
  return B200ALS_OK;
}

// launches als_cg_tile_kernel for one length class; *grid_out = CTAs launched (loss partials to sum)
static int launch_cg_tile(Ctx& c, TileCgParams P, int warps, int cluster, bool full_g, int* grid_out) {
  const int kpad = tile_kpad(P.k);
  const TileCgLayout L{kpad, P.cap, warps, full_g ? 1 : 0, P.nbuf, cluster > 1 ? 1 : 0};
  const size_t smem = L.bytes();
  if (smem > c.smem_optin) return fail(B200ALS_EUNSUPPORTED, "tile kernel: row class does not fit shared memory");
  int per_sm = 1, grid = 1;
  const int threads = warps * 32;
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (e != cudaSuccess) return e;
    grid = std::min(c.sm_count * std::max(1, per_sm), P.n_list);
    kern<<<grid, threads, smem, c.stream>>>(P);
    return cudaSuccess;
  };
  // thread-block clusters: one row per cluster, as many clusters as the device can hold at once
  auto launch_cluster = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)cluster, 1, 1);
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c.stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n_clusters = 0;
    e = cudaOccupancyMaxActiveClusters(&n_clusters, kern, &cfg);
    if (e != cudaSuccess) return e;
    if (n_clusters < 1) return cudaErrorLaunchOutOfResources;
    n_clusters = std::min(n_clusters, P.n_list);
    grid = n_clusters * cluster;
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    return cudaLaunchKernelEx(&cfg, kern, P);
  };
#define B200ALS_TILE_CASE(LPR, CC)                                                                                    \
  do {                                                                                                                \
    if (cluster > 1) CU(full_g ? launch_cluster(als_cg_tile_kernel<LPR, CC, true, true>)                              \
                               : launch_cluster(als_cg_tile_kernel<LPR, CC, false, true>));                           \
    else CU(full_g ? launch(als_cg_tile_kernel<LPR, CC, true, false>) : launch(als_cg_tile_kernel<LPR, CC, false, false>)); \
  } while (0)
  switch (kpad) {
    case 16: B200ALS_TILE_CASE(4, 1); break;
    case 32: B200ALS_TILE_CASE(8, 1); break;
    case 64: B200ALS_TILE_CASE(16, 1); break;
    case 128: B200ALS_TILE_CASE(32, 1); break;
    default: B200ALS_TILE_CASE(32, 2); break;
  }
#undef B200ALS_TILE_CASE
  LAUNCHED(); CU(cudaGetLastError());
  *grid_out = grid;
  return B200ALS_OK;
}

// launches als_cg_gram_kernel (rank 128, long rows); two CTAs per SM (the occupancy query does not see tensor memory)
static int launch_cg_gram(Ctx& c, TileCgParams P, int* grid_out) {
  const size_t smem = sizeof(GramCgSmem);
  CU(cudaFuncSetAttribute(als_cg_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 2;
  if (const char* e = getenv("B200ALS_GRAM_ROWS_PER_SM")) per_sm = std::max(1, std::min(2, atoi(e)));
  const int grid = std::min(c.sm_count * per_sm, std::max(1, P.n_list));
  als_cg_gram_kernel<<<grid, kGcThreads, smem, c.stream>>>(P);
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
  P.diag = nullptr;
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
  const int max_grid = c.sm_count * 32;
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
      // rows with 1..80 non-zeros: row-per-thread / warp-per-system kernels; longer rows: generic kernel; empty rows: zero
      TRY(classify_rows(c, A));
      if (A.n_empty > 0) {
        zero_empty_rows_kernel<T><<<(unsigned)(((long long)A.n_cols * k + 255) / 256), 256, 0, c.stream>>>(P.ptr, A.n_cols, k, Y);
        LAUNCHED(); CU(cudaGetLastError());
      }
      if (A.n_short > 0) {
        P.row_list = A.all_short ? nullptr : A.short_list.i32();
        P.n_list = A.n_short;
        // Kernel choice (measured round 2, profiles/r2/): rank 128 -- row-per-thread panel Cholesky with the per-row Gram on
        // tcgen05 (3xTF32 into TMEM), 111 ms per 1 M rows of 80 entries vs 134 ms for the FFMA2 Gram (kernel = 4);
        // rank 64 -- warp per system (kernel = 9), 22.8 ms vs 25.2 ms per 1 M rows of 50 for the two-warp CTA (kernel = 4).
        // persistent CTAs: exactly as many as are co-resident (registers AND shared memory), else a second wave
        int per_sm = 1, grid = 1;
        const bool tc_gram = (k == 128 && o.kernel != 4);
        const int min_per_sm = tc_gram ? 3 : 1;
        auto launch = [&](auto kern, int threads, size_t smem) -> cudaError_t {
          cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return e;
          e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
          if (e != cudaSuccess) return e;
          // the occupancy query answers 1 CTA/SM for the kernel that allocates tensor memory although 3 are co-resident
          // (registers, shared memory and 3 x 128 of the 512 TMEM columns all fit): with the grid it suggested the kernel
          // ran 2.1x slower (profiles/r2/chol_tc_occupancy.txt)
          per_sm = std::max(per_sm, min_per_sm);
          if (const char* ev = getenv("B200ALS_CHOL_PER_SM")) per_sm = std::max(1, atoi(ev));
          grid = std::min(c.sm_count * std::max(1, per_sm), A.n_short);
          kern<<<grid, threads, smem, c.stream>>>(P);
          return cudaSuccess;
        };
        if (k == 64 && o.kernel != 4) CU(launch(als_chol_warp64_kernel, 32, sizeof(CholRowsSmem<64>)));
        else if (k == 64) CU(launch(als_chol_rows_kernel<64, 8>, 64, sizeof(CholRowsSmem<64>)));
        else if (tc_gram) CU(launch(als_chol_rows_kernel<128, 3, 1>, 128, sizeof(CholRowsSmem<128>)));
        else if (o.ctas == 2) CU(launch(als_chol_rows_kernel<128, 2>, 128, sizeof(CholRowsSmem<128>)));
        else CU(launch(als_chol_rows_kernel<128, 3>, 128, sizeof(CholRowsSmem<128>)));
        LAUNCHED(); CU(cudaGetLastError());
        sum_partials_kernel<<<1, 32, 0, c.stream>>>(P.loss_partials, grid, c.loss_acc.f64(), 1);
        LAUNCHED(); CU(cudaGetLastError());
      }
      if (A.n_long > 0) TRY(run_generic_chol(A.long_list.i32(), A.n_long));
    }
    return B200ALS_OK;
  }

  // ---- conjugate gradient ----
  // fp32, rank % 4 == 0, rank <= 256, no bias terms: rows are binned by length (plan_rows) -- the register-resident kernel
  // takes rows of <= 80 entries at rank 128, the shared-memory tile kernel every other row that fits its buffers, the
  // streaming kernel the rest.  kernel = 1 forces the streaming kernel (the reference's arithmetic, full XtX),
  // kernel = 10 keeps the resident kernel out (tests / A-B runs of the tile kernel).
  bool tile_ok = false;
  if constexpr (sizeof(T) == 4) {
    tile_ok = (k % 4 == 0) && (k <= 256) && (o.kernel != 1) && (o.feedback == B200ALS_EXPLICIT || G || diag) && !biased;
    if (o.kernel == 2 && !(tile_ok && k == kResK)) return fail(B200ALS_EUNSUPPORTED, "resident kernel requires rank 128 fp32");
  }
  if (!tile_ok) {
    if (diag && !G) {
      if constexpr (sizeof(T) == 4) P.diag = (const T*)diag;
      else return fail(B200ALS_EINVAL, "the eigenbasis path is fp32 only");
    }
    return run_generic_cg(nullptr, 0);
  }
  if constexpr (sizeof(T) == 4) {
    using CD = CscDev<T>;
    const bool resident_ok = (k == kResK) && (o.kernel != 10);
    const bool full_g = (o.feedback == B200ALS_IMPLICIT) && !diag;
    bool gram_rows = false;
    TRY(gram_rows_for(c, A, k, o.feedback, &gram_rows));
    TRY(plan_rows(c, A, k, resident_ok, full_g, gram_rows));
    if (sub_range && (A.plan_single < 0 || (A.plan_single == CD::kClsLong && !gram_rows) || A.cls[A.plan_single].stream))
      return fail(B200ALS_EINVAL, "row sub-ranges need a block whose rows all fall into one length class");
    if (A.plan_empty > 0) {
      zero_empty_rows_kernel<T><<<(unsigned)(((long long)A.n_cols * k + 255) / 256), 256, 0, c.stream>>>(P.ptr, A.n_cols, k, Y);
      LAUNCHED(); CU(cudaGetLastError());
    }
    const typename CD::RowClass& RC = A.cls[CD::kClsResident];
    if (RC.count > 0) {
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
      R.row_list = (A.plan_single == CD::kClsResident) ? nullptr : RC.list.i32();
      R.n_list = sub_range ? n_rows_here : RC.count;
      R.n_list_dev = nullptr;
      R.ptr_base = 0;
      R.row_begin = sub_range ? o.row_begin : 0;
      R.loss_partials = P.loss_partials;
      const int ctas = (o.ctas == 3 || o.ctas == 4) ? o.ctas : kDefaultCtas;
      const int grid = std::min(c.sm_count * ctas, R.n_list);
      const size_t smem = sizeof(ResidentSmem);
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
    for (int t = 0; t < CD::kNumTile; t++) {
      const int q = CD::kClsTile0 + t;
      const typename CD::RowClass& TC = A.cls[q];
      if (TC.count == 0) continue;
      if (TC.stream) {
        if (diag && !G) P.diag = (const T*)diag;
        TRY(run_generic_cg(TC.list.i32(), TC.count));
        continue;
      }
      TileCgParams TP;
      TP.ptr = P.ptr;
      TP.idx = P.idx;
      TP.val = (const float*)P.val;
      TP.X = (const float*)X;
      TP.Y = (float*)Y;
      TP.diag = diag;
      TP.G = (const float*)G;
      TP.k = k;
      TP.feedback = o.feedback;
      TP.cg_steps = o.cg_steps;
      TP.dynamic_lambda = o.dynamic_lambda;
      TP.lambda = (float)o.lambda;
      TP.row_list = (A.plan_single == q) ? nullptr : TC.list.i32();
      TP.n_list = sub_range ? n_rows_here : TC.count;
      TP.ptr_base = 0;
      TP.row_begin = sub_range ? o.row_begin : 0;
      TP.cap = TC.cap;
      TP.nbuf = TC.nbuf;
      TP.loss_partials = P.loss_partials;
      int grid = 0;
      TRY(launch_cg_tile(c, TP, TC.warps, TC.cluster, full_g, &grid));
      sum_partials_kernel<<<1, 32, 0, c.stream>>>(P.loss_partials, grid, c.loss_acc.f64(), 1);
      LAUNCHED(); CU(cudaGetLastError());
    }
    const typename CD::RowClass& LC = A.cls[CD::kClsLong];
    if (LC.count > 0 && gram_rows) {
      TileCgParams TP;
      TP.ptr = P.ptr;
      TP.idx = P.idx;
      TP.val = (const float*)P.val;
      TP.X = (const float*)X;
      TP.Y = (float*)Y;
      TP.diag = diag;
      TP.G = (const float*)G;
      TP.k = k;
      TP.feedback = o.feedback;
      TP.cg_steps = o.cg_steps;
      TP.dynamic_lambda = o.dynamic_lambda;
      TP.lambda = (float)o.lambda;
      TP.row_list = (A.plan_single == CD::kClsLong) ? nullptr : LC.list.i32();
      TP.n_list = sub_range ? n_rows_here : LC.count;
      TP.ptr_base = 0;
      TP.row_begin = sub_range ? o.row_begin : 0;
      TP.cap = 0;
      TP.nbuf = 1;
      TP.loss_partials = P.loss_partials;
      int grid = 0;
      TRY(launch_cg_gram(c, TP, &grid));
      sum_partials_kernel<<<1, 32, 0, c.stream>>>(P.loss_partials, grid, c.loss_acc.f64(), 1);
      LAUNCHED(); CU(cudaGetLastError());
    } else if (LC.count > 0) {
      if (diag && !G) P.diag = (const T*)diag;
      TRY(run_generic_cg(LC.list.i32(), LC.count));
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
