// engine_context.inl -- part of engine.cu (included there; not a standalone translation unit).
// ------------------------------------------------------------------------------------------------------
// device context shared by the stateless calls and sessions
// ------------------------------------------------------------------------------------------------------
struct Ctx {
  int device = -1;
  int sm_count = 148;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;
  DevBuf ticket, loss_partials, loss_acc, status, gram_partials, reg_partials, rot_rt;
  bool attrs_set = false;
  int init() {
    if (stream) return B200ALS_OK;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
      return fail(B200ALS_ECUDA, "no CUDA device visible (this engine has no CPU fallback)");
    CU(cudaGetDevice(&device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    sm_count = prop.multiProcessorCount;
    smem_optin = prop.sharedMemPerBlockOptin;
    CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CU(ticket.ensure(sizeof(unsigned long long)));
    CU(loss_acc.ensure(4 * sizeof(double)));
    CU(status.ensure(sizeof(int)));
    return B200ALS_OK;
  }
};
static Ctx& ctx() {
  static thread_local Ctx c;
  return c;
}

template <typename T>
struct CscDev {
  int32_t n_rows = 0, n_cols = 0;
  int64_t nnz = 0;
  DevBuf ptr, idx, val;
  // row classes for the resident kernel (built lazily)
  DevBuf short_list, long_list;
  int n_short = -1, n_long = 0, n_empty = 0;
  bool all_short = false;
  // row-length classes of the CG path (built lazily per rank by plan_rows, engine_solve.inl)
  struct RowClass {
    int lo = 0, hi = 0;       // rows with lo <= nnz <= hi
    int cap = 0, warps = 0;   // tile kernel launch shape (0: not a tile class)
    int cluster = 1;          // CTAs per thread-block cluster (rows of up to cluster * cap entries)
    int nbuf = 2;             // tile buffers per CTA
    bool stream = false;      // this range of row lengths goes to the streaming kernel (a cluster size switched off)
    int count = 0;
    DevBuf list;              // ascending row ids (stable compaction => deterministic launch order)
  };
  // tile classes: 4 warps x 4 CTAs/SM double- / single-buffered, 8 warps x 2 CTAs/SM single-buffered, 16 warps x 1 CTA/SM
  // double-buffered, then thread-block clusters of 2 / 4 / 8 CTAs
  static constexpr int kClsResident = 0, kClsTile0 = 1, kNumTile = 7, kClsLong = 8, kNumCls = 9;
  RowClass cls[kNumCls];
  int all_ge1 = -1;           // implicit confidences: 1 = every value >= 1 (the symmetric Gram-rows kernel applies), 0 = not, -1 = unknown
  int plan_key = -1;          // 0: a plan exists (built for plan_sig), -1: none
  int plan_sig[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int plan_empty = 0;
  int plan_single = -1;       // >= 0: every row of the block lies in this one class (chunked solves allowed)
};

template <typename T>
static int upload_csc(const b200als_csc* A, CscDev<T>& D, cudaStream_t st) {
  if (!A || !A->ptr || (A->nnz > 0 && (!A->idx || (!A->val_f64 && !A->val_f32))))
    return fail(B200ALS_EINVAL, "b200als_csc: null ptr/idx/val");
  if (A->n_cols < 0 || A->n_rows < 0 || A->nnz < 0) return fail(B200ALS_EINVAL, "b200als_csc: negative size");
  D.n_rows = A->n_rows;
  D.n_cols = A->n_cols;
  D.nnz = A->nnz;
  CU(D.ptr.ensure(sizeof(int32_t) * ((size_t)A->n_cols + 1)));
  CU(D.idx.ensure(sizeof(int32_t) * (size_t)A->nnz));
  CU(D.val.ensure(sizeof(T) * (size_t)A->nnz));
  CU(cudaMemcpyAsync(D.ptr.p, A->ptr, sizeof(int32_t) * ((size_t)A->n_cols + 1), cudaMemcpyHostToDevice, st));
  if (A->nnz) {
    CU(cudaMemcpyAsync(D.idx.p, A->idx, sizeof(int32_t) * (size_t)A->nnz, cudaMemcpyHostToDevice, st));
    const bool same_f64 = A->val_f64 && sizeof(T) == 8, same_f32 = !A->val_f64 && sizeof(T) == 4;
    if (same_f64 || same_f32) {
      CU(cudaMemcpyAsync(D.val.p, A->val_f64 ? (const void*)A->val_f64 : (const void*)A->val_f32,
                         sizeof(T) * (size_t)A->nnz, cudaMemcpyHostToDevice, st));
    } else {
      // double -> float (or float -> double) once at upload; the reference converts per visit
      // (wrmf_implicit.hpp:182-183), same rounding
      DevBuf tmp;
      const size_t eb = A->val_f64 ? 8 : 4;
      CU(tmp.ensure(eb * (size_t)A->nnz));
      CU(cudaMemcpyAsync(tmp.p, A->val_f64 ? (const void*)A->val_f64 : (const void*)A->val_f32, eb * (size_t)A->nnz,
                         cudaMemcpyHostToDevice, st));
      const int bs = 256;
      const unsigned gs = (unsigned)((A->nnz + bs - 1) / bs);
      if (A->val_f64) convert_kernel<double, T><<<gs, bs, 0, st>>>(tmp.f64(), D.val.template as<T>(), A->nnz);
      else convert_kernel<float, T><<<gs, bs, 0, st>>>(tmp.f32(), D.val.template as<T>(), A->nnz);
      LAUNCHED(); CU(cudaGetLastError());
      CU(cudaStreamSynchronize(st));
    }
  }
  D.n_short = -1;
  D.plan_key = -1;
  D.all_ge1 = -1;
  return B200ALS_OK;
}

// ------------------------------------------------------------------------------------------------------
// Gram
// ------------------------------------------------------------------------------------------------------
// Gram arithmetic: 0 = default (tcgen05 3xTF32 at rank 128 / 256 with >= 8192 rows, else fp32 FMA), 1 = tcgen05 with bf16
// operands (BASELINE configs[4]: "fp32 vs tensor-core bf16 Gram"), 2 = fp32 FMA everywhere.  The session option
// (b200als_options.reserved[2]) wins over the environment (B200ALS_GRAM = tf32x3 | bf16 | ffma).
static int g_gram_mode_override = -1;   // set around a session's half-iteration
static int gram_mode() {
  if (g_gram_mode_override >= 0) return g_gram_mode_override;
  const char* e = getenv("B200ALS_GRAM");
  if (!e) return 0;
  if (e[0] == 'f' || e[0] == 'F') return 2;
  if (e[0] == 'b' || e[0] == 'B') return 1;
  return 0;
}
template <typename T>
static int run_gram(Ctx& c, const T* X, int k, long long n, double lambda, T* G, double* G64) {
  if (n <= 0) {   // empty matrix (e.g. an empty Gram slice of a multi-GPU run): G = lambda I, no launch geometry to derive
    CU(c.gram_partials.ensure(sizeof(double)));
    gram_reduce_kernel<T><<<(k * k + 255) / 256, 256, 0, c.stream>>>(c.gram_partials.f64(), 0, 1, k, lambda, G, G64);
    LAUNCHED(); CU(cudaGetLastError());
    return B200ALS_OK;
  }
  if constexpr (sizeof(T) == 4) {
    const int mode = gram_mode();
    if ((k == kTcK || k == 2 * kTcK) && n >= 8192 && mode != 2) {   // small inputs: the exact fp32 FMA kernel
      const int nt1 = k / kTcK, n_tiles = nt1 * (nt1 + 1) / 2;
      const long long target = std::max(1, 888 / n_tiles);            // ~6 CTAs per SM in flight over all tiles
      long long rows_per = std::max<long long>(1024, (n + target - 1) / target);
      rows_per = ((rows_per + 255) / 256) * 256;   // whole drain windows
      const long long n_cta = (n + rows_per - 1) / rows_per;
      CU(c.gram_partials.ensure(sizeof(double) * (size_t)n_cta * n_tiles * kTcK * kTcK));
      const size_t smem = sizeof(GramTcSmem2);
      const dim3 grid((unsigned)n_cta, (unsigned)n_tiles);
      if (mode == 1) {
        CU(cudaFuncSetAttribute(gram_tc_blocks_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gram_tc_blocks_kernel<true><<<grid, 128, smem, c.stream>>>((const float*)X, k, n, rows_per, c.gram_partials.f64());
      } else {
        CU(cudaFuncSetAttribute(gram_tc_blocks_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gram_tc_blocks_kernel<false><<<grid, 128, smem, c.stream>>>((const float*)X, k, n, rows_per, c.gram_partials.f64());
      }
      LAUNCHED(); CU(cudaGetLastError());
      gram_reduce_kernel<T><<<(k * k + 255) / 256, 256, 0, c.stream>>>(c.gram_partials.f64(), (int)n_cta, n_tiles, k, lambda, G, G64);
      LAUNCHED(); CU(cudaGetLastError());
      return B200ALS_OK;
    }
  }
  const int nt1 = (k + kGramTile - 1) / kGramTile, n_tiles = nt1 * (nt1 + 1) / 2;
  long long n_cta = std::min<long long>((long long)c.sm_count * 2 / std::max(1, n_tiles) + 1, (n + 255) / 256);
  n_cta = std::max<long long>(1, n_cta);
  long long rows_per = (n + n_cta - 1) / n_cta;
  rows_per = ((rows_per + kGramRows - 1) / kGramRows) * kGramRows;
  n_cta = std::max<long long>(1, (n + rows_per - 1) / rows_per);
  CU(c.gram_partials.ensure(sizeof(double) * (size_t)n_cta * n_tiles * kGramTile * kGramTile));
  gram_partial_kernel<T><<<dim3((unsigned)n_cta, (unsigned)n_tiles), 256, 0, c.stream>>>(X, k, n, rows_per,
                                                                                         c.gram_partials.f64(), nt1);
  LAUNCHED(); CU(cudaGetLastError());
  gram_reduce_kernel<T><<<(k * k + 255) / 256, 256, 0, c.stream>>>(c.gram_partials.f64(), (int)n_cta, n_tiles, k,
                                                                   lambda, G, G64);
  LAUNCHED(); CU(cudaGetLastError());
  return B200ALS_OK;
}
