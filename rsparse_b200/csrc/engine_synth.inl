// engine_synth.inl -- part of engine.cu (included there; not a standalone translation unit).
// ------------------------------------------------------------------------------------------------------
// 4. synthetic workloads
// ------------------------------------------------------------------------------------------------------
extern "C" int b200als_synth_csr_host(int32_t n_rows, int32_t n_cols, int32_t nnz_per_row, uint64_t seed, int explicit_values,
                                      int64_t row_offset, int32_t* ptr, int32_t* idx, float* val_f32, double* val_f64) {
  if (n_rows < 0 || n_cols <= 0 || nnz_per_row <= 0 || nnz_per_row > n_cols || !ptr || !idx)
    return fail(B200ALS_EINVAL, "bad synthetic shape");
  if ((long long)n_rows * nnz_per_row > 2147483647LL) return fail(B200ALS_EINVAL, "nnz exceeds 32-bit row pointers");
  const unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([=]() {
      const int64_t r0 = (int64_t)n_rows * t / nt, r1 = (int64_t)n_rows * (t + 1) / nt;
      for (int64_t r = r0; r < r1; r++) {
        ptr[r] = (int32_t)(r * nnz_per_row);
        for (int j = 0; j < nnz_per_row; j++) {
          int32_t col; float v;
          synth_entry(r + row_offset, j, n_cols, nnz_per_row, seed, explicit_values, &col, &v);
          const int64_t e = r * nnz_per_row + j;
          idx[e] = col;
          if (val_f32) val_f32[e] = v;
          if (val_f64) val_f64[e] = (double)v;
        }
      }
    });
  for (auto& x : th) x.join();
  ptr[n_rows] = (int32_t)((int64_t)n_rows * nnz_per_row);
  return B200ALS_OK;
}

extern "C" int b200als_create_synthetic(b200als_session** out, int32_t n_user_local, int64_t user_offset, int32_t n_user_global,
                                        int32_t n_item, int32_t nnz_per_row, uint64_t seed, int rank,
                                        const b200als_options* opts) {
  Ctx& c = ctx();
  TRY(c.init());
  if (!out || rank <= 0 || n_user_local < 0 || n_item <= 0 || nnz_per_row <= 0 || nnz_per_row > n_item)
    return fail(B200ALS_EINVAL, "bad argument");
  if ((long long)n_user_local * nnz_per_row > 2147483647LL) return fail(B200ALS_EINVAL, "local nnz exceeds 32-bit row pointers");
  b200als_session* s = new b200als_session();
  if (opts) s->opt = *opts; else b200als_default_options(&s->opt);
  s->k = rank;
  s->n_user = n_user_global;
  s->n_item = n_item;
  int rc = session_alloc(s);
  if (rc != B200ALS_OK) { b200als_destroy(s); return rc; }
  CscDev<float>& A = s->csc[B200ALS_USERS];
  A.n_rows = n_item;
  A.n_cols = n_user_local;
  A.nnz = (int64_t)n_user_local * nnz_per_row;
  auto bail = [&](int code, const char* m) { b200als_destroy(s); return fail(code, m); };
  if (A.ptr.ensure(sizeof(int32_t) * ((size_t)n_user_local + 1)) != cudaSuccess) return bail(B200ALS_ECUDA, "alloc ptr");
  if (A.idx.ensure(sizeof(int32_t) * (size_t)A.nnz) != cudaSuccess) return bail(B200ALS_ECUDA, "alloc idx");
  if (A.val.ensure(sizeof(float) * (size_t)A.nnz) != cudaSuccess) return bail(B200ALS_ECUDA, "alloc val");
  const long long total = std::max<long long>(A.nnz, n_user_local + 1);
  synth_csr_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c.stream>>>(n_user_local, n_item, nnz_per_row, seed,
                                                                         s->opt.feedback == B200ALS_EXPLICIT, user_offset,
                                                                         A.ptr.i32(), A.idx.i32(), A.val.f32());
  LAUNCHED(); if (cudaGetLastError() != cudaSuccess) return bail(B200ALS_ECUDA, "synth kernel launch");
  s->has[B200ALS_USERS] = true;
  s->shard_begin[B200ALS_USERS] = (int32_t)user_offset;
  s->shard_end[B200ALS_USERS] = (int32_t)user_offset + n_user_local;
  s->nnz_global[B200ALS_USERS] = (int64_t)n_user_global * nnz_per_row;
  rc = session_counts(s);
  if (rc == B200ALS_OK && cudaStreamSynchronize(c.stream) != cudaSuccess) rc = fail(B200ALS_ECUDA, "sync");
  if (rc != B200ALS_OK) { b200als_destroy(s); return rc; }
  *out = s;
  return B200ALS_OK;
}

// ---- skewed synthetic data (robustness points of bench.py) -------------------------------------------------------------
__device__ __forceinline__ int synth_row_len(int64_t row, int32_t mean, int32_t n_cols, uint64_t seed, int len_dist) {
  if (len_dist == 0) return mean;
  const uint64_t h1 = synth_hash(seed * 0x9E3779B1ull + 0x51ED27ull + (uint64_t)row * 2), h2 = synth_hash(seed * 0x9E3779B1ull + 0x51ED27ull + (uint64_t)row * 2 + 1);
  const float u1 = ((float)((h1 >> 40) & 0xFFFFFF) + 1.0f) / 16777217.0f;
  const float u2 = (float)((h2 >> 40) & 0xFFFFFF) / 16777216.0f;
  const float z = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
  const float v = (float)mean * expf(z - 0.5f);   // log-normal, sigma = 1, mean = `mean`
  return max(1, min(n_cols, (int)lrintf(v)));
}
__global__ void synth_len_kernel(int32_t n_rows, int32_t n_cols, int32_t mean, uint64_t seed, int len_dist, int64_t row_offset,
                                 int32_t* __restrict__ len) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rows) len[r] = synth_row_len(r + row_offset, mean, n_cols, seed, len_dist);
}
// one thread per row: entries in ascending order; Zipf ids are bumped to stay distinct (max(candidate, previous + 1)) and
// clamped so that the rest of the row still fits below n_cols
__global__ void synth_fill_kernel(int32_t n_rows, int32_t n_cols, uint64_t seed, int explicit_values, int col_dist,
                                  int64_t row_offset, const int32_t* __restrict__ ptr, int32_t* __restrict__ idx,
                                  float* __restrict__ val) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int p = ptr[r], n = ptr[r + 1] - p;
  const int64_t row = r + row_offset;
  const float logn = logf((float)n_cols);
  int prev = -1;
  for (int j = 0; j < n; j++) {
    const uint64_t h = synth_hash(seed * 0x100000001B3ull + (uint64_t)row * 0x10001ull + (uint64_t)j);
    int col;
    if (col_dist == 0) {
      const int64_t lo = ((int64_t)j * n_cols) / n, hi = ((int64_t)(j + 1) * n_cols) / n;
      col = (int)(lo + (int64_t)(h % (uint64_t)(hi - lo)));
    } else {
      const float u = ((float)j + (float)(h & 0xFFFF) / 65536.0f) / (float)n;   // stratified in log space
      col = (int)floorf(expf(u * logn)) - 1;
      col = max(col, prev + 1);
      col = min(col, n_cols - (n - j));
    }
    prev = col;
    idx[p + j] = col;
    const float uv = (float)((h >> 40) & 0xFFFFFF) / 16777216.0f;
    val[p + j] = explicit_values ? (1.0f + floorf(uv * 5.0f)) : (1.0f + floorf(10.0f * uv * uv));
  }
}

extern "C" int b200als_create_synthetic_ex(b200als_session** out, int32_t n_user_local, int64_t user_offset, int32_t n_user_global,
                                           int32_t n_item, int32_t nnz_per_row, uint64_t seed, int rank,
                                           const b200als_options* opts, int col_dist, int len_dist) {
  if (col_dist == 0 && len_dist == 0)
    return b200als_create_synthetic(out, n_user_local, user_offset, n_user_global, n_item, nnz_per_row, seed, rank, opts);
  Ctx& c = ctx();
  TRY(c.init());
  if (!out || rank <= 0 || n_user_local < 0 || n_item <= 0 || nnz_per_row <= 0 || nnz_per_row > n_item)
    return fail(B200ALS_EINVAL, "bad argument");
  b200als_session* s = new b200als_session();
  if (opts) s->opt = *opts; else b200als_default_options(&s->opt);
  s->k = rank;
  s->n_user = n_user_global;
  s->n_item = n_item;
  int rc = session_alloc(s);
  if (rc != B200ALS_OK) { b200als_destroy(s); return rc; }
  CscDev<float>& A = s->csc[B200ALS_USERS];
  A.n_rows = n_item;
  A.n_cols = n_user_local;
  auto bail = [&](int code, const char* m) { b200als_destroy(s); return fail(code, m); };
  DevBuf len, temp;
  if (len.ensure(sizeof(int32_t) * ((size_t)n_user_local + 1)) != cudaSuccess) return bail(B200ALS_ECUDA, "alloc len");
  if (A.ptr.ensure(sizeof(int32_t) * ((size_t)n_user_local + 1)) != cudaSuccess) return bail(B200ALS_ECUDA, "alloc ptr");
  cudaMemsetAsync(len.p, 0, sizeof(int32_t) * ((size_t)n_user_local + 1), c.stream);
  if (n_user_local > 0) {
    synth_len_kernel<<<(n_user_local + 255) / 256, 256, 0, c.stream>>>(n_user_local, n_item, nnz_per_row, seed, len_dist, user_offset, len.i32());
    LAUNCHED();
  }
  // row pointers: 64-bit scan first (the total must be checked against the 32-bit limit of a shard), then narrowed
  DevBuf ptr64;
  if (ptr64.ensure(sizeof(long long) * ((size_t)n_user_local + 1)) != cudaSuccess) return bail(B200ALS_ECUDA, "alloc ptr64");
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, len.i32(), ptr64.as<long long>(), n_user_local + 1, c.stream);
  if (temp.ensure(tb) != cudaSuccess) return bail(B200ALS_ECUDA, "alloc scan");
  cub::DeviceScan::ExclusiveSum(temp.p, tb, len.i32(), ptr64.as<long long>(), n_user_local + 1, c.stream);
  LAUNCHED();
  long long total = 0;
  cudaMemcpyAsync(&total, ptr64.as<long long>() + n_user_local, sizeof(long long), cudaMemcpyDeviceToHost, c.stream);
  if (cudaStreamSynchronize(c.stream) != cudaSuccess) return bail(B200ALS_ECUDA, "sync");
  if (total > 2147483647LL) return bail(B200ALS_EINVAL, "local nnz exceeds 32-bit row pointers");
  convert_kernel<long long, int32_t><<<(n_user_local + 1 + 255) / 256, 256, 0, c.stream>>>(ptr64.as<long long>(), A.ptr.i32(), n_user_local + 1);
  LAUNCHED();
  A.nnz = total;
  if (A.idx.ensure(sizeof(int32_t) * (size_t)std::max<long long>(1, total)) != cudaSuccess) return bail(B200ALS_ECUDA, "alloc idx");
  if (A.val.ensure(sizeof(float) * (size_t)std::max<long long>(1, total)) != cudaSuccess) return bail(B200ALS_ECUDA, "alloc val");
  if (n_user_local > 0) {
    synth_fill_kernel<<<(n_user_local + 127) / 128, 128, 0, c.stream>>>(n_user_local, n_item, seed, s->opt.feedback == B200ALS_EXPLICIT,
                                                                     col_dist, user_offset, A.ptr.i32(), A.idx.i32(), A.val.f32());
    LAUNCHED();
  }
  if (cudaGetLastError() != cudaSuccess) return bail(B200ALS_ECUDA, "synthetic kernels");
  s->has[B200ALS_USERS] = true;
  s->shard_begin[B200ALS_USERS] = (int32_t)user_offset;
  s->shard_end[B200ALS_USERS] = (int32_t)user_offset + n_user_local;
  long long nnz_g = total;
  if (g_comm.world > 1) {
    DevBuf t;
    if (t.ensure(sizeof(long long)) != cudaSuccess) return bail(B200ALS_ECUDA, "alloc");
    cudaMemcpyAsync(t.p, &nnz_g, sizeof(nnz_g), cudaMemcpyHostToDevice, c.stream);
    if (g_nccl.AllReduce(t.p, t.p, 1, ncclInt64, ncclSum, g_comm.comm, c.stream) != ncclSuccess) return bail(B200ALS_ENCCL, "allreduce nnz");
    cudaMemcpyAsync(&nnz_g, t.p, sizeof(nnz_g), cudaMemcpyDeviceToHost, c.stream);
    cudaStreamSynchronize(c.stream);
  }
  s->nnz_global[B200ALS_USERS] = nnz_g;
  rc = session_counts(s);
  if (rc == B200ALS_OK && cudaStreamSynchronize(c.stream) != cudaSuccess) rc = fail(B200ALS_ECUDA, "sync");
  if (rc != B200ALS_OK) { b200als_destroy(s); return rc; }
  *out = s;
  return B200ALS_OK;
}
