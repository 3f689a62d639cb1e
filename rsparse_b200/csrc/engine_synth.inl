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
