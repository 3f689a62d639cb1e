// engine_topk.inl -- part of engine.cu (included there; not a standalone translation unit).
// ------------------------------------------------------------------------------------------------------
// top-k recommendation: `top_product` of src/matrix_top_product.cpp:20-102
// ------------------------------------------------------------------------------------------------------
__global__ void set_bits_kernel(const int32_t* __restrict__ ids_1based, int n, int n_item, uint32_t* __restrict__ bits) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int i = ids_1based[e] - 1;   // R indices
  if (i >= 0 && i < n_item) atomicOr(&bits[i >> 5], 1u << (i & 31));
}
static int run_topk(Ctx& c, const float* dX, long long n_user, const float* dY, int n_item, int rank, int top_k,
                    const int32_t* h_nr_ptr, const int32_t* h_nr_idx, const int32_t* h_exclude, int n_exclude,
                    double glob_mean, int32_t* h_idx_out, double* h_scores_out) {
  if (rank > kTopMaxRank) return fail(B200ALS_EUNSUPPORTED, "top_product: rank > 256 is not supported");
  if (top_k < 1 || top_k > kTopMaxK) return fail(B200ALS_EUNSUPPORTED, "top_product: k must be in 1..128");
  if (n_user <= 0) return B200ALS_OK;
  DevBuf nr_ptr, nr_idx, excl, bits, d_idx, d_sc;
  TopkParams P;
  P.x = dX; P.y = dY; P.n_user = n_user; P.n_item = n_item; P.rank = rank; P.top_k = top_k;
  P.nr_ptr = nullptr; P.nr_idx = nullptr; P.exclude_bits = nullptr; P.glob_mean = glob_mean;
  if (h_nr_ptr) {
    const long long nnz = h_nr_ptr[n_user];
    if (nnz > 0) {   // src/matrix_top_product.cpp:33: an empty filter matrix is ignored
      if (!h_nr_idx) return fail(B200ALS_EINVAL, "top_product: not_recommend indices missing");
      CU(nr_ptr.ensure(sizeof(int32_t) * (size_t)(n_user + 1)));
      CU(nr_idx.ensure(sizeof(int32_t) * (size_t)nnz));
      CU(cudaMemcpyAsync(nr_ptr.p, h_nr_ptr, sizeof(int32_t) * (size_t)(n_user + 1), cudaMemcpyHostToDevice, c.stream));
      CU(cudaMemcpyAsync(nr_idx.p, h_nr_idx, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, c.stream));
      P.nr_ptr = nr_ptr.i32();
      P.nr_idx = nr_idx.i32();
    }
  }
  if (n_exclude > 0) {
    if (!h_exclude) return fail(B200ALS_EINVAL, "top_product: exclude list missing");
    const size_t words = ((size_t)n_item + 31) / 32;
    CU(excl.ensure(sizeof(int32_t) * (size_t)n_exclude));
    CU(bits.ensure(sizeof(uint32_t) * words));
    CU(cudaMemsetAsync(bits.p, 0, sizeof(uint32_t) * words, c.stream));
    CU(cudaMemcpyAsync(excl.p, h_exclude, sizeof(int32_t) * (size_t)n_exclude, cudaMemcpyHostToDevice, c.stream));
    set_bits_kernel<<<(n_exclude + 255) / 256, 256, 0, c.stream>>>(excl.i32(), n_exclude, n_item, (uint32_t*)bits.p);
    LAUNCHED(); CU(cudaGetLastError());
    P.exclude_bits = (const uint32_t*)bits.p;
  }
  CU(d_idx.ensure(sizeof(int32_t) * (size_t)n_user * top_k));
  CU(d_sc.ensure(sizeof(double) * (size_t)n_user * top_k));
  P.idx_out = d_idx.i32();
  P.score_out = d_sc.f64();
  const size_t smem = sizeof(TopkSmem);
  CU(cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (n_user + kTopUB - 1) / kTopUB;
  topk_kernel<<<(unsigned)blocks, 256, smem, c.stream>>>(P);
  LAUNCHED(); CU(cudaGetLastError());
  CU(cudaMemcpyAsync(h_idx_out, d_idx.p, sizeof(int32_t) * (size_t)n_user * top_k, cudaMemcpyDeviceToHost, c.stream));
  if (h_scores_out)
    CU(cudaMemcpyAsync(h_scores_out, d_sc.p, sizeof(double) * (size_t)n_user * top_k, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  return B200ALS_OK;
}

extern "C" int b200als_top_product(const float* user_emb, int64_t n_user, const float* item_emb, int32_t n_item, int rank,
                                   int top_k, const int32_t* not_recommend_ptr, const int32_t* not_recommend_idx,
                                   const int32_t* exclude, int n_exclude, double glob_mean, int32_t* idx_out,
                                   double* scores_out) {
  Ctx& c = ctx();
  TRY(c.init());
  if (!user_emb || !item_emb || !idx_out || n_user < 0 || n_item <= 0 || rank <= 0) return fail(B200ALS_EINVAL, "bad argument");
  DevBuf dX, dY;
  CU(dX.ensure(sizeof(float) * (size_t)rank * (size_t)std::max<int64_t>(1, n_user)));
  CU(dY.ensure(sizeof(float) * (size_t)rank * (size_t)n_item));
  CU(cudaMemcpyAsync(dX.p, user_emb, sizeof(float) * (size_t)rank * (size_t)n_user, cudaMemcpyHostToDevice, c.stream));
  CU(cudaMemcpyAsync(dY.p, item_emb, sizeof(float) * (size_t)rank * (size_t)n_item, cudaMemcpyHostToDevice, c.stream));
  return run_topk(c, dX.f32(), n_user, dY.f32(), n_item, rank, top_k, not_recommend_ptr, not_recommend_idx, exclude, n_exclude,
                  glob_mean, idx_out, scores_out);
}
