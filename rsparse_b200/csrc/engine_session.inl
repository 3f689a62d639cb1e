// engine_session.inl -- part of engine.cu (included there; not a standalone translation unit).
// ------------------------------------------------------------------------------------------------------
// 2. session
// ------------------------------------------------------------------------------------------------------
struct b200als_session {
  b200als_options opt;
  int k = 0;
  int32_t n_user = 0, n_item = 0;
  // orientation [B200ALS_ITEMS]: columns = items (local block), idx = users ; [B200ALS_USERS]: columns = users
  CscDev<float> csc[2];
  bool has[2] = {false, false};
  int32_t shard_begin[2] = {0, 0}, shard_end[2] = {0, 0};
  int64_t nnz_global[2] = {0, 0};
  DevBuf fac[2];   // full factor matrices (stored in basis B): [ITEMS] k x n_item, [USERS] k x n_user
  DevBuf cnt[2];   // cnt[w][j] = nnz of row j of factor matrix w (global), for the dynamic-lambda regulariser
  DevBuf G, G64, Vt, Q, Qt, diag, B64, Btmp, Bf, scratch;
  // warm start of the eigen-decomposition: Weig[f] = the stored basis right after a half-iteration in which side f was the fixed
  // matrix = the eigenvectors of that side's Gram in TRUE coordinates; V0, T1, A0: k x k fp64 work space
  DevBuf Weig[2], V0, T1, A0;
  bool eig_warm[2] = {false, false};
  // bias terms (with_user_item_bias / global_bias, b200als_set_bias): factor matrices are `k` = rank + 2 wide as in R
  int with_biases = 0;
  double global_bias = 0.0;
  DevBuf gbb;                      // global_bias_base, [k]
  BiasScratch<float> bias_w;
  bool basis_identity = true;
  std::vector<int32_t> ranges[2];   // [3*world]: every rank's [begin, end, can_chunk) per orientation (multi-GPU)
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_chunk[8] = {}, ev_comm_done = nullptr;
  // peer-memory exchange (multi-GPU): every rank maps every other rank's factor matrices (CUDA IPC) and pushes its
  // freshly solved rows straight into them with the copy engines over NVLink -- no SM is taken from the solve
  static constexpr int kMaxPeers = 16;
  int p2p_state = 0;                      // 0 not tried yet, 1 active, -1 unavailable (NCCL broadcasts instead)
  float* peer_fac[2][kMaxPeers] = {};
  cudaStream_t push_stream[kMaxPeers] = {};
  cudaEvent_t ev_push[kMaxPeers] = {};
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float t_gram = 0, t_prep = 0, t_solve = 0, t_comm = 0;
};

extern "C" void b200als_default_options(b200als_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->feedback = B200ALS_IMPLICIT;
  o->solver = B200ALS_CONJUGATE_GRADIENT;
  o->cg_steps = 3;
  o->dynamic_lambda = 1;
  o->lambda = 0.0;
  o->kernel = 0;
}

static int session_alloc(b200als_session* s) {
  Ctx& c = ctx();
  const size_t k = (size_t)s->k;
  CU(s->fac[B200ALS_ITEMS].ensure(sizeof(float) * k * (size_t)s->n_item));
  CU(s->fac[B200ALS_USERS].ensure(sizeof(float) * k * (size_t)s->n_user));
  CU(cudaMemsetAsync(s->fac[0].p, 0, s->fac[0].bytes, c.stream));
  CU(cudaMemsetAsync(s->fac[1].p, 0, s->fac[1].bytes, c.stream));
  CU(s->G.ensure(sizeof(float) * k * k));
  CU(s->G64.ensure(sizeof(double) * k * k));
  CU(s->Vt.ensure(sizeof(double) * k * k));
  CU(s->Q.ensure(sizeof(float) * k * k));
  CU(s->Qt.ensure(sizeof(float) * k * k));
  CU(s->diag.ensure(sizeof(float) * k));
  CU(s->B64.ensure(sizeof(double) * k * k));
  CU(s->Btmp.ensure(sizeof(double) * k * k));
  CU(s->Bf.ensure(sizeof(float) * k * k));
  for (int w = 0; w < 2; w++) CU(s->Weig[w].ensure(sizeof(double) * k * k));
  CU(s->V0.ensure(sizeof(double) * k * k));
  CU(s->T1.ensure(sizeof(double) * k * k));
  CU(s->A0.ensure(sizeof(double) * k * k));
  set_identity_kernel<<<(unsigned)((k * k + 255) / 256), 256, 0, c.stream>>>(s->B64.f64(), (int)k);
  LAUNCHED(); CU(cudaGetLastError());
  s->basis_identity = true;
  for (auto& e : s->ev) CU(cudaEventCreate(&e));
  for (auto& e : s->ev_chunk) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&s->ev_comm_done, cudaEventDisableTiming));
  int lo = 0, hi = 0;
  CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CU(cudaStreamCreateWithPriority(&s->comm_stream, cudaStreamNonBlocking, hi));  // the exchange must get SM slots early
  return B200ALS_OK;
}

// cnt[which][j] += number of entries of row j seen in the local block of the *other* orientation
__global__ void count_idx_kernel(const int32_t* __restrict__ idx, long long nnz, float* __restrict__ cnt) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nnz) atomicAdd(&cnt[idx[e]], 1.0f);
}

static int session_counts(b200als_session* s) {
  // cnt_X for the dynamic-lambda regulariser (R/model_WRMF.R:305-315): nnz per row of the FIXED matrix, i.e. for
  // the user half (X = items) the nnz per item.  Counted from whichever orientation is present.
  Ctx& c = ctx();
  for (int w = 0; w < 2; w++) {
    const int32_t n = (w == B200ALS_ITEMS) ? s->n_item : s->n_user;
    CU(s->cnt[w].ensure(sizeof(float) * (size_t)std::max(1, n)));
    CU(cudaMemsetAsync(s->cnt[w].p, 0, sizeof(float) * (size_t)n, c.stream));
  }
  // orientation USERS has idx = items -> counts per item ; orientation ITEMS has idx = users -> counts per user
  for (int w = 0; w < 2; w++) {
    if (!s->has[w] || s->csc[w].nnz == 0) continue;
    const int other = 1 - w;
    count_idx_kernel<<<(unsigned)((s->csc[w].nnz + 255) / 256), 256, 0, c.stream>>>(s->csc[w].idx.i32(), s->csc[w].nnz,
                                                                                   s->cnt[other].f32());
    LAUNCHED(); CU(cudaGetLastError());
  }
  if (g_comm.world > 1) {
    for (int w = 0; w < 2; w++) {
      const int32_t n = (w == B200ALS_ITEMS) ? s->n_item : s->n_user;
      if (s->has[1 - w]) NC(g_nccl.AllReduce(s->cnt[w].p, s->cnt[w].p, (size_t)n, ncclFloat, ncclSum, g_comm.comm, c.stream));
    }
  }
  return B200ALS_OK;
}

extern "C" int b200als_create(b200als_session** out, const b200als_csc* c_ui, const b200als_csc* c_iu, int32_t n_user,
                              int32_t n_item, int rank, const b200als_options* opts) {
  Ctx& c = ctx();
  TRY(c.init());
  if (!out || rank <= 0 || n_user < 0 || n_item < 0) return fail(B200ALS_EINVAL, "bad argument");
  if (rank > 256) return fail(B200ALS_EUNSUPPORTED, "rank > 256 is not supported");
  b200als_session* s = new b200als_session();
  if (opts) s->opt = *opts; else b200als_default_options(&s->opt);
  s->k = rank;
  s->n_user = n_user;
  s->n_item = n_item;
  int rc = session_alloc(s);
  if (rc == B200ALS_OK && c_ui) {
    rc = upload_csc<float>(c_ui, s->csc[B200ALS_ITEMS], c.stream);
    s->has[B200ALS_ITEMS] = true;
    s->shard_begin[B200ALS_ITEMS] = 0;
    s->shard_end[B200ALS_ITEMS] = c_ui->n_cols;
  }
  if (rc == B200ALS_OK && c_iu) {
    rc = upload_csc<float>(c_iu, s->csc[B200ALS_USERS], c.stream);
    s->has[B200ALS_USERS] = true;
    s->shard_begin[B200ALS_USERS] = 0;
    s->shard_end[B200ALS_USERS] = c_iu->n_cols;
  }
  if (rc == B200ALS_OK) rc = session_counts(s);
  for (int w = 0; w < 2 && rc == B200ALS_OK; w++) {
    long long nnz = s->has[w] ? s->csc[w].nnz : 0;
    if (g_comm.world > 1) {
      DevBuf t;
      if (t.ensure(sizeof(long long)) != cudaSuccess) { rc = fail(B200ALS_ECUDA, "alloc"); break; }
      cudaMemcpyAsync(t.p, &nnz, sizeof(nnz), cudaMemcpyHostToDevice, c.stream);
      if (g_nccl.AllReduce(t.p, t.p, 1, ncclInt64, ncclSum, g_comm.comm, c.stream) != ncclSuccess) { rc = fail(B200ALS_ENCCL, "allreduce nnz"); break; }
      cudaMemcpyAsync(&nnz, t.p, sizeof(nnz), cudaMemcpyDeviceToHost, c.stream);
      cudaStreamSynchronize(c.stream);
    }
    s->nnz_global[w] = nnz;
  }
  if (rc == B200ALS_OK && cudaStreamSynchronize(c.stream) != cudaSuccess) rc = fail(B200ALS_ECUDA, "sync after upload");
  if (rc != B200ALS_OK) {
    b200als_destroy(s);
    return rc;
  }
  *out = s;
  return B200ALS_OK;
}

// ------------------------------------------------------------------------------------------------------
// Format ingest on the device: build the other orientation of the sparse matrix, i.e. what
// `MatrixExtra::as.csr.matrix` / `t_shallow` do on the host in R/model_WRMF.R:184-189.  A stable LSD radix sort
// (CUB) of the entries by their row id keeps, inside every new column, the source order = ascending source
// column, so the result satisfies the dgCMatrix invariant and is bit-reproducible.
// ------------------------------------------------------------------------------------------------------
__global__ void expand_columns_kernel(const int32_t* __restrict__ ptr, int n_cols, int32_t* __restrict__ col_of) {
  const int cidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (cidx >= n_cols) return;
  for (int e = ptr[cidx]; e < ptr[cidx + 1]; e++) col_of[e] = cidx;
}
__global__ void iota_kernel(int32_t* __restrict__ a, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (int32_t)i;
}
__global__ void gather_transposed_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ col_of,
                                         const float* __restrict__ val, long long nnz, int32_t* __restrict__ idx_out,
                                         float* __restrict__ val_out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnz) return;
  const int32_t e = perm[t];
  idx_out[t] = col_of[e];
  val_out[t] = val[e];
}
// ptr_out[r] = first position in the sorted key array with key >= r  (r = 0 .. n_rows)
__global__ void row_starts_kernel(const int32_t* __restrict__ sorted_keys, long long nnz, int n_rows, int32_t* __restrict__ ptr_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  long long lo = 0, hi = nnz;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (sorted_keys[mid] < r) lo = mid + 1; else hi = mid;
  }
  ptr_out[r] = (int32_t)lo;
}
static int transpose_on_device(Ctx& c, const CscDev<float>& src, CscDev<float>& dst) {
  dst.n_rows = src.n_cols;
  dst.n_cols = src.n_rows;
  dst.nnz = src.nnz;
  dst.n_short = -1;
  dst.plan_key = -1;
  const long long nnz = src.nnz;
  CU(dst.ptr.ensure(sizeof(int32_t) * ((size_t)dst.n_cols + 1)));
  CU(dst.idx.ensure(sizeof(int32_t) * (size_t)nnz));
  CU(dst.val.ensure(sizeof(float) * (size_t)nnz));
  if (nnz == 0) {
    CU(cudaMemsetAsync(dst.ptr.p, 0, sizeof(int32_t) * ((size_t)dst.n_cols + 1), c.stream));
    return B200ALS_OK;
  }
  DevBuf col_of, iota, perm, keys_out, temp;
  CU(col_of.ensure(sizeof(int32_t) * (size_t)nnz));
  CU(iota.ensure(sizeof(int32_t) * (size_t)nnz));
  CU(perm.ensure(sizeof(int32_t) * (size_t)nnz));
  CU(keys_out.ensure(sizeof(int32_t) * (size_t)nnz));
  expand_columns_kernel<<<(src.n_cols + 255) / 256, 256, 0, c.stream>>>(src.ptr.i32(), src.n_cols, col_of.i32());
  LAUNCHED(); CU(cudaGetLastError());
  iota_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, c.stream>>>(iota.i32(), nnz);
  LAUNCHED(); CU(cudaGetLastError());
  int bits = 1;
  while (bits < 31 && (1ll << bits) < (long long)std::max(1, src.n_rows)) bits++;
  size_t temp_bytes = 0;
  CU(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, src.idx.i32(), keys_out.i32(), iota.i32(), perm.i32(), (int)nnz, 0,
                                     bits, c.stream));
  CU(temp.ensure(temp_bytes));
  CU(cub::DeviceRadixSort::SortPairs(temp.p, temp_bytes, src.idx.i32(), keys_out.i32(), iota.i32(), perm.i32(), (int)nnz, 0,
                                     bits, c.stream));
  LAUNCHED();
  gather_transposed_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, c.stream>>>(perm.i32(), col_of.i32(), src.val.f32(), nnz,
                                                                              dst.idx.i32(), dst.val.f32());
  LAUNCHED(); CU(cudaGetLastError());
  row_starts_kernel<<<(dst.n_cols + 1 + 255) / 256, 256, 0, c.stream>>>(keys_out.i32(), nnz, dst.n_cols, dst.ptr.i32());
  LAUNCHED(); CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c.stream));
  return B200ALS_OK;
}

// ---- the same on a sharded matrix (multi-GPU) ---------------------------------------------------------------------
// Every rank holds a block of columns of orientation `have` (say its users, idx = item ids) and needs its block of
// columns of the other orientation (its items, idx = user ids).  (1) Each rank transposes its own block on the
// device: entries grouped by item, users ascending.  (2) The entries of the items owned by rank r are one contiguous
// range of that result; ranks exchange these ranges (and the matching column pointers) with grouped ncclSend /
// ncclRecv over NVLink.  (3) The received pieces arrive ordered by (source rank, item, user); sources own ascending
// user blocks, so a STABLE sort by item alone yields (item, user ascending): the dgCMatrix invariant, bit-identical
// to a single-GPU transpose of the whole matrix.
__global__ void add_offset_kernel(int32_t* __restrict__ a, long long n, int32_t off) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] += off;
}
// col_of[e] = c for the entries of column c of a piece whose pointers are ptr[0..n_cols] (absolute; piece starts at ptr[0])
__global__ void expand_piece_kernel(const int32_t* __restrict__ ptr, int n_cols, int32_t* __restrict__ col_of) {
  const int cidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (cidx >= n_cols) return;
  const int base = ptr[0];
  for (int e = ptr[cidx] - base; e < ptr[cidx + 1] - base; e++) col_of[e] = cidx;
}
static void block_of(long long n, int r, int world, int32_t* b, int32_t* e) {   // same split as parallel.shard_range
  const long long base = n / world, rem = n % world;
  *b = (int32_t)(r * base + std::min<long long>(r, rem));
  *e = (int32_t)(*b + base + (r < rem ? 1 : 0));
}
static int transpose_sharded(b200als_session* s, int have, int need) {
  Ctx& c = ctx();
  const int W = g_comm.world, me = g_comm.rank;
  const int32_t n_other = (need == B200ALS_ITEMS) ? s->n_item : s->n_user;   // columns of the needed orientation (global)
  const int32_t n_have_global = (have == B200ALS_ITEMS) ? s->n_item : s->n_user;
  // (1) local transpose: columns = all n_other ids, idx = LOCAL column ids of the block -> made global
  CscDev<float> T;
  CscDev<float>& src = s->csc[have];
  if (src.n_rows != n_other) return fail(B200ALS_EINVAL, "sharded transpose: the uploaded block does not span the other dimension");
  TRY(transpose_on_device(c, src, T));
  if (T.nnz > 0 && s->shard_begin[have] != 0) {
    add_offset_kernel<<<(unsigned)((T.nnz + 255) / 256), 256, 0, c.stream>>>(T.idx.i32(), T.nnz, s->shard_begin[have]);
    LAUNCHED(); CU(cudaGetLastError());
  }
  // (2) ranges per destination
  std::vector<int32_t> nb(W), ne(W);
  for (int r = 0; r < W; r++) block_of(n_other, r, W, &nb[r], &ne[r]);
  std::vector<int32_t> bounds(W + 1);
  for (int r = 0; r < W; r++)
    CU(cudaMemcpyAsync(&bounds[r], T.ptr.i32() + nb[r], sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
  CU(cudaMemcpyAsync(&bounds[W], T.ptr.i32() + n_other, sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  std::vector<long long> cnt((size_t)W * W, 0);   // cnt[src * W + dst]
  for (int r = 0; r < W; r++) cnt[(size_t)me * W + r] = (long long)bounds[r + 1] - bounds[r];
  DevBuf dcnt;
  CU(dcnt.ensure(sizeof(long long) * (size_t)W * W));
  CU(cudaMemcpyAsync(dcnt.as<long long>() + (size_t)me * W, &cnt[(size_t)me * W], sizeof(long long) * W, cudaMemcpyHostToDevice, c.stream));
  NC(g_nccl.AllGather(dcnt.as<long long>() + (size_t)me * W, dcnt.p, W, ncclInt64, g_comm.comm, c.stream));
  CU(cudaMemcpyAsync(cnt.data(), dcnt.p, sizeof(long long) * (size_t)W * W, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  const int32_t n_loc = ne[me] - nb[me];
  std::vector<long long> roff(W + 1, 0);
  for (int r = 0; r < W; r++) roff[r + 1] = roff[r] + cnt[(size_t)r * W + me];
  const long long nnz_in = roff[W];
  if (nnz_in > 2147483647LL) return fail(B200ALS_EUNSUPPORTED, "a shard of the transposed matrix exceeds 2^31-1 entries");
  DevBuf ridx, rval, rptr, col_of, iota, perm, keys_out, temp;
  CU(ridx.ensure(sizeof(int32_t) * (size_t)std::max<long long>(1, nnz_in)));
  CU(rval.ensure(sizeof(float) * (size_t)std::max<long long>(1, nnz_in)));
  CU(rptr.ensure(sizeof(int32_t) * (size_t)W * ((size_t)n_loc + 1)));
  NC(g_nccl.GroupStart());
  for (int r = 0; r < W; r++) {
    const long long sc = cnt[(size_t)me * W + r], rc = cnt[(size_t)r * W + me];
    NC(g_nccl.Send(T.ptr.i32() + nb[r], (size_t)(ne[r] - nb[r]) + 1, ncclInt32, r, g_comm.comm, c.stream));
    NC(g_nccl.Recv(rptr.i32() + (size_t)r * ((size_t)n_loc + 1), (size_t)n_loc + 1, ncclInt32, r, g_comm.comm, c.stream));
    if (sc > 0) {
      NC(g_nccl.Send(T.idx.i32() + bounds[r], (size_t)sc, ncclInt32, r, g_comm.comm, c.stream));
      NC(g_nccl.Send(T.val.f32() + bounds[r], (size_t)sc, ncclFloat, r, g_comm.comm, c.stream));
    }
    if (rc > 0) {
      NC(g_nccl.Recv(ridx.i32() + roff[r], (size_t)rc, ncclInt32, r, g_comm.comm, c.stream));
      NC(g_nccl.Recv(rval.f32() + roff[r], (size_t)rc, ncclFloat, r, g_comm.comm, c.stream));
    }
  }
  NC(g_nccl.GroupEnd());
  // (3) stable sort of the received pieces by local column
  CscDev<float>& dst = s->csc[need];
  dst.n_rows = n_have_global;
  dst.n_cols = n_loc;
  dst.nnz = nnz_in;
  dst.n_short = -1;
  dst.plan_key = -1;
  CU(dst.ptr.ensure(sizeof(int32_t) * ((size_t)n_loc + 1)));
  CU(dst.idx.ensure(sizeof(int32_t) * (size_t)std::max<long long>(1, nnz_in)));
  CU(dst.val.ensure(sizeof(float) * (size_t)std::max<long long>(1, nnz_in)));
  if (nnz_in == 0) {
    CU(cudaMemsetAsync(dst.ptr.p, 0, sizeof(int32_t) * ((size_t)n_loc + 1), c.stream));
  } else {
    CU(col_of.ensure(sizeof(int32_t) * (size_t)nnz_in));
    CU(iota.ensure(sizeof(int32_t) * (size_t)nnz_in));
    CU(perm.ensure(sizeof(int32_t) * (size_t)nnz_in));
    CU(keys_out.ensure(sizeof(int32_t) * (size_t)nnz_in));
    for (int r = 0; r < W; r++) {
      if (cnt[(size_t)r * W + me] == 0 || n_loc == 0) continue;
      expand_piece_kernel<<<(n_loc + 255) / 256, 256, 0, c.stream>>>(rptr.i32() + (size_t)r * ((size_t)n_loc + 1), n_loc,
                                                                   col_of.i32() + roff[r]);
      LAUNCHED(); CU(cudaGetLastError());
    }
    iota_kernel<<<(unsigned)((nnz_in + 255) / 256), 256, 0, c.stream>>>(iota.i32(), nnz_in);
    LAUNCHED(); CU(cudaGetLastError());
    int bits = 1;
    while (bits < 31 && (1ll << bits) < (long long)std::max(1, n_loc)) bits++;
    size_t temp_bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, col_of.i32(), keys_out.i32(), iota.i32(), perm.i32(), (int)nnz_in, 0, bits, c.stream));
    CU(temp.ensure(temp_bytes));
    CU(cub::DeviceRadixSort::SortPairs(temp.p, temp_bytes, col_of.i32(), keys_out.i32(), iota.i32(), perm.i32(), (int)nnz_in, 0, bits, c.stream));
    LAUNCHED();
    // dst.idx[t] = ridx[perm[t]], dst.val[t] = rval[perm[t]]  (gather_transposed_kernel with col_of := ridx)
    gather_transposed_kernel<<<(unsigned)((nnz_in + 255) / 256), 256, 0, c.stream>>>(perm.i32(), ridx.i32(), rval.f32(), nnz_in,
                                                                                   dst.idx.i32(), dst.val.f32());
    LAUNCHED(); CU(cudaGetLastError());
    row_starts_kernel<<<(n_loc + 1 + 255) / 256, 256, 0, c.stream>>>(keys_out.i32(), nnz_in, n_loc, dst.ptr.i32());
    LAUNCHED(); CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(c.stream));
  s->has[need] = true;
  s->shard_begin[need] = nb[me];
  s->shard_end[need] = ne[me];
  s->nnz_global[need] = s->nnz_global[have];
  s->ranges[need].clear();
  return B200ALS_OK;
}

extern "C" int b200als_build_missing_orientation(b200als_session* s) {
  Ctx& c = ctx();
  if (!s) return fail(B200ALS_EINVAL, "null session");
  if (s->has[0] && s->has[1]) return B200ALS_OK;
  if (!s->has[0] && !s->has[1]) return fail(B200ALS_EINVAL, "the session holds no sparse matrix");
  const int have = s->has[0] ? 0 : 1, need = 1 - have;
  if (g_comm.world > 1) {
    TRY(transpose_sharded(s, have, need));
    return session_counts(s);
  }
  TRY(transpose_on_device(c, s->csc[have], s->csc[need]));
  s->has[need] = true;
  s->shard_begin[need] = 0;
  s->shard_end[need] = s->csc[need].n_cols;
  s->nnz_global[need] = s->csc[need].nnz;
  s->ranges[need].clear();
  return session_counts(s);
}
// copy one orientation back to the host (tests / export): ptr[n_cols+1], idx[nnz], val[nnz]
extern "C" int b200als_get_orientation(b200als_session* s, int which, int32_t* ptr, int32_t* idx, float* val, int64_t* nnz_out) {
  Ctx& c = ctx();
  if (!s || which < 0 || which > 1 || !s->has[which]) return fail(B200ALS_EINVAL, "orientation not present");
  const CscDev<float>& A = s->csc[which];
  if (nnz_out) *nnz_out = A.nnz;
  if (ptr) CU(cudaMemcpyAsync(ptr, A.ptr.p, sizeof(int32_t) * ((size_t)A.n_cols + 1), cudaMemcpyDeviceToHost, c.stream));
  if (idx && A.nnz) CU(cudaMemcpyAsync(idx, A.idx.p, sizeof(int32_t) * (size_t)A.nnz, cudaMemcpyDeviceToHost, c.stream));
  if (val && A.nnz) CU(cudaMemcpyAsync(val, A.val.p, sizeof(float) * (size_t)A.nnz, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  return B200ALS_OK;
}

extern "C" int b200als_destroy(b200als_session* s) {
  if (!s) return B200ALS_OK;
  for (auto& e : s->ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : s->ev_chunk)
    if (e) cudaEventDestroy(e);
  if (s->ev_comm_done) cudaEventDestroy(s->ev_comm_done);
  if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
  if (s->p2p_state == 1) {
    cudaDeviceSynchronize();
    for (int w = 0; w < 2; w++)
      for (int r = 0; r < b200als_session::kMaxPeers; r++)
        if (s->peer_fac[w][r]) cudaIpcCloseMemHandle(s->peer_fac[w][r]);
    for (int r = 0; r < b200als_session::kMaxPeers; r++) {
      if (s->push_stream[r]) cudaStreamDestroy(s->push_stream[r]);
      if (s->ev_push[r]) cudaEventDestroy(s->ev_push[r]);
    }
    // nobody frees a matrix a peer still has mapped: destroy is collective while the communicator lives
    if (g_comm.comm) {
      Ctx& c = ctx();
      cudaMemsetAsync(c.status.p, 0, sizeof(int), c.stream);
      g_nccl.AllReduce(c.status.p, c.status.p, 1, ncclInt32, ncclSum, g_comm.comm, c.stream);
      cudaStreamSynchronize(c.stream);
    }
  }
  delete s;
  return B200ALS_OK;
}

extern "C" int b200als_exchange_mode(b200als_session* s, int* mode) {
  if (!s || !mode) return fail(B200ALS_EINVAL, "bad argument");
  *mode = (g_comm.world <= 1 || s->p2p_state == 0) ? 0 : (s->p2p_state == 1 ? 1 : 2);
  return B200ALS_OK;
}

extern "C" int b200als_set_shard(b200als_session* s, int which, int32_t begin, int32_t end) {
  if (!s || which < 0 || which > 1) return fail(B200ALS_EINVAL, "bad argument");
  const int32_t n = (which == B200ALS_ITEMS) ? s->n_item : s->n_user;
  if (begin < 0 || end < begin || end > n || (s->has[which] && end - begin != s->csc[which].n_cols))
    return fail(B200ALS_EINVAL, "shard range does not match the uploaded block");
  s->shard_begin[which] = begin;
  s->shard_end[which] = end;
  s->ranges[which].clear();
  return B200ALS_OK;
}

static int rotate_matrix(Ctx& c, float* M, long long n, const float* R, int k) {
  if (n <= 0) return B200ALS_OK;
  if (k != kTcK) {
    // other ranks (k % 4 == 0, k <= 256): fp32 FFMA kernel, R streamed through shared memory
    if (k % 4 != 0 || k > 256) return fail(B200ALS_EUNSUPPORTED, "change of basis needs rank % 4 == 0 and rank <= 256");
    const long long blocks = (n + kRotRows - 1) / kRotRows;
    auto launch = [&](auto kern, size_t smem) -> cudaError_t {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      int per_sm = 1;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem);
      if (e != cudaSuccess) return e;
      const int grid = (int)std::min<long long>(blocks, (long long)c.sm_count * std::max(1, per_sm));
      kern<<<grid, 256, smem, c.stream>>>(M, M, R, n, k);
      return cudaSuccess;
    };
    if (k <= 64) CU(launch(rotate_any_kernel<64>, sizeof(RotAnySmem<64>)));
    else if (k <= 128) CU(launch(rotate_any_kernel<128>, sizeof(RotAnySmem<128>)));
    else CU(launch(rotate_any_kernel<256>, sizeof(RotAnySmem<256>)));
    LAUNCHED(); CU(cudaGetLastError());
    return B200ALS_OK;
  }
  // large matrices: tcgen05 3xTF32 kernel (B200ALS_ROTATE=ffma forces the fp32 FMA kernel)
  const char* env = getenv("B200ALS_ROTATE");
  const bool force_tc = env && (env[0] == 't' || env[0] == 'T');   // tests
  const bool tc = force_tc || (!(env && (env[0] == 'f' || env[0] == 'F')) && n >= 65536);
  if (tc) {
    CU(c.rot_rt.ensure(sizeof(float) * kTcK * kTcK));
    transpose_128_kernel<<<(kTcK * kTcK + 255) / 256, 256, 0, c.stream>>>(R, c.rot_rt.f32());
    LAUNCHED(); CU(cudaGetLastError());
    const size_t smem = sizeof(RotTcSmem);
    CU(cudaFuncSetAttribute(rotate_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tiles = (n + 127) / 128;
    const int grid = (int)std::min<long long>(tiles, c.sm_count);
    rotate_tc_kernel<<<grid, 128, smem, c.stream>>>(M, M, c.rot_rt.f32(), n);
    LAUNCHED(); CU(cudaGetLastError());
    return B200ALS_OK;
  }
  const size_t smem = sizeof(RotSmem);
  CU(cudaFuncSetAttribute(rotate_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (n + kRotRows - 1) / kRotRows;
  const int grid = (int)std::min<long long>(blocks, c.sm_count * 2);
  rotate_rows_kernel<<<grid, 256, smem, c.stream>>>(M, M, R, n);
  LAUNCHED(); CU(cudaGetLastError());
  return B200ALS_OK;
}

// true = stored * B'  <=> stored = true * B.  Export / import copies through a scratch buffer.
extern "C" int b200als_set_factors(b200als_session* s, int which, const float* host) {
  Ctx& c = ctx();
  if (!s || !host || which < 0 || which > 1) return fail(B200ALS_EINVAL, "bad argument");
  const long long n = (which == B200ALS_ITEMS) ? s->n_item : s->n_user;
  CU(cudaMemcpyAsync(s->fac[which].p, host, sizeof(float) * (size_t)s->k * (size_t)n, cudaMemcpyHostToDevice, c.stream));
  s->eig_warm[which] = false;   // a replaced matrix has an unrelated Gram
  if (!s->basis_identity) {
    convert_kk_kernel<<<(s->k * s->k + 255) / 256, 256, 0, c.stream>>>(s->B64.f64(), s->Bf.f32(), s->k, 0);
    LAUNCHED(); CU(cudaGetLastError());
    TRY(rotate_matrix(c, s->fac[which].f32(), n, s->Bf.f32(), s->k));
  }
  CU(cudaStreamSynchronize(c.stream));
  return B200ALS_OK;
}
static int export_rotated(b200als_session* s, const float* dev, long long n, float* host) {
  Ctx& c = ctx();
  const size_t bytes = sizeof(float) * (size_t)s->k * (size_t)n;
  if (s->basis_identity) {
    CU(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c.stream));
  } else {
    CU(s->scratch.ensure(bytes));
    CU(cudaMemcpyAsync(s->scratch.p, dev, bytes, cudaMemcpyDeviceToDevice, c.stream));
    convert_kk_kernel<<<(s->k * s->k + 255) / 256, 256, 0, c.stream>>>(s->B64.f64(), s->Bf.f32(), s->k, 1);
    LAUNCHED(); CU(cudaGetLastError());
    TRY(rotate_matrix(c, s->scratch.f32(), n, s->Bf.f32(), s->k));
    CU(cudaMemcpyAsync(host, s->scratch.p, bytes, cudaMemcpyDeviceToHost, c.stream));
  }
  CU(cudaStreamSynchronize(c.stream));
  return B200ALS_OK;
}
extern "C" int b200als_get_factors(b200als_session* s, int which, float* host) {
  if (!s || !host || which < 0 || which > 1) return fail(B200ALS_EINVAL, "bad argument");
  const long long n = (which == B200ALS_ITEMS) ? s->n_item : s->n_user;
  return export_rotated(s, s->fac[which].f32(), n, host);
}

extern "C" int b200als_init_factors(b200als_session* s, uint64_t seed) {
  Ctx& c = ctx();
  if (!s) return fail(B200ALS_EINVAL, "null session");
  const long long nu = (long long)s->k * s->n_user, ni = (long long)s->k * s->n_item;
  if (nu) init_normal_kernel<<<(unsigned)((nu + 255) / 256), 256, 0, c.stream>>>(s->fac[B200ALS_USERS].f32(), nu, seed, 0.01f);
  LAUNCHED(); CU(cudaGetLastError());
  if (s->opt.solver == B200ALS_CONJUGATE_GRADIENT) {
    CU(cudaMemsetAsync(s->fac[B200ALS_ITEMS].p, 0, sizeof(float) * (size_t)ni, c.stream));  // R/model_WRMF.R:217-230
  } else if (ni) {
    init_normal_kernel<<<(unsigned)((ni + 255) / 256), 256, 0, c.stream>>>(s->fac[B200ALS_ITEMS].f32(), ni,
                                                                           seed ^ 0xA5A5A5A5ull, 0.01f);
    LAUNCHED(); CU(cudaGetLastError());
  }
  set_identity_kernel<<<(s->k * s->k + 255) / 256, 256, 0, c.stream>>>(s->B64.f64(), s->k);
  LAUNCHED(); CU(cudaGetLastError());
  s->basis_identity = true;
  s->eig_warm[0] = s->eig_warm[1] = false;
  CU(cudaStreamSynchronize(c.stream));
  return B200ALS_OK;
}

extern "C" int b200als_randomize_factors(b200als_session* s, int which, uint64_t seed, float scale, float decay) {
  Ctx& c = ctx();
  if (!s || which < 0 || which > 1) return fail(B200ALS_EINVAL, "bad argument");
  const long long n = (long long)s->k * ((which == B200ALS_ITEMS) ? s->n_item : s->n_user);
  if (n) init_normal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(s->fac[which].f32(), n, seed, scale, s->k, decay);
  LAUNCHED(); CU(cudaGetLastError());
  s->eig_warm[which] = false;   // a replaced matrix has an unrelated Gram
  CU(cudaStreamSynchronize(c.stream));
  return B200ALS_OK;
}

// every rank learns every rank's [begin, end) of the solved matrix (cached until set_shard)
static int gather_ranges(b200als_session* s, int which) {
  if (!s->ranges[which].empty()) return B200ALS_OK;
  Ctx& c = ctx();
  // chunked solves (exchange overlapped with the solve) need every row of the block in ONE length class of the CG path
  const bool cg_fast = (s->k % 4 == 0) && (s->k <= 256);
  bool gram_rows = false;
  if (cg_fast) {
    TRY(gram_rows_for(c, s->csc[which], s->k, s->opt.feedback, &gram_rows));
    TRY(plan_rows(c, s->csc[which], s->k, s->k == kResK && s->opt.kernel != 10, false, gram_rows));
  }
  const bool one_class = cg_fast && s->csc[which].plan_single >= 0 &&
                         (s->csc[which].plan_single != CscDev<float>::kClsLong || gram_rows) &&
                         !s->csc[which].cls[s->csc[which].plan_single].stream;
  std::vector<int32_t> ranges(3 * g_comm.world);
  DevBuf d;
  CU(d.ensure(sizeof(int32_t) * 3 * g_comm.world));
  int32_t mine[3] = {s->shard_begin[which], s->shard_end[which],
                     (one_class && s->csc[which].n_cols >= 8 * 4096) ? 1 : 0};
  CU(cudaMemcpyAsync(d.i32() + 3 * g_comm.rank, mine, sizeof(mine), cudaMemcpyHostToDevice, c.stream));
  NC(g_nccl.AllGather(d.i32() + 3 * g_comm.rank, d.p, 3, ncclInt32, g_comm.comm, c.stream));
  CU(cudaMemcpyAsync(ranges.data(), d.p, sizeof(int32_t) * 3 * g_comm.world, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  s->ranges[which] = ranges;
  return B200ALS_OK;
}
// exchange of the freshly solved rows: chunk `ch` of `n_ch` of every rank's block, one broadcast per owner
// (unequal block sizes allowed), grouped, on stream `st`.
static int exchange_chunk(b200als_session* s, int which, int ch, int n_ch, cudaStream_t st) {
  if (g_comm.world <= 1) return B200ALS_OK;
  const std::vector<int32_t>& ranges = s->ranges[which];
  float* M = s->fac[which].f32();
  NC(g_nccl.GroupStart());
  for (int r = 0; r < g_comm.world; r++) {
    const long long rb = ranges[3 * r], len = ranges[3 * r + 1] - rb;
    const long long cb = rb + len * ch / n_ch, ce = rb + len * (ch + 1) / n_ch;
    if (ce <= cb) continue;
    float* p = M + (size_t)cb * s->k;
    NC(g_nccl.Broadcast(p, p, (size_t)(ce - cb) * (size_t)s->k, ncclFloat, r, g_comm.comm, st));
  }
  NC(g_nccl.GroupEnd());
  return B200ALS_OK;
}
// Maps the peers' factor matrices.  Collective: every rank calls it at the same point.  Falls back to NCCL (state -1)
// unless every rank could open every handle (B200ALS_EXCHANGE=nccl forces the fallback, =p2p makes failure an error).
static int p2p_setup(b200als_session* s) {
  if (s->p2p_state != 0) return B200ALS_OK;
  Ctx& c = ctx();
  const char* ev = getenv("B200ALS_EXCHANGE");
  const bool force_nccl = ev && !strcmp(ev, "nccl"), force_p2p = ev && !strcmp(ev, "p2p");
  const int W = g_comm.world, me = g_comm.rank;
  int ok = (!force_nccl && W <= b200als_session::kMaxPeers) ? 1 : 0;
  cudaIpcMemHandle_t mine[2];
  std::memset(mine, 0, sizeof(mine));
  if (ok)
    for (int w = 0; w < 2; w++)
      if (cudaIpcGetMemHandle(&mine[w], s->fac[w].p) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  const size_t hb = sizeof(mine);
  DevBuf d, flag;
  CU(d.ensure(hb * (size_t)W));
  CU(flag.ensure(sizeof(int)));
  CU(cudaMemcpyAsync((char*)d.p + hb * me, mine, hb, cudaMemcpyHostToDevice, c.stream));
  NC(g_nccl.AllGather((char*)d.p + hb * me, d.p, hb, ncclChar, g_comm.comm, c.stream));
  std::vector<cudaIpcMemHandle_t> all(2 * (size_t)W);
  CU(cudaMemcpyAsync(all.data(), d.p, hb * (size_t)W, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaMemcpyAsync(flag.p, &ok, sizeof(int), cudaMemcpyHostToDevice, c.stream));
  NC(g_nccl.AllReduce(flag.p, flag.p, 1, ncclInt32, ncclMin, g_comm.comm, c.stream));
  CU(cudaMemcpyAsync(&ok, flag.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  if (ok) {
    for (int r = 0; r < W && ok; r++) {
      if (r == me) continue;
      for (int w = 0; w < 2; w++) {
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, all[2 * (size_t)r + w], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          ok = 0;
          cudaGetLastError();
          break;
        }
        s->peer_fac[w][r] = (float*)q;
      }
    }
    CU(cudaMemcpyAsync(flag.p, &ok, sizeof(int), cudaMemcpyHostToDevice, c.stream));
    NC(g_nccl.AllReduce(flag.p, flag.p, 1, ncclInt32, ncclMin, g_comm.comm, c.stream));
    CU(cudaMemcpyAsync(&ok, flag.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    CU(cudaStreamSynchronize(c.stream));
  }
  if (!ok) {
    for (int w = 0; w < 2; w++)
      for (int r = 0; r < b200als_session::kMaxPeers; r++)
        if (s->peer_fac[w][r]) { cudaIpcCloseMemHandle(s->peer_fac[w][r]); s->peer_fac[w][r] = nullptr; }
    s->p2p_state = -1;
    if (force_p2p) return fail(B200ALS_ECUDA, "B200ALS_EXCHANGE=p2p: peer mapping of the factor matrices failed on some rank");
    return B200ALS_OK;
  }
  for (int r = 0; r < W; r++) {
    if (r == me) continue;
    CU(cudaStreamCreateWithFlags(&s->push_stream[r], cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&s->ev_push[r], cudaEventDisableTiming));
  }
  s->p2p_state = 1;
  return B200ALS_OK;
}
// pushes chunk `ch` of `n_ch` of this rank's block into every peer's copy of the matrix, one copy-engine stream per
// peer, after `ready` (the chunk's solve).  Completion is collected by p2p_join().
static int p2p_push_chunk(b200als_session* s, int which, int ch, int n_ch, cudaEvent_t ready) {
  const std::vector<int32_t>& ranges = s->ranges[which];
  const int me = g_comm.rank;
  const long long rb = ranges[3 * me], len = ranges[3 * me + 1] - rb;
  const long long cb = rb + len * ch / n_ch, ce = rb + len * (ch + 1) / n_ch;
  if (ce <= cb) return B200ALS_OK;
  const size_t off = (size_t)cb * s->k, bytes = sizeof(float) * (size_t)(ce - cb) * s->k;
  const float* src = s->fac[which].f32() + off;
  for (int i = 1; i < g_comm.world; i++) {
    const int r = (me + i) % g_comm.world;   // staggered start: no two ranks open on the same destination
    CU(cudaStreamWaitEvent(s->push_stream[r], ready, 0));
    CU(cudaMemcpyAsync(s->peer_fac[which][r] + off, src, bytes, cudaMemcpyDeviceToDevice, s->push_stream[r]));
  }
  return B200ALS_OK;
}
static int p2p_join(b200als_session* s, cudaStream_t st) {
  for (int r = 0; r < g_comm.world; r++) {
    if (r == g_comm.rank) continue;
    CU(cudaEventRecord(s->ev_push[r], s->push_stream[r]));
    CU(cudaStreamWaitEvent(st, s->ev_push[r], 0));
  }
  return B200ALS_OK;
}
__global__ void set_col_kernel(float* __restrict__ M, int ld, int col, int n, float v) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) M[(size_t)r * ld + col] = v;
}
__global__ void status_to_double_kernel(const int* __restrict__ status, double* __restrict__ out) { out[0] = (status[0] != 0) ? 1.0 : 0.0; }
__global__ void finalize_gram_kernel(double* __restrict__ G64, float* __restrict__ G, int k, double lambda) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= k * k) return;
  const double v = G64[e] + (((e / k) == (e % k)) ? lambda : 0.0);
  G64[e] = v;
  G[e] = (float)v;
}

// solve for `which`; Yout == nullptr: in place into the session's factors.  Yout != nullptr (transform_):
// solve into Yout (device, local block) starting from zeros.
static int session_half(b200als_session* s, int which, int solver, float* Yout, double* loss) {
  Ctx& c = ctx();
  if (!s->has[which]) return fail(B200ALS_EINVAL, "the orientation needed for this half-iteration was not supplied");
  const int fixed = 1 - which;
  const long long n_fixed = (fixed == B200ALS_ITEMS) ? s->n_item : s->n_user;
  float* X = s->fac[fixed].f32();
  float* Yfull = s->fac[which].f32();
  float* Y = Yout ? Yout : (Yfull + (size_t)s->shard_begin[which] * s->k);
  HalfOpts o{s->opt.feedback, solver, s->opt.cg_steps, s->opt.dynamic_lambda, s->opt.kernel, s->opt.lambda, s->opt.reserved[0], s->opt.reserved[1]};
  CscDev<float>& A = s->csc[which];
  const bool implicit = (o.feedback == B200ALS_IMPLICIT);
  if (s->with_biases || (implicit && s->global_bias >= std::sqrt((double)std::numeric_limits<float>::epsilon()))) {
    // bias terms inside the device-resident session (wrmf_implicit.hpp:105-157,190-232 / wrmf_explicit.hpp:57-64,87-91):
    // the same device code as the stateless calls (half_on_device), no host round trip.  The ITEM half is the call with
    // is_bias_last_row = TRUE, the USER half (and transform_) the one with FALSE (R/model_WRMF.R:321,327,436).
    if (g_comm.world > 1) return fail(B200ALS_EUNSUPPORTED, "bias terms in a multi-GPU session are not implemented");
    NvtxRange nvb("b200als/half_iteration/biased");
    CU(s->gbb.ensure(sizeof(float) * (size_t)s->k));
    CU(cudaEventRecord(s->ev[0], c.stream));
    TRY(half_on_device<float>(c, A, s->k, X, Y, n_fixed, (long long)A.n_cols, nullptr, o, s->with_biases,
                              which == B200ALS_ITEMS ? 1 : 0, s->global_bias, s->gbb.f32(), 1, s->bias_w));
    CU(cudaEventRecord(s->ev[3], c.stream));
    TRY(finish_loss<float>(c, X, s->k, n_fixed, s->cnt[fixed].f32(), o, s->nnz_global[which], 0.0, false, loss));
    s->t_gram = s->t_prep = s->t_comm = 0.f;
    cudaEventElapsedTime(&s->t_solve, s->ev[0], s->ev[3]);
    return B200ALS_OK;
  }
  NvtxRange nv_half(which == B200ALS_USERS ? "b200als/half_iteration/users" : "b200als/half_iteration/items");
  struct GramModeScope {   // the session's Gram arithmetic (options.reserved[2]) applies to every run_gram of this call
    explicit GramModeScope(int m) { g_gram_mode_override = (m >= 1 && m <= 3) ? (m == 3 ? 0 : m) : -1; }
    ~GramModeScope() { g_gram_mode_override = -1; }
  } gram_scope(s->opt.reserved[2]);
  CU(cudaEventRecord(s->ev[0], c.stream));
  const float* G = nullptr;
  const float* diag = nullptr;
  if (g_comm.world > 1) TRY(gather_ranges(s, which));
  if (g_comm.world > 1 && !Yout) TRY(p2p_setup(s));
  const bool p2p = (g_comm.world > 1 && !Yout && s->p2p_state == 1);
  if (p2p && !implicit) {
    // one-sided pushes need every rank to have finished its earlier writes to the matrix (set_factors, randomize, the
    // previous half-iteration) before any peer writes into it: implicit feedback gets that from the Gram all-reduce
    // below, explicit feedback from this one-word all-reduce
    CU(cudaMemsetAsync(c.status.p, 0, sizeof(int), c.stream));
    NC(g_nccl.AllReduce(c.status.p, c.status.p, 1, ncclInt32, ncclSum, g_comm.comm, c.stream));
  }
  if (implicit) {
    NvtxRange nv("b200als/gram");
    if (g_comm.world > 1) {
      // each rank reduces its 1/world slice of the fixed matrix; the k x k partials are summed over NVLink
      const long long b = n_fixed * g_comm.rank / g_comm.world, e = n_fixed * (g_comm.rank + 1) / g_comm.world;
      TRY(run_gram<float>(c, X + (size_t)b * s->k, s->k, e - b, 0.0, s->G.f32(), s->G64.f64()));
      NC(g_nccl.AllReduce(s->G64.p, s->G64.p, (size_t)s->k * s->k, ncclDouble, ncclSum, g_comm.comm, c.stream));
      finalize_gram_kernel<<<(s->k * s->k + 255) / 256, 256, 0, c.stream>>>(s->G64.f64(), s->G.f32(), s->k, o.lambda);
      LAUNCHED(); CU(cudaGetLastError());
    } else {
      TRY(run_gram<float>(c, X, s->k, n_fixed, o.lambda, s->G.f32(), s->G64.f64()));
    }
    G = s->G.f32();
  }
  CU(cudaEventRecord(s->ev[1], c.stream));
  // eigenbasis path: implicit CG, rank 128, resident kernel, enough rows to amortise the rotation
  // The decision uses GLOBAL quantities only (rank, options, the global number of solved rows), so every rank of a
  // multi-GPU run takes the same branch: rows of any length are solved in the eigenbasis (resident / tile / streaming
  // kernels all take `diag`), a shard with long rows no longer opts out on its own.
  const long long n_solved_global = (which == B200ALS_ITEMS) ? s->n_item : s->n_user;
  bool use_diag = implicit && solver == B200ALS_CONJUGATE_GRADIENT && (s->k % 4 == 0) && s->k <= 256 && o.kernel != 1 &&
                  o.kernel != 2 && !Yout;
  if (use_diag && o.kernel != 3 && o.kernel != 10 && n_solved_global < 50000) use_diag = false;
  if (use_diag) {
    NvtxRange nv("b200als/eigenbasis");
    const size_t jsm = sizeof(double) * (size_t)s->k * (s->k + 1);
    const int a_in_smem = (jsm + 8192 <= c.smem_optin) ? 1 : 0;
    if (a_in_smem) CU(cudaFuncSetAttribute(jacobi_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jsm));
    // Warm start (B200ALS_EIG_WARM=0 disables): the Gram of a side changes slowly from one ALS iteration to the next, so
    // its previous eigenvectors W (kept in true coordinates) nearly diagonalise it.  In the stored basis B they read
    // V0 = B' W; Jacobi then runs on A0 = V0' G V0 (2-3 sweeps instead of ~9) and Q = V0 V1.
    const char* ew = getenv("B200ALS_EIG_WARM");
    const bool warm = s->eig_warm[fixed] && !(ew && ew[0] == '0');
    const unsigned kkb = (unsigned)((s->k * s->k + 255) / 256);
    if (warm) {
      matmul_kk_kernel<<<kkb, 256, 0, c.stream>>>(s->B64.f64(), s->Weig[fixed].f64(), s->V0.f64(), s->k, 1);   // V0 = B' W
      LAUNCHED();
      matmul_kk_kernel<<<kkb, 256, 0, c.stream>>>(s->G64.f64(), s->V0.f64(), s->T1.f64(), s->k, 0);            // T1 = G V0
      LAUNCHED();
      matmul_kk_kernel<<<kkb, 256, 0, c.stream>>>(s->V0.f64(), s->T1.f64(), s->A0.f64(), s->k, 1);             // A0 = V0' T1
      LAUNCHED();
      jacobi_eig_kernel<<<1, kJacobiThreads, a_in_smem ? jsm : 0, c.stream>>>(s->A0.f64(), s->Vt.f64(), s->k, s->Q.f32(),
                                                                            s->diag.f32(), s->T1.f64(), 30, a_in_smem);
      LAUNCHED(); CU(cudaGetLastError());
      matmul_kk_kernel<<<kkb, 256, 0, c.stream>>>(s->V0.f64(), s->T1.f64(), s->Btmp.f64(), s->k, 0);           // Q = V0 V1
      LAUNCHED();
      convert_kk_kernel<<<kkb, 256, 0, c.stream>>>(s->Btmp.f64(), s->Q.f32(), s->k, 0);
      LAUNCHED(); CU(cudaGetLastError());
    } else {
      jacobi_eig_kernel<<<1, kJacobiThreads, a_in_smem ? jsm : 0, c.stream>>>(s->G64.f64(), s->Vt.f64(), s->k, s->Q.f32(),
                                                                            s->diag.f32(), s->Btmp.f64(), 30, a_in_smem);
      LAUNCHED(); CU(cudaGetLastError());
    }
    // fixed <- fixed Q (whole matrix), solved slice <- slice Q, B <- B Q
    TRY(rotate_matrix(c, X, n_fixed, s->Q.f32(), s->k));
    TRY(rotate_matrix(c, Y, A.n_cols, s->Q.f32(), s->k));
    matmul_kk_kernel<<<kkb, 256, 0, c.stream>>>(s->B64.f64(), s->Btmp.f64(), s->Vt.f64(), s->k);
    LAUNCHED(); CU(cudaGetLastError());
    CU(cudaMemcpyAsync(s->B64.p, s->Vt.p, sizeof(double) * (size_t)s->k * s->k, cudaMemcpyDeviceToDevice, c.stream));
    CU(cudaMemcpyAsync(s->Weig[fixed].p, s->Vt.p, sizeof(double) * (size_t)s->k * s->k, cudaMemcpyDeviceToDevice, c.stream));
    s->eig_warm[fixed] = true;
    s->basis_identity = false;
    diag = s->diag.f32();
    G = nullptr;
  }
  CU(cudaEventRecord(s->ev[2], c.stream));
  nvtxRangePushA("b200als/solve+exchange");
  if (g_comm.world > 1 && !Yout) {
    // the block is solved in chunks; chunk c travels to the other ranks (priority stream) while chunk c+1 is solved
    // every rank must take the same decision: chunk only if every block qualifies.  More chunks = shorter
    // exposed tail of the exchange (only the last chunk's broadcast is not hidden behind a solve)
    int n_ch = 8;
    if (const char* ev = getenv("B200ALS_EXCHANGE_CHUNKS")) n_ch = std::max(1, std::min(8, atoi(ev)));
    for (int r = 0; r < g_comm.world; r++)
      if (!s->ranges[which][3 * r + 2]) n_ch = 1;
    for (int ch = 0; ch < n_ch; ch++) {
      HalfOpts oc = o;
      if (n_ch > 1) {
        oc.row_begin = (int)((long long)A.n_cols * ch / n_ch);
        oc.row_count = (int)((long long)A.n_cols * (ch + 1) / n_ch) - oc.row_begin;
      }
      oc.reset_loss = (ch == 0);
      TRY(solve_rows<float>(c, A, X, Y, G, diag, s->k, oc));
      CU(cudaEventRecord(s->ev_chunk[ch], c.stream));
      if (p2p) {
        TRY(p2p_push_chunk(s, which, ch, n_ch, s->ev_chunk[ch]));
      } else {
        CU(cudaStreamWaitEvent(s->comm_stream, s->ev_chunk[ch], 0));
        TRY(exchange_chunk(s, which, ch, n_ch, s->comm_stream));
      }
    }
    CU(cudaEventRecord(s->ev[3], c.stream));
    if (p2p) {
      // own pushes done; the loss all-reduce below completes only when every rank got here, i.e. when every push
      // into this rank's matrix has landed
      TRY(p2p_join(s, c.stream));
    } else {
      CU(cudaEventRecord(s->ev_comm_done, s->comm_stream));
      CU(cudaStreamWaitEvent(c.stream, s->ev_comm_done, 0));
    }
  } else {
    TRY(solve_rows<float>(c, A, X, Y, G, diag, s->k, o));
    CU(cudaEventRecord(s->ev[3], c.stream));
  }
  CU(cudaEventRecord(s->ev[4], c.stream));
  nvtxRangePop();
  NvtxRange nv_loss("b200als/loss");
  // loss: local row sums -> global
  if (g_comm.world > 1) {
    // One all-reduce carries the row sums, the regulariser (each rank squares its 1 / world slice of the fixed matrix instead of
    // every rank reading all of it) and the status word (a non-SPD row on ONE rank must fail the call on EVERY rank).  With
    // the peer-memory exchange it is also the completion barrier of the pushes (see above).
    const long long b = n_fixed * g_comm.rank / g_comm.world, e = n_fixed * (g_comm.rank + 1) / g_comm.world;
    CU(cudaMemsetAsync(c.loss_acc.f64() + 1, 0, sizeof(double), c.stream));
    if (o.lambda > 0 && e > b) {
      const bool weighted = (o.feedback == B200ALS_EXPLICIT) && o.dynamic_lambda;
      const int grid = c.sm_count * 2;
      CU(c.reg_partials.ensure(sizeof(double) * (size_t)grid));
      sqnorm_kernel<float><<<grid, 256, 0, c.stream>>>(X + (size_t)b * s->k, s->k, e - b, weighted ? s->cnt[fixed].f32() + b : nullptr,
                                                     c.reg_partials.f64());
      LAUNCHED(); CU(cudaGetLastError());
      sum_partials_kernel<<<1, 32, 0, c.stream>>>(c.reg_partials.f64(), grid, c.loss_acc.f64() + 1, 0);
      LAUNCHED(); CU(cudaGetLastError());
    }
    status_to_double_kernel<<<1, 1, 0, c.stream>>>(c.status.i32(), c.loss_acc.f64() + 2);
    LAUNCHED(); CU(cudaGetLastError());
    NC(g_nccl.AllReduce(c.loss_acc.p, c.loss_acc.p, 3, ncclDouble, ncclSum, g_comm.comm, c.stream));
    double h[3] = {0, 0, 0};
    CU(cudaMemcpyAsync(h, c.loss_acc.p, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
    CU(cudaStreamSynchronize(c.stream));
    if (h[2] != 0.0) return fail(B200ALS_ENOTSPD, "a per-row system was not positive definite (Cholesky pivot <= 0) on some rank");
    if (loss) *loss = (double)(float)((h[0] + o.lambda * h[1]) / (double)s->nnz_global[which]);
  } else {
    TRY(finish_loss<float>(c, X, s->k, n_fixed, s->cnt[fixed].f32(), o, s->nnz_global[which], 0.0, false, loss));
  }
  cudaEventElapsedTime(&s->t_gram, s->ev[0], s->ev[1]);
  cudaEventElapsedTime(&s->t_prep, s->ev[1], s->ev[2]);
  cudaEventElapsedTime(&s->t_solve, s->ev[2], s->ev[3]);
  cudaEventElapsedTime(&s->t_comm, s->ev[3], s->ev[4]);
  return B200ALS_OK;
}

extern "C" int b200als_half_iteration(b200als_session* s, int which, int solver_override, double* loss) {
  if (!s || which < 0 || which > 1) return fail(B200ALS_EINVAL, "bad argument");
  return session_half(s, which, solver_override >= 0 ? solver_override : s->opt.solver, nullptr, loss);
}

extern "C" int b200als_fit(b200als_session* s, int n_iter, double convergence_tol, double* loss_trace, int* n_iter_done) {
  if (!s || n_iter < 0) return fail(B200ALS_EINVAL, "bad argument");
  double loss_prev = INFINITY;
  int done = 0;
  for (int i = 0; i < n_iter; i++) {  // R/model_WRMF.R:318-338
    double li = 0, lu = 0;
    TRY(session_half(s, B200ALS_ITEMS, s->opt.solver, nullptr, &li));
    TRY(session_half(s, B200ALS_USERS, s->opt.solver, nullptr, &lu));
    if (loss_trace) { loss_trace[2 * i] = li; loss_trace[2 * i + 1] = lu; }
    done = i + 1;
    if (loss_prev / lu - 1 < convergence_tol) break;
    loss_prev = lu;
  }
  if (n_iter_done) *n_iter_done = done;
  return B200ALS_OK;
}

extern "C" int b200als_transform(b200als_session* s, float* host_out, double* loss) {
  Ctx& c = ctx();
  if (!s || !host_out) return fail(B200ALS_EINVAL, "bad argument");
  if (!s->has[B200ALS_USERS]) return fail(B200ALS_EINVAL, "transform needs the users orientation");
  CscDev<float>& A = s->csc[B200ALS_USERS];
  DevBuf res;
  const size_t bytes = sizeof(float) * (size_t)s->k * (size_t)A.n_cols;
  CU(res.ensure(bytes));
  CU(cudaMemsetAsync(res.p, 0, bytes, c.stream));  // res = zeros (R/model_WRMF.R:423-427)
  if (s->with_biases && A.n_cols > 0) {            // res[1, ] = 1 (R/model_WRMF.R:429-431)
    set_col_kernel<<<(A.n_cols + 255) / 256, 256, 0, c.stream>>>(res.f32(), s->k, 0, A.n_cols, 1.0f);
    LAUNCHED(); CU(cudaGetLastError());
  }
  const int solver = (s->opt.solver == B200ALS_CONJUGATE_GRADIENT) ? B200ALS_CHOLESKY : s->opt.solver;  // avoid_cg (:112)
  TRY(session_half(s, B200ALS_USERS, solver, res.f32(), loss));
  return export_rotated(s, res.f32(), A.n_cols, host_out);
}

extern "C" int b200als_set_bias(b200als_session* s, int with_user_item_bias, double global_bias) {
  if (!s) return fail(B200ALS_EINVAL, "null session");
  if (with_user_item_bias && s->k < 3) return fail(B200ALS_EINVAL, "with_user_item_bias needs factor matrices of rank + 2 >= 3 rows");
  s->with_biases = with_user_item_bias ? 1 : 0;
  s->global_bias = global_bias;
  return B200ALS_OK;
}

extern "C" int b200als_row_plan(b200als_session* s, int which, int32_t counts[10], int32_t caps[9], int64_t* nnz_local) {
  if (!s || which < 0 || which > 1 || !s->has[which]) return fail(B200ALS_EINVAL, "orientation not present");
  const CscDev<float>& A = s->csc[which];
  for (int q = 0; q < CscDev<float>::kNumCls; q++) {
    if (counts) counts[q] = (A.plan_key >= 0) ? A.cls[q].count : 0;
    if (caps) caps[q] = (A.plan_key >= 0) ? A.cls[q].hi : 0;
  }
  if (counts) counts[9] = (A.plan_key >= 0) ? A.plan_empty : 0;
  if (nnz_local) *nnz_local = A.nnz;
  return B200ALS_OK;
}

extern "C" int b200als_last_timing(b200als_session* s, float* gram_ms, float* prep_ms, float* solve_ms, float* comm_ms) {
  if (!s) return fail(B200ALS_EINVAL, "null session");
  if (gram_ms) *gram_ms = s->t_gram;
  if (prep_ms) *prep_ms = s->t_prep;
  if (solve_ms) *solve_ms = s->t_solve;
  if (comm_ms) *comm_ms = s->t_comm;
  return B200ALS_OK;
}
