// als_cg_tile.cuh -- fixed-step CG half-iteration with the row's gathered factor tile RESIDENT IN SHARED MEMORY:
// the general fast path -- any rank k <= 256 with k % 4 == 0, any row length that fits the tile buffer (the host bins
// rows by length, engine_solve.inl), fp32, eigenbasis / full XtX / explicit feedback.  als_cg_resident_kernel
// (als_resident.cuh) stays the specialised kernel for rank 128 with rows of <= 80 entries; everything it cannot take
// lands here instead of on the streaming kernel (als_generic.cuh), which re-gathers the tile from L2 on every sweep.
//
// Reference semantics: cg_solver_implicit (inst/include/wrmf_implicit.hpp:8-32), cg_solver_explicit
// (inst/include/wrmf_explicit.hpp:8-31) and the column loops around them (wrmf_implicit.hpp:175-282,
// wrmf_explicit.hpp:71-146); same iterates, same `rsnew < CG_TOL` exit.
//
// Shape of the kernel:
//   * one CTA solves one row at a time (static interleave over a persistent grid => fixed loss summation order);
//     the row's tile X_nnz (n x k floats) is gathered from HBM ONCE by 16-byte cp.async into one of two tile
//     buffers -- the tile of the CTA's next row lands while this row computes; CSR indices / values travel
//     global -> shared two rows ahead, row pointers three rows ahead (registers): no dependent HBM chain per row;
//   * a gathered row of k floats is held by LPR = KPAD/4 (<= 32) lanes as one float4 each (two float4 per lane at
//     rank 256), so a warp works on 32/LPR gathered rows at a time and small ranks do not idle lanes;
//   * a CG step is one sweep over the tile: per batch of 4 row steps  u_j = x_j . v  (packed FFMA2 + a transposing
//     halving reduction: 3 + log2(LPR/4) shuffles for 4 x 32/LPR rows, select-free because every lane loads its
//     batch permuted by the slot it will own),  w_j = k0 c_j + (k1 c_j + k2) u_j  by the owner lanes, 4 broadcast
//     shuffles,  acc += w_j x_j  from the same registers; whole batches run without bounds checks (the padding
//     columns of the shared-memory tile are zero for the whole kernel); then one cross-warp sum through shared
//     memory (two-stage when the CTA has more than 4 warps);
//   * XtX p: d (.) p in the eigenbasis of XtX (eig.cuh), or (kFullG) every (warp, lane group) multiplies its slab
//     of XtX rows from L1/L2 and the slabs ride the same reduction; lambda_use p for explicit feedback;
//   * the loss term X_nnz' y comes from the u vectors already computed (X_nnz'y = X_nnz'x0 + sum_k alpha_k X_nnz'p_k).
//   * rows too long for one CTA's buffers (the item half-iteration; heavy-tailed data) are solved by a THREAD-BLOCK
//     CLUSTER of 2, 4 or 8 CTAs (kCluster): CTA r of the cluster stages and keeps the r-th contiguous slab of the row's
//     tile, every sweep ends with one cluster barrier after which each CTA adds the per-CTA partial sums of all slabs in
//     rank order through distributed shared memory (mapa + ld.shared::cluster) -- the tile is still read from HBM once.
//     All CTAs of a cluster carry the same CG state and take the same exits; CTA 0 writes the row and its lambda |y|^2.
// Algorithmic HBM bytes per row: 4nk + 8n + 4 + 4k + 4k.
#pragma once
#include "als_resident.cuh"   // packed-fp32 helpers (dot4, axpy4, fma4, ...)

namespace b200als {

struct TileCgParams {
  const int32_t* ptr;
  const int32_t* idx;
  const float* val;
  const float* X;      // k x n_src
  float* Y;            // k x n_targets
  const float* diag;   // [k] eigenvalues of XtX (+lambda): implicit, eigenbasis
  const float* G;      // k x k XtX + lambda I: implicit, kFullG
  int k;               // rank, k % 4 == 0, k <= KPAD of the instantiation
  int feedback;
  int cg_steps;
  int dynamic_lambda;
  float lambda;
  const int32_t* row_list;  // rows of this launch's length class (nullptr: rows row_begin .. row_begin + n_list - 1)
  int n_list;
  int ptr_base;
  int row_begin;
  int cap;                  // tile-buffer capacity in gathered rows (every listed row has 1 <= nnz <= cap)
  int nbuf;                 // 2: double-buffered tiles, 1: single buffer
  double* loss_partials;    // [gridDim.x]
};

constexpr int kTileBatch = 4;   // row steps per reduction batch

// shared-memory carve-up (bytes), shared by host and device
struct TileCgLayout {
  int kpad, cap, warps, full_g;
  int nbuf = 2;      // tile buffers: 2 = the next row's tile lands while this one computes; 1 = single buffer (twice the capacity
                     // per byte of shared memory; the load of a row is then hidden by the OTHER CTAs of the SM)
  int cluster = 0;   // 1: the kernel runs on thread-block clusters (needs the per-CTA partial-sum buffer its peers read)
  __host__ __device__ size_t tile_off(int b) const { return (size_t)b * cap * kpad * 4; }
  __host__ __device__ size_t ybuf_off(int b) const { return tile_off(nbuf) + (size_t)b * kpad * 4; }
  // cross-warp exchange: double-buffered per sweep with <= 4 warps (one barrier per sweep); a single buffer with more
  // warps (the two-stage sum has a second barrier per sweep, which also separates consecutive sweeps)
  __host__ __device__ int n_vbuf() const { return warps > 4 ? 1 : 2; }
  __host__ __device__ size_t vbuf_off(int b) const { return ybuf_off(nbuf) + (size_t)b * warps * kpad * 4; }
  __host__ __device__ size_t sum_off() const { return vbuf_off(n_vbuf()); }                              // [2][kpad] two-stage cross-warp sum (> 4 warps)
  __host__ __device__ size_t csum_off() const { return sum_off() + (warps > 4 ? (size_t)2 * kpad * 4 : 0); }   // [2][kpad] this CTA's partial, read by its cluster peers
  __host__ __device__ size_t vec_off() const { return csum_off() + (cluster ? (size_t)2 * kpad * 4 : 0); }     // [warps][kpad], kFullG only
  __host__ __device__ size_t idx_off(int s) const { return vec_off() + (full_g ? (size_t)warps * kpad * 4 : 0) + (size_t)s * cap * 4; }
  __host__ __device__ size_t val_off(int s) const { return idx_off(3) + (size_t)s * cap * 4; }
  __host__ __device__ size_t ubuf_off(int b) const { return val_off(3) + (size_t)b * cap * 4; }
  __host__ __device__ size_t uy_off() const { return ubuf_off(2); }
  __host__ __device__ size_t red_off() const { return (uy_off() + (size_t)cap * 4 + 15) & ~(size_t)15; }
  __host__ __device__ size_t bytes() const { return red_off() + 32 * 8; }
};

// thread-block cluster primitives (PTX): rank / size of the cluster, barrier with release / acquire semantics over the
// cluster's shared memories, and a 4-byte load from the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(const float* local_ptr, uint32_t rank) {
  uint32_t ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}

template <int LPR, int C, bool kFullG, bool kCluster = false>
__global__ void __launch_bounds__(512) als_cg_tile_kernel(TileCgParams P) {
  static_assert(LPR == 4 || LPR == 8 || LPR == 16 || LPR == 32, "lanes per gathered row");
  static_assert(C == 1 || (C == 2 && LPR == 32), "two chunks per lane only at rank 256");
  constexpr int KPAD = LPR * 4 * C;
  constexpr int RPW = 32 / LPR;                  // gathered rows per warp per row step
  constexpr int LOG_LPR = (LPR == 4) ? 2 : (LPR == 8) ? 3 : (LPR == 16) ? 4 : 5;
  constexpr int SLOT_SHIFT = LOG_LPR - 2;        // slot of a lane = top two bits of its index within the group
  constexpr int B = kTileBatch;
  static_assert(B == 4, "the halving reduction below is written for batches of 4 row steps");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int W = blockDim.x >> 5, T = blockDim.x;
  const int gi = lane / LPR, gl = lane % LPR;    // lane group within the warp, lane within the group
  const int slot = gl >> SLOT_SHIFT;             // which of a batch's 4 row steps this lane owns after the reduction
  const int k = P.k;
  const TileCgLayout L{KPAD, P.cap, W, kFullG ? 1 : 0, P.nbuf, kCluster ? 1 : 0};
  const bool single = (P.nbuf == 1);
  auto tile_of = [&](int b) { return reinterpret_cast<float*>(smem_raw + L.tile_off(b)); };
  auto ybuf_of = [&](int b) { return reinterpret_cast<float*>(smem_raw + L.ybuf_off(b)); };
  auto vbuf_of = [&](int b) { return reinterpret_cast<float*>(smem_raw + L.vbuf_off(b)); };
  auto idx_of = [&](int s) { return reinterpret_cast<int*>(smem_raw + L.idx_off(s)); };
  auto val_of = [&](int s) { return reinterpret_cast<float*>(smem_raw + L.val_off(s)); };
  auto ubuf_of = [&](int b) { return reinterpret_cast<float*>(smem_raw + L.ubuf_off(b)); };
  float* uy = reinterpret_cast<float*>(smem_raw + L.uy_off());
  double* red = reinterpret_cast<double*>(smem_raw + L.red_off());
  float* sumv = reinterpret_cast<float*>(smem_raw + L.sum_off());                       // [2][KPAD] two-stage cross-warp sum
  float* vecw = reinterpret_cast<float*>(smem_raw + L.vec_off()) + (size_t)w * KPAD;   // kFullG: this warp's copy of v

  const bool implicit = (P.feedback == 0);
  // a cluster of CL CTAs works on one row; CTA `crank` of it owns the crank-th contiguous slab of the row's entries
  const int CL = kCluster ? (int)cluster_nctarank() : 1;
  const int crank = kCluster ? (int)cluster_ctarank() : 0;
  const int first = (int)blockIdx.x / CL;
  const int stride = (int)gridDim.x / CL;
  const long long n_list = P.n_list;
  float* csum = reinterpret_cast<float*>(smem_raw + L.csum_off());
  auto valid = [&](int i) -> bool { return (long long)first + (long long)i * stride < n_list; };
  auto row_of = [&](int i) -> int {
    const long long t = (long long)first + (long long)i * stride;
    return P.row_list ? __ldg(P.row_list + t) : (int)t + P.row_begin;
  };
  // this CTA's slab of a row with n entries starting at p: [p + lo, p + lo + cnt)
  auto slab_lo = [&](int n) -> int { const int sl = (n + CL - 1) / CL; return min(n, crank * sl); };
  auto slab_cnt = [&](int n) -> int { const int sl = (n + CL - 1) / CL; const int lo = min(n, crank * sl); return min(n, lo + sl) - lo; };
  // which features this lane holds: chunk c covers [128 c + 4 gl, +4); beyond k (k % 4 == 0) the shared-memory copies
  // stay zero for the whole kernel (zero-filled below, never written by a copy), so only global accesses are guarded
  bool fvalid[C];
  int foff[C];
#pragma unroll
  for (int c = 0; c < C; c++) { foff[c] = c * 128 + 4 * gl; fvalid[c] = foff[c] < k; }
  for (size_t e = (size_t)tid * 16; e < L.vec_off(); e += (size_t)T * 16)   // tiles, y buffers, exchange buffers
    *reinterpret_cast<float4*>(smem_raw + e) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  // ---- staging helpers ------------------------------------------------------------------------------------
  auto issue_meta = [&](int sl, int p, int n) {          // CSR indices / values of a row -> shared (4-byte cp.async)
    int* si = idx_of(sl);
    float* sv = val_of(sl);
    for (int j = tid; j < n; j += T) {
      cp_async_4(si + j, P.idx + p + j);
      cp_async_4(sv + j, P.val + p + j);
    }
  };
  auto issue_tile = [&](int buf, int sl, int row, int n) {   // gathered rows + warm-start y -> shared (16-byte cp.async)
    // A warp instruction copies CW consecutive 16-byte chunks of RP gathered rows: one index load and one 64-bit
    // address per (row, lane) instead of per chunk (the per-chunk loop was 24 % of the kernel's instructions at rank 256).
    const int* si = idx_of(sl);
    float* tl = tile_of(buf);
    constexpr int CPR = KPAD / 4;                            // 16-byte chunks per padded row (power of two)
    constexpr int CW = (CPR < 32) ? CPR : 32;                // chunks of one row per warp instruction
    constexpr int RP = 32 / CW;                              // rows per warp instruction
    constexpr int NC = CPR / CW;                             // instructions per row (2 at rank > 128)
    const int cl = lane % CW;
#pragma unroll 4
    for (int j = w * RP + lane / CW; j < n; j += W * RP) {
      const float* src = P.X + (size_t)si[j] * k + 4 * cl;
      float* dst = tl + (size_t)j * KPAD + 4 * cl;
#pragma unroll
      for (int cc = 0; cc < NC; cc++)
        if (4 * (cl + cc * CW) < k) cp_async_16(dst + cc * CW * 4, src + cc * CW * 4);
    }
    constexpr int CPRy = CPR;
    if (tid < CPRy && 4 * tid < k) cp_async_16(ybuf_of(buf) + 4 * tid, P.Y + (size_t)row * k + 4 * tid);
  };

  // ---- pipeline prologue ----------------------------------------------------------------------------------
  //   row i   : rid0, n0            tile landing / landed in buffer i & 1, metadata in slot i % 3
  //   row i+1 : rid1, n1            metadata landed (slot (i+1) % 3); tile issued at the start of row i
  //   row i+2 : rid2, p2, n2        metadata issued at the start of row i
  //   row i+3 : rid3                row pointers loaded during row i
  int rid0 = -1, rid1 = -1, rid2 = -1, rid3 = -1;
  int n0 = 0, n1 = 0, n2 = 0, p2 = 0;
  if (valid(0)) {
    rid0 = row_of(0);
    const int p = __ldg(P.ptr + rid0) - P.ptr_base;
    n0 = __ldg(P.ptr + rid0 + 1) - P.ptr_base - p;
    issue_meta(0, p + slab_lo(n0), slab_cnt(n0));
  }
  if (valid(1)) {
    rid1 = row_of(1);
    const int p = __ldg(P.ptr + rid1) - P.ptr_base;
    n1 = __ldg(P.ptr + rid1 + 1) - P.ptr_base - p;
    issue_meta(1, p + slab_lo(n1), slab_cnt(n1));
  }
  if (valid(2)) { rid2 = row_of(2); p2 = __ldg(P.ptr + rid2) - P.ptr_base; n2 = __ldg(P.ptr + rid2 + 1) - P.ptr_base - p2; }
  if (valid(3)) rid3 = row_of(3);
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  if (valid(0) && !single) issue_tile(0, 0, rid0, slab_cnt(n0));

  float4 dg[C];
#pragma unroll
  for (int c = 0; c < C; c++)
    dg[c] = (!kFullG && implicit && fvalid[c]) ? ldg_f4(P.diag + foff[c]) : make_float4(0.f, 0.f, 0.f, 0.f);
  double lane_loss = 0.0;
  int sweep = 0;

  // dot over the k features of two replicated vectors (every lane of the CTA ends with the same value)
  auto vdot = [&](const float4 (&a)[C], const float4 (&b)[C]) -> float {
    float s = dot4(a[0], b[0]);
    if constexpr (C == 2) s += dot4(a[1], b[1]);
#pragma unroll
    for (int m = LPR / 2; m > 0; m >>= 1) s += __shfl_xor_sync(kFull, s, m);
    return s;
  };

  for (int i = 0; valid(i); i++) {
    const int n_row = n0;              // entries of the whole row (lambda_use)
    const int n = slab_cnt(n0);        // entries of this CTA's slab
    const int buf = single ? 0 : (i & 1), sl = i % 3;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();   // tile i and the metadata of row i+1 have landed; every warp is done with row i-1
    if (single && i == 0) {   // one buffer: the first row's tile is fetched now; later ones right after the previous row's last sweep
      issue_tile(0, sl, rid0, n);
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();
    }
    // ---- prefetch (nothing here waits on memory) ----
    if (valid(i + 1) && !single) issue_tile(buf ^ 1, (i + 1) % 3, rid1, slab_cnt(n1));
    if (valid(i + 2)) issue_meta((i + 2) % 3, p2 + slab_lo(n2), slab_cnt(n2));
    int p3 = 0, p3e = 0, rid4 = -1;
    if (valid(i + 3)) { p3 = ld_pinned_i32(P.ptr + rid3); p3e = ld_pinned_i32(P.ptr + rid3 + 1); }
    if (valid(i + 4)) rid4 = P.row_list ? ld_pinned_i32(P.row_list + ((long long)first + (long long)(i + 4) * stride))
                                        : row_of(i + 4);

    const float* sv = val_of(sl);
    const float lam_use = implicit ? P.lambda : (P.lambda * (P.dynamic_lambda ? (float)n_row : 1.0f));
    // Row steps: step rs of warp w covers gathered rows (rs * W + w) * RPW + gi.  Register s of a batch holds row step
    // b0 + (s ^ slot): lanes of different slots keep their batch permuted, which makes both halving levels of the
    // reduction the same instruction stream for every lane (keep registers 0/1, send 2/3; keep 0, send 1) -- no selects.
    const int n_steps_total = (n + RPW - 1) / RPW;                        // over the whole CTA
    const int my_steps = (n_steps_total > w) ? (n_steps_total - w + W - 1) / W : 0;
    const int full_steps_total = n / RPW;                                 // row steps whose RPW rows all exist
    const int my_full = (full_steps_total > w) ? (full_steps_total - w + W - 1) / W : 0;
    const int nb_full = my_full / B;                                      // batches that need no bounds checks
    const int step_stride = W * RPW * KPAD;                               // floats between consecutive row steps of a warp
    const float* lane_tile = tile_of(buf) + (size_t)(w * RPW + gi) * KPAD + foff[0];

    // One sweep: acc = sum_j w_j x_j with w_j = k0 c_j + (k1 c_j + k2) (x_j . v)  (+ / - XtX v when kFullG), summed
    // over the CTA; u_j = x_j . v -> ubuf.
    //   implicit r0: c - (c-1) u  (1,-1, 1)     implicit Ap: (c-1) u  (0, 1,-1)
    //   explicit r0: c - u        (1, 0,-1)     explicit Ap: u        (0, 0, 1)
    auto run_sweep = [&](const float4 (&v)[C], float k0, float k1, float k2, int gmode, float4 (&out)[C]) {
      float* ub = ubuf_of(sweep & 1);
      float4 acc[C];
#pragma unroll
      for (int c = 0; c < C; c++) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      auto batch = [&](int b0, bool checked) {
        float4 x[B][C];
        float t[B];
#pragma unroll
        for (int s = 0; s < B; s++) {
          const int rs = b0 + (s ^ slot);
          const float* src = lane_tile + (size_t)rs * step_stride;
          bool ok = true;
          if (checked) ok = (rs < my_steps) && ((rs * W + w) * RPW + gi < n);
#pragma unroll
          for (int c = 0; c < C; c++)
            x[s][c] = ok ? *reinterpret_cast<const float4*>(src + c * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
          t[s] = dot4(x[s][0], v[0]);
          if constexpr (C == 2) t[s] += dot4(x[s][1], v[1]);
        }
        // transposing halving reduction over the LPR lanes of the group, then plain butterflies
        const float o0 = t[0] + __shfl_xor_sync(kFull, t[2], LPR / 2);
        const float o1 = t[1] + __shfl_xor_sync(kFull, t[3], LPR / 2);
        float u = o0 + __shfl_xor_sync(kFull, o1, LPR / 4);
#pragma unroll
        for (int m = LPR / 8; m > 0; m >>= 1) u += __shfl_xor_sync(kFull, u, m);
        // u = x_j . v of this lane's own row step (register 0 = step b0 + slot)
        const int mrs = b0 + slot;
        const int mj = (mrs * W + w) * RPW + gi;
        float wq = 0.0f;
        if (!checked || ((mrs < my_steps) && (mj < n))) {
          const float cj = sv[mj];
          wq = fmaf(fmaf(k1, cj, k2), u, k0 * cj);
          if ((gl & ((1 << SLOT_SHIFT) - 1)) == 0) ub[mj] = u;
        }
#pragma unroll
        for (int s = 0; s < B; s++) {
          const float ws = __shfl_sync(kFull, wq, gi * LPR + ((s ^ slot) << SLOT_SHIFT));
#pragma unroll
          for (int c = 0; c < C; c++) acc[c] = axpy4(ws, x[s][c], acc[c]);
        }
      };
      int b0 = 0;
      for (int nb = 0; nb < nb_full; nb++, b0 += B) batch(b0, false);
      for (; b0 < my_steps; b0 += B) batch(b0, true);
      if constexpr (kFullG) {
        if (gmode != 0) {
          // this (warp, group)'s slab of XtX rows: j = unit, unit + U, ...; v_j from this warp's shared copy of v
#pragma unroll
          for (int c = 0; c < C; c++)
            if (gi == 0) *reinterpret_cast<float4*>(vecw + foff[c]) = v[c];
          __syncwarp();
          float4 g[C];
#pragma unroll
          for (int c = 0; c < C; c++) g[c] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int U = CL * W * RPW;   // the rows of XtX are split over every (CTA of the cluster, warp, lane group)
          for (int j = (crank * W + w) * RPW + gi; j < k; j += U) {
            const float vj = vecw[j];
#pragma unroll
            for (int c = 0; c < C; c++)
              if (fvalid[c]) g[c] = axpy4(vj, ldg_f4(P.G + (size_t)j * k + foff[c]), g[c]);
          }
          const float sg = (gmode == 1) ? -1.0f : 1.0f;
#pragma unroll
          for (int c = 0; c < C; c++) acc[c] = axpy4(sg, g[c], acc[c]);
          __syncwarp();
        }
      }
      // sum over the lane groups of the warp, then over the warps (fixed order)
#pragma unroll
      for (int m = LPR; m < 32; m <<= 1) {
#pragma unroll
        for (int c = 0; c < C; c++) {
          acc[c].x += __shfl_xor_sync(kFull, acc[c].x, m);
          acc[c].y += __shfl_xor_sync(kFull, acc[c].y, m);
          acc[c].z += __shfl_xor_sync(kFull, acc[c].z, m);
          acc[c].w += __shfl_xor_sync(kFull, acc[c].w, m);
        }
      }
      float* vb = vbuf_of((W > 4) ? 0 : (sweep & 1));
      if (gi == 0) {
#pragma unroll
        for (int c = 0; c < C; c++) *reinterpret_cast<float4*>(vb + (size_t)w * KPAD + foff[c]) = acc[c];
      }
      __syncthreads();
      if (W <= 4) {
        // few warps: every lane adds the W partials of its own features
#pragma unroll
        for (int c = 0; c < C; c++) {
          float4 s4 = *reinterpret_cast<const float4*>(vb + foff[c]);
          for (int ww = 1; ww < W; ww++) s4 = add4(s4, *reinterpret_cast<const float4*>(vb + (size_t)ww * KPAD + foff[c]));
          out[c] = s4;
        }
      } else {
        // many warps: thread f < KPAD adds the W partials of feature f (conflict-free 4-byte reads), the sums are
        // published once and read back by everyone -- W + 1 reads per thread instead of W 16-byte reads per lane
        float* sb = sumv + (sweep & 1) * KPAD;
        if constexpr (kCluster) {
          // this CTA's slab sum -> csum; after the cluster barrier every CTA adds the CL slab sums in rank order
          float* cs = csum + (sweep & 1) * KPAD;
          if (tid < KPAD) {
            float s1 = vb[tid];
            for (int ww = 1; ww < W; ww++) s1 += vb[(size_t)ww * KPAD + tid];
            cs[tid] = s1;
          }
          cluster_sync_all();
          if (tid < KPAD) {
            float s2 = ld_dsmem_f32(cs + tid, 0);
            for (int rr = 1; rr < CL; rr++) s2 += ld_dsmem_f32(cs + tid, (uint32_t)rr);
            sb[tid] = s2;
          }
        } else if (tid < KPAD) {
          float s1 = vb[tid];
          for (int ww = 1; ww < W; ww++) s1 += vb[(size_t)ww * KPAD + tid];
          sb[tid] = s1;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < C; c++) out[c] = *reinterpret_cast<const float4*>(sb + foff[c]);
      }
      sweep++;
    };

    // ---- CG (cg_solver_implicit / cg_solver_explicit) ---------------------------------------------------------
    float4 x[C], r[C], p[C], v[C], Ap[C];
#pragma unroll
    for (int c = 0; c < C; c++) x[c] = *reinterpret_cast<const float4*>(ybuf_of(buf) + foff[c]);
    run_sweep(x, 1.0f, implicit ? -1.0f : 0.0f, implicit ? 1.0f : -1.0f, (kFullG && implicit) ? 1 : 0, v);
#pragma unroll
    for (int c = 0; c < C; c++) {
      if (implicit) r[c] = kFullG ? v[c] : fma4(neg4(dg[c]), x[c], v[c]);     // v - d (.) x   (wrmf_implicit.hpp:16)
      else r[c] = axpy4(-lam_use, x[c], v[c]);                                // v - lambda_use x (wrmf_explicit.hpp:15)
      p[c] = r[c];
    }
    // X_nnz' y for the loss starts as X_nnz' x0 (the u of the first sweep)
    {
      const float* ub = ubuf_of((sweep - 1) & 1);
      for (int j = tid; j < n; j += T) uy[j] = ub[j];
    }
    float rsold = vdot(r, r);
    // guard the reference lacks (rsold / p'Ap = 0/0 once a row has converged exactly): a zero residual skips the loop
    const int n_cg = (rsold > 0.0f) ? P.cg_steps : 0;
    for (int it = 0; it < n_cg; it++) {
      run_sweep(p, 0.0f, implicit ? 1.0f : 0.0f, implicit ? -1.0f : 1.0f, (kFullG && implicit) ? 2 : 0, v);
#pragma unroll
      for (int c = 0; c < C; c++) {
        if (implicit) Ap[c] = kFullG ? v[c] : fma4(dg[c], p[c], v[c]);        // XtX p + X_nnz((c-1) . X_nnz'p)  (:22)
        else Ap[c] = axpy4(lam_use, p[c], v[c]);                              // X_nnz X_nnz' p + lambda p       (:21)
      }
      const float pAp = vdot(p, Ap);
      const float a = (pAp != 0.0f) ? __fdividef(rsold, pAp) : 0.0f;   // hardware reciprocal, <= 2 ulp (as in als_resident.cuh)
#pragma unroll
      for (int c = 0; c < C; c++) {
        x[c] = axpy4(a, p[c], x[c]);
        r[c] = axpy4(-a, Ap[c], r[c]);
      }
      {
        const float* ub = ubuf_of((sweep - 1) & 1);
        for (int j = tid; j < n; j += T) uy[j] = fmaf(a, ub[j], uy[j]);   // own entries only: no barrier needed
      }
      if (it + 1 == n_cg) break;
      const float rsnew = vdot(r, r);
      if (rsnew < (float)B200ALS_CG_TOL) break;                          // identical in every warp
      const float bt = __fdividef(rsnew, rsold);
#pragma unroll
      for (int c = 0; c < C; c++) p[c] = axpy4(bt, p[c], r[c]);
      rsold = rsnew;
    }
    // One buffer: nobody reads the tile after the last sweep (every warp has passed its closing barrier), and the next
    // row's metadata landed a row ago -- its copies run under the store of y, the loss pass and the next row's set-up.
    if (single && valid(i + 1)) issue_tile(0, (i + 1) % 3, rid1, slab_cnt(n1));
    if (w == 0 && gi == 0 && crank == 0) {
#pragma unroll
      for (int c = 0; c < C; c++)
        if (fvalid[c]) *reinterpret_cast<float4*>(P.Y + (size_t)rid0 * k + foff[c]) = x[c];
    }
    // ---- loss (wrmf_implicit.hpp:259-261 / wrmf_explicit.hpp:131-132); every thread reads back its own uy entries ----
    {
      float l = 0.0f;
      for (int j = tid; j < n; j += T) {
        const float cj = sv[j];
        const float d = implicit ? (1.0f - uy[j]) : (cj - uy[j]);
        l += implicit ? d * d * cj : d * d;
      }
      // per-lane fp64 sums (one block reduction at the end of the kernel): lane group 0 of warp 0 adds lambda |y|^2
      if (w == 0 && gi == 0 && crank == 0) {
        float yy = dot4(x[0], x[0]);
        if constexpr (C == 2) yy += dot4(x[1], x[1]);
        l = fmaf(lam_use, yy, l);
      }
      lane_loss += (double)l;
    }
    // ---- advance the pipeline ----
    n0 = n1; rid0 = rid1;
    n1 = n2; rid1 = rid2;
    p2 = p3 - P.ptr_base; n2 = p3e - p3; rid2 = rid3;
    rid3 = rid4;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  if constexpr (kCluster) cluster_sync_all();   // no CTA leaves while a peer may still read its shared memory
  const double tot = block_sum_double(lane_loss, red);
  if (tid == 0) P.loss_partials[blockIdx.x] = tot;
}

}  // namespace b200als
