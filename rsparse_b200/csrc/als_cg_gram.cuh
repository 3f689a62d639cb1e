// als_cg_gram.cuh -- fixed-step CG half-iteration for LONG rows at rank 128 (the item half-iteration, heavy-tailed data):
// the row's k x k system matrix is formed explicitly on the tensor cores in ONE pass over the gathered tile, then the
// reference's CG steps run on that 128 x 128 matrix.
//
// Why: the sweep-based kernels (als_resident.cuh, als_cg_tile.cuh) read the tile once but execute ~37 instructions per
// gathered row and sweep on CUDA cores, five sweeps per row; that is fine while n <~ k, but for n >> k the same CG
// iterates are cheaper through  A = XtX + X_nnz diag(c-1) X_nnz'  (explicit: X_nnz X_nnz' + lambda_u I),  b = X_nnz c:
//     r0 = b - A x0,   Ap = A p                       (wrmf_implicit.hpp:16,22 / wrmf_explicit.hpp:15,21, re-associated)
// -- 2 n k^2 flop on tcgen05 (3xTF32 split, fp32-grade) instead of 5 x 4 n k on FFMA, and no row-length limit (the
// streaming kernel moves the tile five times; measured HBM-bound on those re-reads).
//
// X_nnz diag(w) X_nnz' = Z Z' with z_j = sqrt(w_j) x_j (w_j = c_j - 1 >= 0, or 1 for explicit feedback): BOTH operands of
// the MMA are the same tile, so one hi/lo pair is staged per chunk.  The host enables the kernel for implicit feedback only
// when every confidence is >= 1 (engine_solve.inl); otherwise such rows take the cluster / streaming kernels.
//
// One CTA of 256 threads per row, two CTAs per SM (2 x 128 of the 512 TMEM columns), persistent, static row interleave:
//   phase A  chunks of 32 gathered rows travel global -> shared by 16-byte cp.async two chunks ahead (ring of two raw
//            buffers; CSR indices / values three chunks ahead); warp w transposes 4-row group w of the chunk into the
//            K-major core-matrix layout of gram_tc.cuh (hi / lo of z, double-buffered operand tiles) and accumulates its
//            share of b = sum_j c_j x_j in registers; one thread issues 3 tcgen05.mma.kind::tf32 (M128 N128 K8) per 8
//            gathered rows, tcgen05.commit -> mbarrier per tile buffer; every 4 chunks the accumulator is drained
//            (tcgen05.ld.32x32b) into registers with round-to-nearest adds (the tensor core accumulates with truncation:
//            short chains keep the bias ~1e-6);
//   phase B  threads (r, h), h = 0 / 1, own columns [64 h, 64 h + 64) of row r of A (64 registers) and a copy of element r of
//            x, r, p: mat-vecs read the vector from shared memory (broadcast LDS.128) and add the two half-row dots through
//            shared memory; the scalars of a CG step are block sums; same iterates, same `rsnew < CG_TOL` exit as
//            cg_solver_implicit / cg_solver_explicit;
//   phase C  loss: one more pass over the gathered rows (u_j = x_j . y from L2 / HBM, 16 rows in flight per warp),
//            wrmf_implicit.hpp:259-261.
// Algorithmic HBM bytes per row as the other CG kernels (4nk + 8n + 4 + 8k); this kernel moves the tile twice (phase C).
#pragma once
#include "als_cg_tile.cuh"   // TileCgParams
#include "rotate_tc.cuh"     // gram_tc.cuh: tc_smem_desc / tc_mma_tf32 / tc_commit / kTc* ; tf32_hi

namespace b200als {

constexpr int kGcThreads = 256;
constexpr int kGcWarps = kGcThreads / 32;
constexpr int kGcDrainChunks = 4;   // chunks of 32 gathered rows between two drains of the TMEM accumulator

struct alignas(128) GramCgSmem {
  unsigned char op[2][2][kTcTileBytes];    // [buffer][hi / lo] operand tiles of z = sqrt(w) x, 32 gathered rows each
  alignas(16) float raw[2][kTcRows][kTcK]; // gathered rows as they lie in HBM (cp.async ring)
  int idx[4][kTcRows];                     // CSR indices / values of four consecutive chunks (ring)
  float cs[4][kTcRows];
  alignas(16) float vec[2][kTcK];          // vector of the current mat-vec (double buffered)
  float half[2][2][kTcK];                  // the two half-row dots of a mat-vec (double buffered)
  float bpart[kGcWarps][kTcK];             // per-warp shares of b
  float part[2][kGcWarps];                 // per-warp partials of the block sums (double buffered)
  alignas(8) uint64_t mma_done[2];         // per operand-tile buffer
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kGcThreads, 2) als_cg_gram_kernel(TileCgParams P) {
  constexpr int K = kTcK;
  constexpr int KH = K / 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  GramCgSmem& S = *reinterpret_cast<GramCgSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int quad = warp & 3, h = warp >> 2;              // TMEM lane quadrant / column half of this warp
  const int r = quad * 32 + lane;                        // the row of A / the vector element this thread owns
  const bool implicit = (P.feedback == 0);
  const bool full_g = implicit && (P.diag == nullptr);
  if (tid == 0) {
    mbar_init(&S.mma_done[0], 1);
    mbar_init(&S.mma_done[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;
  for (int e = tid; e < 2 * kTcRows * K / 4; e += kGcThreads)   // raw ring: finite contents from the start (padding rows are read)
    reinterpret_cast<float4*>(&S.raw[0][0][0])[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  // per operand-tile buffer b (bit b): parity of the next completion to wait for / a commit on that buffer's mbarrier
  // has not been waited for yet (the same values in every thread; each completion is consumed exactly once)
  uint32_t ph_bits = 0, pend_bits = 0;
  auto wait_buf = [&](int b) {
    if (pend_bits & (1u << b)) {
      mbar_wait(&S.mma_done[b], (ph_bits >> b) & 1u);
      ph_bits ^= (1u << b);
      pend_bits &= ~(1u << b);
    }
  };
  const float dgr = (implicit && !full_g) ? __ldg(P.diag + r) : 0.0f;
  double cta_loss = 0.0;
  int sumbuf = 0;                   // parity of S.vec / S.half / S.part

  // block sum over the CTA of one float per (r, h = 0) thread (fixed order); every thread receives the result
  auto block_sum = [&](float v) -> float {
    v = warp_sum(h == 0 ? v : 0.0f);
    if (lane == 0) S.part[sumbuf][warp] = v;
    __syncthreads();
    const float t = ((S.part[sumbuf][0] + S.part[sumbuf][1]) + S.part[sumbuf][2]) + S.part[sumbuf][3];
    sumbuf ^= 1;
    return t;
  };

  const long long n_list = P.n_list;
  for (long long t = blockIdx.x; t < n_list; t += gridDim.x) {
    const int row = P.row_list ? __ldg(P.row_list + t) : (int)t + P.row_begin;
    const int p0 = __ldg(P.ptr + row) - P.ptr_base;
    const int n = __ldg(P.ptr + row + 1) - P.ptr_base - p0;
    const int n_chunks = (n + kTcRows - 1) / kTcRows;
    const float lam_use = implicit ? P.lambda : (P.lambda * (P.dynamic_lambda ? (float)n : 1.0f));
    const float x0r = __ldg(P.Y + (size_t)row * K + r);   // warm start (this thread's element), in flight during phase A

    auto issue_meta = [&](int c) {   // chunk c -> ring slot c % 4 (padding entries: index 0, value 0)
      if (tid < kTcRows) {
        const int j = c * kTcRows + tid;
        if (j < n) {
          cp_async_4(&S.idx[c & 3][tid], P.idx + p0 + j);
          cp_async_4(&S.cs[c & 3][tid], P.val + p0 + j);
        } else {
          S.idx[c & 3][tid] = 0;
          S.cs[c & 3][tid] = 0.0f;
        }
      }
    };
    // the 32 x 128 floats of chunk c -> raw[c & 1]: 1024 16-byte pieces, 4 per thread (a warp fetches whole 512-byte rows)
    auto issue_rows = [&](int c) {
#pragma unroll
      for (int i = 0; i < (kTcRows * K / 4) / kGcThreads; i++) {
        const int piece = tid + i * kGcThreads;
        const int jl = piece >> 5, c4 = piece & 31;
        if (c * kTcRows + jl < n)
          cp_async_16(&S.raw[c & 1][jl][c4 * 4], P.X + (size_t)S.idx[c & 3][jl] * K + c4 * 4);
      }
    };

    __syncthreads();   // the previous row's phases B / C are done with S.vec / S.half / S.part / the rings
    issue_meta(0);
    if (n_chunks > 1) issue_meta(1);
    if (n_chunks > 2) issue_meta(2);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    issue_rows(0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (n_chunks > 1) issue_rows(1);
    asm volatile("cp.async.commit_group;" ::: "memory");

    float acc[KH];   // columns [64 h, 64 h + 64) of row r of Z Z' (drained from TMEM)
#pragma unroll
    for (int c = 0; c < KH; c++) acc[c] = 0.0f;
    float bp[4] = {0.f, 0.f, 0.f, 0.f};   // this warp's share of b for features 4 lane .. 4 lane + 3

    for (int ch = 0; ch < n_chunks; ch++) {
      const int ob = ch & 1;          // operand-tile buffer and raw buffer of this chunk
      const int ms = ch & 3;          // metadata slot
      const int jbase = ch * kTcRows;
      // chunk ch has landed (the group committed after it may still be in flight); metadata of chunk ch + 2 as well
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncthreads();
      wait_buf(ob);                   // the MMAs that read op[ob] two chunks ago must have completed
      // ---- stage: warp w transposes the 4-row group w of the chunk into the K-major layout; z = s x, hi / lo split ------
      {
        const int kb = warp;
        // padding rows: their metadata is (index 0, value 0) => c = 0 and s = 0, and the raw buffer holds finite stale data
        // (zero-filled once per kernel), so no bounds checks are needed here
        const float4 c4v = *reinterpret_cast<const float4*>(&S.cs[ms][kb * 4]);
        const float cv[4] = {c4v.x, c4v.y, c4v.z, c4v.w};
        float sv[4];
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          // s = sqrt(c - 1) to ~2 ulp (MUFU.RSQ): s^2 differs from c - 1 by ~2e-7 relative, far below the fp32 tolerance
          const float wgt = implicit ? fmaxf(cv[kk] - 1.0f, 0.0f) : ((jbase + kb * 4 + kk < n) ? 1.0f : 0.0f);
          sv[kk] = (wgt > 0.0f) ? wgt * __frsqrt_rn(wgt) : 0.0f;
        }
        float4 v[4];
#pragma unroll
        for (int kk = 0; kk < 4; kk++) v[kk] = *reinterpret_cast<const float4*>(&S.raw[ob][kb * 4 + kk][lane * 4]);
        const float col[4][4] = {{v[0].x, v[1].x, v[2].x, v[3].x}, {v[0].y, v[1].y, v[2].y, v[3].y},
                                 {v[0].z, v[1].z, v[2].z, v[3].z}, {v[0].w, v[1].w, v[2].w, v[3].w}};
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
          const int m = 4 * lane + jj;   // feature
          const float x0 = col[jj][0], x1 = col[jj][1], x2 = col[jj][2], x3 = col[jj][3];
          bp[jj] = fmaf(cv[0], x0, fmaf(cv[1], x1, fmaf(cv[2], x2, fmaf(cv[3], x3, bp[jj]))));
          const float z0 = x0 * sv[0], z1 = x1 * sv[1], z2 = x2 * sv[2], z3 = x3 * sv[3];
          float4 zh, zl;
          zh.x = tf32_hi(z0); zl.x = z0 - zh.x;
          zh.y = tf32_hi(z1); zl.y = z1 - zh.y;
          zh.z = tf32_hi(z2); zl.z = z2 - zh.z;
          zh.w = tf32_hi(z3); zl.w = z3 - zh.w;
          const int off = (m >> 3) * kTcSBO + kb * kTcLBO + (m & 7) * 16;
          *reinterpret_cast<float4*>(&S.op[ob][0][off]) = zh;
          *reinterpret_cast<float4*>(&S.op[ob][1][off]) = zl;
        }
      }
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core's async proxy
      __syncthreads();       // tiles complete; raw[ob] and metadata slot ms have been read by everyone
      const bool window_start = (ch % kGcDrainChunks) == 0;
      const bool window_end = ((ch % kGcDrainChunks) == kGcDrainChunks - 1) || (ch == n_chunks - 1);
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int ksteps = min(kTcRows / 8, (n - jbase + 7) / 8);   // 8 gathered rows per MMA; padding rows are zero
        for (int ks = 0; ks < ksteps; ks++) {
          const uint64_t dh = tc_smem_desc(&S.op[ob][0][ks * 2 * kTcLBO]);
          const uint64_t dl = tc_smem_desc(&S.op[ob][1][ks * 2 * kTcLBO]);
          tc_mma_tf32(tmem, dh, dh, (window_start && ks == 0) ? 0u : 1u);   // z_hi' z_hi
          tc_mma_tf32(tmem, dh, dl, 1u);                                     // z_hi' z_lo
          tc_mma_tf32(tmem, dl, dh, 1u);                                     // z_lo' z_hi
        }
        tc_commit(&S.mma_done[ob]);
      }
      pend_bits |= (1u << ob);
      // ---- prefetch: metadata of chunk ch + 3 (its own group, committed FIRST: `wait_group 1` at the start of the next
      //      chunk then covers it), rows of chunk ch + 2 into the raw buffer just consumed ------------------------------
      if (ch + 3 < n_chunks) issue_meta(ch + 3);
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (ch + 2 < n_chunks) issue_rows(ch + 2);
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (window_end) {
        // every MMA of the window must have completed (older buffer first: each completion is consumed exactly once)
        wait_buf(ob ^ 1);
        wait_buf(ob);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < KH; c0 += 32) {
          uint32_t d[32];
          const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(h * KH + c0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]),
                "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]),
                "=r"(d[17]), "=r"(d[18]), "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]),
                "=r"(d[25]), "=r"(d[26]), "=r"(d[27]), "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
              : "r"(taddr)
              : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int c = 0; c < 32; c++) acc[c0 + c] += __uint_as_float(d[c]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        // after a drain both tile buffers are idle: nothing is pending on either mbarrier
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");

    // ---- b = sum over the warps' shares ------------------------------------------------------------------------------
    *reinterpret_cast<float4*>(&S.bpart[warp][4 * lane]) = make_float4(bp[0], bp[1], bp[2], bp[3]);
    __syncthreads();   // also: every thread has read the accumulator before the next row's first (overwriting) MMA
    float br = 0.0f;
#pragma unroll
    for (int w8 = 0; w8 < kGcWarps; w8++) br += S.bpart[w8][r];

    // ---- A = XtX (diag(d) in the eigenbasis) + Z Z'  /  X_nnz X_nnz' + lambda_u I ---------------------------------------
    if (full_g) {
#pragma unroll
      for (int c = 0; c < KH; c++) acc[c] += __ldg(P.G + (size_t)(h * KH + c) * K + r);   // symmetric: row r read as column r
    }
    const float dshift = implicit ? dgr : lam_use;   // on the diagonal (0 with the full XtX: it already carries lambda)
    // mat-vec with this thread's half row; the vector comes from shared memory (broadcast reads)
    auto matvec = [&](float vr) -> float {
      float* vb = S.vec[sumbuf];
      if (h == 0) vb[r] = vr;
      __syncthreads();
      float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < KH; c += 4) {
        const float4 v4 = *reinterpret_cast<const float4*>(vb + h * KH + c);
        s0 = __ffma2_rn(make_float2(acc[c], acc[c + 1]), make_float2(v4.x, v4.y), s0);
        s1 = __ffma2_rn(make_float2(acc[c + 2], acc[c + 3]), make_float2(v4.z, v4.w), s1);
      }
      S.half[sumbuf][h][r] = (s0.x + s0.y) + (s1.x + s1.y);
      __syncthreads();
      return fmaf(dshift, vr, S.half[sumbuf][0][r] + S.half[sumbuf][1][r]);
    };

    // ---- CG (cg_solver_implicit / cg_solver_explicit); both threads of a row carry the same x, r, p --------------------
    float xr = x0r;
    float rr = br - matvec(xr);
    float pr = rr;
    float rsold = block_sum(rr * rr);
    // guard the reference lacks (rsold / p'Ap = 0/0 once a row has converged exactly): a zero residual skips the loop
    const int n_cg = (rsold > 0.0f) ? P.cg_steps : 0;
    for (int it = 0; it < n_cg; it++) {
      const float Apr = matvec(pr);
      const float pAp = block_sum(pr * Apr);
      const float a = (pAp != 0.0f) ? __fdiv_rn(rsold, pAp) : 0.0f;
      xr = fmaf(a, pr, xr);
      rr = fmaf(-a, Apr, rr);
      if (it + 1 == n_cg) break;
      const float rsnew = block_sum(rr * rr);
      if (rsnew < (float)B200ALS_CG_TOL) break;          // identical in every thread
      pr = fmaf(__fdiv_rn(rsnew, rsold), pr, rr);
      rsold = rsnew;
    }
    if (h == 0) P.Y[(size_t)row * K + r] = xr;

    // ---- loss (wrmf_implicit.hpp:259-261 / wrmf_explicit.hpp:131-132): u_j = x_j . y, one more pass over the tile ----
    {
      float* vb = S.vec[sumbuf];
      if (h == 0) vb[r] = xr;
      __syncthreads();
      const float4 y4 = *reinterpret_cast<const float4*>(vb + 4 * lane);
      float l = 0.0f;
      // 16 gathered rows in flight per warp; their 16 dot products are reduced by a transposing halving reduction
      // (16 shuffles for 16 rows).  Register s of lane L holds row q + (s ^ slot), slot = L >> 1, which makes every halving
      // level the same instruction stream for all lanes; after the four levels + one butterfly lanes 2 slot and 2 slot + 1
      // hold u of row q + slot.
      const int slot = lane >> 1;
      for (int base = warp * 32; base < n; base += kGcWarps * 32) {    // this warp's groups of 32 gathered rows
        const int cnt = min(32, n - base);
        int my_idx = 0;
        float my_c = 0.0f;
        if (lane < cnt) {
          my_idx = __ldg(P.idx + p0 + base + lane);
          my_c = __ldg(P.val + p0 + base + lane);
        }
        for (int q = 0; q < cnt; q += 16) {
          float4 xv[16];
#pragma unroll
          for (int s16 = 0; s16 < 16; s16++) {
            const int jl = q + (s16 ^ slot);
            const int src = __shfl_sync(kFull, my_idx, jl & 31);
            xv[s16] = (jl < cnt) ? ldg_f4(P.X + (size_t)src * K + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          float t16[16];
#pragma unroll
          for (int s16 = 0; s16 < 16; s16++) t16[s16] = dot4(xv[s16], y4);
          float t8[8], t4[4], t2[2];
#pragma unroll
          for (int v = 0; v < 8; v++) t8[v] = t16[v] + __shfl_xor_sync(kFull, t16[v + 8], 16);
#pragma unroll
          for (int v = 0; v < 4; v++) t4[v] = t8[v] + __shfl_xor_sync(kFull, t8[v + 4], 8);
#pragma unroll
          for (int v = 0; v < 2; v++) t2[v] = t4[v] + __shfl_xor_sync(kFull, t4[v + 2], 4);
          float d = t2[0] + __shfl_xor_sync(kFull, t2[1], 2);
          d += __shfl_xor_sync(kFull, d, 1);
          const int jo = q + slot;
          const float cj = __shfl_sync(kFull, my_c, jo & 31);
          if (jo < cnt && (lane & 1) == 0) {
            const float e = implicit ? (1.0f - d) : (cj - d);
            l += implicit ? e * e * cj : e * e;
          }
        }
      }
      // per-warp sums -> CTA (all eight warps contribute here, unlike block_sum), then the regulariser term once per row
      l = warp_sum(l);
      if (lane == 0) S.part[sumbuf][warp] = l;
      __syncthreads();
      float tot = 0.0f;
#pragma unroll
      for (int w8 = 0; w8 < kGcWarps; w8++) tot += S.part[sumbuf][w8];
      sumbuf ^= 1;
      tot = fmaf(lam_use, block_sum(xr * xr), tot);
      if (tid == 0) cta_loss += (double)tot;
    }
  }
  if (tid == 0) P.loss_partials[blockIdx.x] = cta_loss;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

}  // namespace b200als
