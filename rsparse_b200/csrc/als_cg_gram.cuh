// als_cg_gram.cuh -- fixed-step CG half-iteration for LONG rows at rank 128 (the item half-iteration, heavy-tailed data):
// the row's k x k system matrix is formed explicitly on the tensor cores in ONE pass over the gathered tile, then the
// reference's CG steps run on that 128 x 128 matrix.
//
// Why: the sweep-based kernels (als_resident.cuh, als_cg_tile.cuh) read the tile once but execute ~37 instructions per
// gathered row and sweep on CUDA cores, five sweeps per row; that is fine while n <~ k, but for n >> k the same CG
// iterates are cheaper through  A = XtX + X_nnz diag(c-1) X_nnz'  (explicit: X_nnz X_nnz' + lambda_u I),  b = X_nnz c:
//     r0 = b - A x0,   Ap = A p                       (wrmf_implicit.hpp:16,22 / wrmf_explicit.hpp:15,21, re-associated)
// -- 2 n k^2 flop on tcgen05 (3xTF32 split, fp32-grade) instead of 5 x 4 n k on FFMA, ~11 instructions per gathered row,
// and no row-length limit (the streaming kernel moves the tile five times; measured HBM-bound on those re-reads).
//
// One CTA of 128 threads per row, two CTAs per SM (2 x 128 of the 512 TMEM columns), persistent, static row interleave:
//   phase A  chunks of 32 gathered rows: global -> registers (the next chunk's rows are requested before this chunk is
//            staged) -> transposed hi/lo operand tiles in the K-major core-matrix layout of gram_tc.cuh ((c-1) x for the M
//            side, x for the N side); one thread issues 3 tcgen05.mma.kind::tf32 (M128 N128 K8) per 8 gathered rows into a
//            TMEM accumulator, tcgen05.commit -> mbarrier per chunk; b += c_j x_j from the N-side tiles while the MMAs run;
//            every 4 chunks the accumulator is drained (tcgen05.ld.32x32b: thread r receives row r) into registers with
//            round-to-nearest adds (the tensor core accumulates with truncation: short chains keep the bias ~1e-6);
//   phase B  thread r owns row r of A (128 registers) and element r of x, r, p: mat-vecs read the vector from shared
//            memory (broadcast LDS.128), the three scalars of a CG step are block sums; same iterates, same
//            `rsnew < CG_TOL` exit as cg_solver_implicit / cg_solver_explicit;
//   phase C  loss: one more pass over the gathered rows (u_j = x_j . y from L2 / HBM), wrmf_implicit.hpp:259-261.
// Algorithmic HBM bytes per row as the other CG kernels (4nk + 8n + 4 + 8k); this kernel moves the tile twice (phase C).
#pragma once
#include "als_cg_tile.cuh"   // TileCgParams
#include "rotate_tc.cuh"     // gram_tc.cuh: tc_smem_desc / tc_mma_tf32 / tc_commit / kTc* ; tf32_hi

namespace b200als {

constexpr int kGcThreads = 128;
constexpr int kGcDrainChunks = 4;   // chunks of 32 gathered rows between two drains of the TMEM accumulator

struct alignas(128) GramCgSmem {
  unsigned char op[4][kTcTileBytes];       // operand tiles: (w x) hi, (w x) lo, x hi, x lo -- 32 gathered rows each
  int idx[3][kTcRows];                     // CSR indices / values of three consecutive chunks (ring)
  float cs[3][kTcRows];
  alignas(16) float vec[2][kTcK];          // vector of the current mat-vec (double buffered)
  float part[2][4];                        // per-warp partials of the block sums (double buffered)
  alignas(8) double red[32];
  alignas(8) uint64_t mma_done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kGcThreads, 2) als_cg_gram_kernel(TileCgParams P) {
  constexpr int K = kTcK;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  GramCgSmem& S = *reinterpret_cast<GramCgSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = tid;                                   // the row of A / the vector element this thread owns
  const bool implicit = (P.feedback == 0);
  const bool full_g = implicit && (P.diag == nullptr);
  if (tid == 0) {
    mbar_init(&S.mma_done, 1);
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
  uint32_t mma_phase = 0;
  const float dgr = (implicit && !full_g) ? __ldg(P.diag + r) : 0.0f;
  double cta_loss = 0.0;
  int sumbuf = 0;   // parity of S.vec / S.part

  // block sum of one float per thread (fixed order); every thread receives the result
  auto block_sum = [&](float v) -> float {
    v = warp_sum(v);
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
    // warm start (this thread's element) -- in flight during phase A
    const float x0r = __ldg(P.Y + (size_t)row * K + r);

    auto issue_meta = [&](int c) {   // chunk c -> ring slot c % 3 (padding entries: index 0, value 0)
      if (tid < kTcRows) {
        const int j = c * kTcRows + tid;
        if (j < n) {
          cp_async_4(&S.idx[c % 3][tid], P.idx + p0 + j);
          cp_async_4(&S.cs[c % 3][tid], P.val + p0 + j);
        } else {
          S.idx[c % 3][tid] = 0;
          S.cs[c % 3][tid] = 0.0f;
        }
      }
    };
    // this warp's two 4-row groups {warp, warp + 4} of chunk c: 8 x LDG.128 per lane (features 4 lane .. 4 lane + 3)
    auto load_rows = [&](int c, float4 (&v)[2][4]) {
#pragma unroll
      for (int g = 0; g < 2; g++)
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          const int jl = (warp + 4 * g) * 4 + kk;
          v[g][kk] = (c * kTcRows + jl < n) ? ldg_f4(P.X + (size_t)S.idx[c % 3][jl] * K + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };

    __syncthreads();   // the previous row's phases B / C are done with S.vec / S.part / the metadata ring
    issue_meta(0);
    if (n_chunks > 1) issue_meta(1);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    float4 nxt[2][4];
    load_rows(0, nxt);

    float acc[K];   // row r of X_nnz diag(w) X_nnz' (drained from TMEM)
#pragma unroll
    for (int c = 0; c < K; c++) acc[c] = 0.0f;
    float br = 0.0f;

    for (int ch = 0; ch < n_chunks; ch++) {
      float4 cur[2][4];
#pragma unroll
      for (int g = 0; g < 2; g++)
#pragma unroll
        for (int kk = 0; kk < 4; kk++) cur[g][kk] = nxt[g][kk];
      if (ch + 2 < n_chunks) issue_meta(ch + 2);
      if (ch + 1 < n_chunks) load_rows(ch + 1, nxt);   // metadata of chunk ch + 1 landed before the barrier that ended chunk ch - 1
      const int slot = ch % 3;
      const int jbase = ch * kTcRows;
      // ---- stage (cf. gram_tc_blocks_kernel / als_chol_rows_kernel): transpose to K-major, hi / lo split --------------
#pragma unroll
      for (int g = 0; g < 2; g++) {
        const int kb = warp + 4 * g;
        float wv[4];
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          const int jl = kb * 4 + kk;
          wv[kk] = (jbase + jl < n) ? (implicit ? (S.cs[slot][jl] - 1.0f) : 1.0f) : 0.0f;
        }
        const float4 (&v)[4] = cur[g];
        const float col[4][4] = {{v[0].x, v[1].x, v[2].x, v[3].x}, {v[0].y, v[1].y, v[2].y, v[3].y},
                                 {v[0].z, v[1].z, v[2].z, v[3].z}, {v[0].w, v[1].w, v[2].w, v[3].w}};
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
          const int m = 4 * lane + jj;   // feature
          float4 bh, bl, ah, al;
          const float x0 = col[jj][0], x1 = col[jj][1], x2 = col[jj][2], x3 = col[jj][3];
          const float y0 = x0 * wv[0], y1 = x1 * wv[1], y2 = x2 * wv[2], y3 = x3 * wv[3];
          bh.x = tf32_hi(x0); bl.x = x0 - bh.x;
          bh.y = tf32_hi(x1); bl.y = x1 - bh.y;
          bh.z = tf32_hi(x2); bl.z = x2 - bh.z;
          bh.w = tf32_hi(x3); bl.w = x3 - bh.w;
          ah.x = tf32_hi(y0); al.x = y0 - ah.x;
          ah.y = tf32_hi(y1); al.y = y1 - ah.y;
          ah.z = tf32_hi(y2); al.z = y2 - ah.z;
          ah.w = tf32_hi(y3); al.w = y3 - ah.w;
          const int off = (m >> 3) * kTcSBO + kb * kTcLBO + (m & 7) * 16;
          *reinterpret_cast<float4*>(&S.op[0][off]) = ah;
          *reinterpret_cast<float4*>(&S.op[1][off]) = al;
          *reinterpret_cast<float4*>(&S.op[2][off]) = bh;
          *reinterpret_cast<float4*>(&S.op[3][off]) = bl;
        }
      }
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core's async proxy
      __syncthreads();
      const bool window_start = (ch % kGcDrainChunks) == 0;
      const bool window_end = ((ch % kGcDrainChunks) == kGcDrainChunks - 1) || (ch == n_chunks - 1);
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int ksteps = min(kTcRows / 8, (n - jbase + 7) / 8);   // 8 gathered rows per MMA; padding rows are zero
        for (int ks = 0; ks < ksteps; ks++) {
          const uint64_t dah = tc_smem_desc(&S.op[0][ks * 2 * kTcLBO]);
          const uint64_t dal = tc_smem_desc(&S.op[1][ks * 2 * kTcLBO]);
          const uint64_t dbh = tc_smem_desc(&S.op[2][ks * 2 * kTcLBO]);
          const uint64_t dbl = tc_smem_desc(&S.op[3][ks * 2 * kTcLBO]);
          tc_mma_tf32(tmem, dah, dbh, (window_start && ks == 0) ? 0u : 1u);   // (w x)_hi' x_hi
          tc_mma_tf32(tmem, dah, dbl, 1u);                                     // (w x)_hi' x_lo
          tc_mma_tf32(tmem, dal, dbh, 1u);                                     // (w x)_lo' x_hi
        }
        tc_commit(&S.mma_done);
      }
      // rhs while the MMAs run: b_r += c_j x_j[r]; x = hi + lo exactly, from the N-side tiles (row r of a tile = feature r)
      {
        const int roff = (r >> 3) * kTcSBO + (r & 7) * 16;
#pragma unroll
        for (int kb = 0; kb < kTcRows / 4; kb++) {
          if (jbase + 4 * kb >= n) break;   // CTA-uniform
          const float4 h4 = *reinterpret_cast<const float4*>(&S.op[2][roff + kb * kTcLBO]);
          const float4 l4 = *reinterpret_cast<const float4*>(&S.op[3][roff + kb * kTcLBO]);
          const float4 c4 = *reinterpret_cast<const float4*>(&S.cs[slot][4 * kb]);   // padding entries are zero
          br = fmaf(c4.x, h4.x + l4.x, fmaf(c4.y, h4.y + l4.y, fmaf(c4.z, h4.z + l4.z, fmaf(c4.w, h4.w + l4.w, br))));
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");   // metadata of chunk ch + 2 (consumed two chunks from now)
      mbar_wait(&S.mma_done, mma_phase);
      mma_phase ^= 1;
      if (window_end) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < K; c0 += 32) {
          uint32_t d[32];
          const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
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
      }
      __syncthreads();   // tiles and the oldest metadata slot are free; the accumulator has been read before the next window
    }

    // ---- A = XtX (diag(d) in the eigenbasis) + X_nnz diag(w) X_nnz'  /  X_nnz X_nnz' + lambda_u I ----------------------
    if (full_g) {
#pragma unroll
      for (int c = 0; c < K; c++) acc[c] += __ldg(P.G + (size_t)c * K + r);   // symmetric: row r read as column r (coalesced)
    }
    const float dshift = implicit ? dgr : lam_use;   // on the diagonal (0 with the full XtX: it already carries lambda)
    // mat-vec with this thread's row: vector from shared memory (broadcast reads)
    auto matvec = [&](float vr) -> float {
      float* vb = S.vec[sumbuf];
      vb[r] = vr;
      __syncthreads();
      float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < K; c += 4) {
        const float4 v4 = *reinterpret_cast<const float4*>(vb + c);
        s0 = __ffma2_rn(make_float2(acc[c], acc[c + 1]), make_float2(v4.x, v4.y), s0);
        s1 = __ffma2_rn(make_float2(acc[c + 2], acc[c + 3]), make_float2(v4.z, v4.w), s1);
      }
      return fmaf(dshift, vr, (s0.x + s0.y) + (s1.x + s1.y));
    };

    // ---- CG (cg_solver_implicit / cg_solver_explicit) ------------------------------------------------------------
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
    P.Y[(size_t)row * K + r] = xr;

    // ---- loss (wrmf_implicit.hpp:259-261 / wrmf_explicit.hpp:131-132): u_j = x_j . y, one more pass over the tile ----
    {
      float* vb = S.vec[sumbuf];
      vb[r] = xr;
      __syncthreads();
      const float4 y4 = *reinterpret_cast<const float4*>(vb + 4 * lane);
      float l = 0.0f;
      // 16 gathered rows in flight per warp; their 16 dot products are reduced by a transposing halving reduction
      // (16 shuffles for 16 rows).  Register s of lane L holds row q + (s ^ slot), slot = L >> 1, which makes every halving
      // level the same instruction stream for all lanes; after the four levels + one butterfly lanes 2 slot and 2 slot + 1
      // hold u of row q + slot.
      const int slot = lane >> 1;
      for (int base = warp * 32; base < n; base += 4 * 32) {          // this warp's groups of 32 gathered rows
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
      l = warp_sum(l);
      // every lane of a warp carries the same l; the regulariser term once per row
      float tot = block_sum((lane == 0) ? l : 0.0f);
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
