// als_chol_rows.cuh -- Cholesky branch of the half-iteration, second generation (rank 64 / 128, fp32, rows with
// 1..80 non-zeros): ONE THREAD PER MATRIX ROW, the whole system in registers, panels of 4 columns.
// Reference: lhs = XtX + X_nnz diag(c-1) X_nnz', rhs = X_nnz c, solve(lhs, rhs)   (wrmf_implicit.hpp:207-236)
//            lhs = X_nnz X_nnz' + lambda_u I, rhs = X_nnz r, solve(lhs, rhs)      (wrmf_explicit.hpp:103-108)
//
// (The first-generation kernel kept 16 x 16 register blocks and paid one CTA barrier per COLUMN -- 35.7 k
// warp-instructions per row at rank 64, barrier-bound, 2x slower; it was removed in round 2, profiles/r1 keeps its
// numbers.)  Here a CTA of K threads owns one system, thread r holds row r of the lower triangle (K registers)
// and the right-looking factorisation advances FOUR columns per pair of barriers:
//   P1  the four threads of the diagonal block publish their 4 x 4 block (+ their rhs entries)      -> barrier
//   P2  every thread factors that 4 x 4 block redundantly (4 MUFU.RSQ + ~30 flop, no serial owner), solves its own
//       four entries of the panel against it (l = a L_D^-T), publishes them transposed (Lt[j][c] = L[c][j], conflict-
//       free scalar stores), advances the forward substitution (z = L^-1 rhs)                        -> barrier
//   P3  rank-4 update of the thread's own row from the published panel: 4 broadcast LDS.128 + 8 FFMA2 per 4 columns,
//       only the columns right of the panel and left of the warp's last row (uniform branches, no masks).
// The register file is indexed statically without unrolling the panel loop: P3 writes the updated column block one
// block to the LEFT of where it read it (d = a*b + c with d != c is free), so the current panel is always registers
// 0..3 and the trip count of P3 shrinks by one block per panel -- triangular work, constant code.
// The shape (thread r owns row r) is exactly what `tcgen05.ld.32x32b` delivers, so the FFMA Gram below can be
// swapped for a TMEM accumulator without touching the factorisation.
// The gathered tile is staged once (16-byte cp.async) and its shared memory is re-used for Lt; the loss re-reads the
// gathered rows through L2 (they were fetched microseconds earlier).  Back substitution: blocked by 32 rows, all
// warps subtract the solved part, warp 0 solves each 32 x 32 triangle with one shuffle per step.
// Algorithmic work per row: 2nK^2 / 2 .. 2nK^2 (Gram, triangular per warp) + K^3/3 flop; bytes as the CG path.
#pragma once
#include "als_generic.cuh"  // SolveParams
#include "rotate_tc.cuh"   // gram_tc.cuh + rt_desc / tf32_hi

namespace b200als {

constexpr int kCholMaxN = 80;   // longest row (gathered rows) the register / shared-memory layout of these kernels holds

template <int K>
struct alignas(128) CholRowsSmem {
  static constexpr int LDT = K + 4;      // rows stay 16-byte aligned, consecutive rows 4 banks apart
  union {
    float tile[kCholMaxN * K];           // gathered rows (FFMA Gram); 16-byte cp.async destinations
    float Lt[K * LDT];                   // Lt[j][c] = L[c][j] for c >= j (factorisation, back substitution)
    unsigned char op[4][(K == kTcK) ? kTcTileBytes : 16];   // tensor-core Gram (rank 128): tiles A hi/lo (= w x), B hi/lo (= x),
                                         // 32 gathered rows each, K-major core-matrix layout of gram_tc.cuh
  };
  alignas(16) float D[4][8];             // diagonal 4 x 4 block, row i = [d_i0 .. d_i3, rhs_i, -, -, -]
  float cs[2][kCholMaxN];                // confidences / ratings of the row, double buffered (next row lands by cp.async)
  int idx[2][kCholMaxN];
  float rs[K];                           // 1 / L[j][j]
  alignas(16) float zz[K];               // z = L^-1 rhs, then y
  float part[K / 32][32];                // back substitution: per-warp partial sums of the solved part
  alignas(8) double red[32];
  alignas(8) uint64_t mma_done[2];       // tensor-core Gram: the chunk's MMAs have completed (tcgen05.commit), per tile buffer
  uint32_t tmem_base;
  int fail;
};

// kCtas: resident CTAs per SM the register allocation is sized for (rank 64: 8; rank 128: 3 = 168 registers with 56 B of
// cold spills, the default -- 12 instead of 8 warps per SM is worth 27 % -- or 2 = 226 registers)
// kTc (rank 128 only): the per-row Gram X_nnz diag(w) X_nnz' on tcgen05 (3xTF32, one 128 x 128 TMEM accumulator) instead of
// the FFMA loop: the gathered rows go global -> registers -> transposed hi/lo operand tiles (the staging of gram_tc.cuh
// with a weighted copy for the M side), one thread issues 3 MMAs per 8 gathered rows, and `tcgen05.ld.32x32b` hands
// thread r row r of the result -- the layout the factorisation below works in.
template <int K, int kCtas, int kTc = 0>
__global__ void __launch_bounds__(K, kCtas) als_chol_rows_kernel(SolveParams<float> P) {
  static_assert(!kTc || K == kTcK, "tensor-core Gram: rank 128 only");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using SM = CholRowsSmem<K>;
  SM& S = *reinterpret_cast<SM*>(smem_raw);
  constexpr int LDT = SM::LDT;
  constexpr int NW = K / 32;
  constexpr int NB4 = K / 4;             // 4-column blocks per row
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = tid;                     // the matrix row this thread owns
  const int wmax = warp * 32 + 31;       // last row owned by this warp: columns beyond it are never needed here
  const bool implicit = (P.feedback == 0);
  const int total = P.n_list_dev ? __ldg(P.n_list_dev) : P.n_list;
  double cta_loss = 0.0;
  uint32_t tmem = 0, mma_phase = 0;
  if constexpr (kTc) {
    if (tid == 0) {
      mbar_init(&S.mma_done[0], 1);
      mbar_init(&S.mma_done[1], 1);
      mbar_fence_init();
    }
    if (warp == 0) {   // one 128-lane x 128-column fp32 accumulator; 3 CTAs/SM x 128 <= 512 TMEM columns
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem = S.tmem_base;
  }
  // Software pipeline of the CSR metadata: the row pointers of row i+2 are loaded (pinned in program order) during
  // row i, the indices / values of row i+1 travel global -> shared by 4-byte cp.async during row i, next to its tile.
  auto row_id = [&](int tt) -> int { return P.row_list ? __ldg(P.row_list + tt) : tt + P.row_begin; };
  auto fetch_meta = [&](int buf, int p, int cnt) {
    for (int j = tid; j < cnt; j += K) {
      cp_async_4(&S.idx[buf][j], P.idx + p + j);
      cp_async_4(&S.cs[buf][j], P.val + p + j);
    }
  };
  int rowA = -1, nA = 0;                  // current row
  int rowB = -1, pB = 0, nB = 0;          // next row
  {
    const int t0 = blockIdx.x, t1 = blockIdx.x + gridDim.x;
    if (t0 < total) {
      rowA = row_id(t0);
      const int pA = __ldg(P.ptr + rowA) - P.ptr_base;
      nA = __ldg(P.ptr + rowA + 1) - P.ptr_base - pA;
      fetch_meta(0, pA, nA);
    }
    if (t1 < total) {
      rowB = row_id(t1);
      pB = __ldg(P.ptr + rowB) - P.ptr_base;
      nB = __ldg(P.ptr + rowB + 1) - P.ptr_base - pB;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  int buf = 0;
  for (int t = blockIdx.x; t < total; t += gridDim.x, buf ^= 1) {
    const int row = rowA, n = nA;
    const int* s_idx = S.idx[buf];
    const float* s_cs = S.cs[buf];
    __syncthreads();   // previous row fully consumed; this row's indices / values (landed last iteration) visible
    fetch_meta(buf ^ 1, pB, nB);
    if (tid == 0) S.fail = 0;
    if constexpr (!kTc) {
      for (int e = tid; e < n * NB4; e += K) {
        const int j = e / NB4, c4 = e - j * NB4;
        cp_async_16(&S.tile[j * K + c4 * 4], P.X + (size_t)s_idx[j] * K + c4 * 4);
      }
    }
    // ---- while the tile is in flight: row r of XtX (implicit) or lambda_u on the diagonal (explicit) -------------
    const float lam_use = implicit ? 0.0f : (float)(P.lambda * (P.dynamic_lambda ? (double)(float)n : 1.));
    float2 a[K / 2];   // a[i] = columns (2i, 2i+1) of row r; after panel p: columns (4p + 2i, 4p + 2i + 1)
    auto init_a = [&]() {
  #pragma unroll
      for (int c4 = 0; c4 < NB4; c4++) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (implicit) {
          // XtX is symmetric: row r is read as column r, so that the 32 lanes of a warp read 32 consecutive floats
          if (4 * c4 <= wmax) {
            g.x = __ldg(P.G + (size_t)(4 * c4 + 0) * K + r);
            g.y = __ldg(P.G + (size_t)(4 * c4 + 1) * K + r);
            g.z = __ldg(P.G + (size_t)(4 * c4 + 2) * K + r);
            g.w = __ldg(P.G + (size_t)(4 * c4 + 3) * K + r);
          }
        } else {
          if (4 * c4 + 0 == r) g.x = lam_use;
          if (4 * c4 + 1 == r) g.y = lam_use;
          if (4 * c4 + 2 == r) g.z = lam_use;
          if (4 * c4 + 3 == r) g.w = lam_use;
        }
        a[2 * c4] = make_float2(g.x, g.y);
        a[2 * c4 + 1] = make_float2(g.z, g.w);
      }
    };
    if constexpr (!kTc) init_a();   // FFMA Gram: overlaps the tile's flight; tensor-core Gram: after the MMAs (registers)
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    float br = 0.0f;
    if constexpr (kTc == 1) {
      // ---- Gram on the tensor core, 32 gathered rows per chunk ----------------------------------------------------
      const int n_chunks = (n + kTcRows - 1) / kTcRows;
      constexpr int kMaxChunks = (kCholMaxN + kTcRows - 1) / kTcRows;   // 3
      // all gathered rows of this warp's 4-row groups are requested up front: ONE memory round trip per solved row
      float4 vv[kMaxChunks][kTcRows / 16][4];
#pragma unroll
      for (int ch = 0; ch < kMaxChunks; ch++)
#pragma unroll
        for (int g = 0; g < kTcRows / 16; g++)
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            const int j = ch * kTcRows + (warp + 4 * g) * 4 + kk;
            vv[ch][g][kk] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < n) vv[ch][g][kk] = ldg_f4(P.X + (size_t)s_idx[j] * K + lane * 4);
          }
#pragma unroll
      for (int ch = 0; ch < kMaxChunks; ch++) {
        if (ch >= n_chunks) break;   // CTA-uniform
        const int jbase = ch * kTcRows;
        // stage: this warp transposes the 4-row groups {warp, warp + 4} of the chunk (cf. gram_tc_blocks_kernel)
#pragma unroll
        for (int g = 0; g < kTcRows / 16; g++) {
          const int kb = warp + 4 * g;
          const float4 (&v)[4] = vv[ch][g];
          float wv[4];
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            const int j = jbase + kb * 4 + kk;
            wv[kk] = (j < n) ? (implicit ? (s_cs[j] - 1.0f) : 1.0f) : 0.f;
          }
          const float col[4][4] = {{v[0].x, v[1].x, v[2].x, v[3].x}, {v[0].y, v[1].y, v[2].y, v[3].y},
                                   {v[0].z, v[1].z, v[2].z, v[3].z}, {v[0].w, v[1].w, v[2].w, v[3].w}};
#pragma unroll
          for (int jj = 0; jj < 4; jj++) {
            const int m = 4 * lane + jj;   // feature
            float4 bh, bl, ah, al;
            const float x0 = col[jj][0], x1 = col[jj][1], x2 = col[jj][2], x3 = col[jj][3];
            const float y0 = x0 * wv[0], y1 = x1 * wv[1], y2 = x2 * wv[2], y3 = x3 * wv[3];
            bh.x = __uint_as_float(__float_as_uint(x0) & 0xFFFFE000u); bl.x = x0 - bh.x;
            bh.y = __uint_as_float(__float_as_uint(x1) & 0xFFFFE000u); bl.y = x1 - bh.y;
            bh.z = __uint_as_float(__float_as_uint(x2) & 0xFFFFE000u); bl.z = x2 - bh.z;
            bh.w = __uint_as_float(__float_as_uint(x3) & 0xFFFFE000u); bl.w = x3 - bh.w;
            ah.x = __uint_as_float(__float_as_uint(y0) & 0xFFFFE000u); al.x = y0 - ah.x;
            ah.y = __uint_as_float(__float_as_uint(y1) & 0xFFFFE000u); al.y = y1 - ah.y;
            ah.z = __uint_as_float(__float_as_uint(y2) & 0xFFFFE000u); al.z = y2 - ah.z;
            ah.w = __uint_as_float(__float_as_uint(y3) & 0xFFFFE000u); al.w = y3 - ah.w;
            const int off = (m >> 3) * kTcSBO + kb * kTcLBO + (m & 7) * 16;
            *reinterpret_cast<float4*>(&S.op[0][off]) = ah;
            *reinterpret_cast<float4*>(&S.op[1][off]) = al;
            *reinterpret_cast<float4*>(&S.op[2][off]) = bh;
            *reinterpret_cast<float4*>(&S.op[3][off]) = bl;
          }
        }
        fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core's async proxy
        __syncthreads();
        if (tid == 0) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int ksteps = min(kTcRows / 8, (n - jbase + 7) / 8);   // 8 gathered rows per MMA; the padding rows are zero
          for (int ks = 0; ks < ksteps; ks++) {
            const uint64_t dah = tc_smem_desc(&S.op[0][ks * 2 * kTcLBO]);
            const uint64_t dal = tc_smem_desc(&S.op[1][ks * 2 * kTcLBO]);
            const uint64_t dbh = tc_smem_desc(&S.op[2][ks * 2 * kTcLBO]);
            const uint64_t dbl = tc_smem_desc(&S.op[3][ks * 2 * kTcLBO]);
            tc_mma_tf32(tmem, dah, dbh, (ch == 0 && ks == 0) ? 0u : 1u);   // (w x)_hi' x_hi
            tc_mma_tf32(tmem, dah, dbl, 1u);                                // (w x)_hi' x_lo
            tc_mma_tf32(tmem, dal, dbh, 1u);                                // (w x)_lo' x_hi
          }
          tc_commit(&S.mma_done[0]);
        }
        // rhs while the MMAs run: b_r += c_j x_j[r], x = hi + lo exactly, from the B tiles (row r of the tile = feature r)
        {
          const int roff = (r >> 3) * kTcSBO + (r & 7) * 16;
#pragma unroll
          for (int kb = 0; kb < kTcRows / 4; kb++) {
            if (jbase + 4 * kb >= n) break;   // CTA-uniform
            const float4 h4 = *reinterpret_cast<const float4*>(&S.op[2][roff + kb * kTcLBO]);
            const float4 l4 = *reinterpret_cast<const float4*>(&S.op[3][roff + kb * kTcLBO]);
            const int j = jbase + 4 * kb;
            const float c0 = s_cs[j], c1 = (j + 1 < n) ? s_cs[j + 1] : 0.f, c2 = (j + 2 < n) ? s_cs[j + 2] : 0.f,
                        c3 = (j + 3 < n) ? s_cs[j + 3] : 0.f;
            br = fmaf(c0, h4.x + l4.x, fmaf(c1, h4.y + l4.y, fmaf(c2, h4.z + l4.z, fmaf(c3, h4.w + l4.w, br))));
          }
        }
        mbar_wait(&S.mma_done[0], mma_phase);
        mma_phase ^= 1;
        __syncthreads();   // every thread has read its rhs entries before the tiles are overwritten (next chunk / Lt)
      }
      // accumulator -> registers: thread r receives row r (lanes 32 warp .. 32 warp + 31 of TMEM)
      init_a();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < K; c0 += 32) {
        if (c0 > wmax) break;   // warp-uniform: columns beyond the warp's last row are never needed
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
        for (int c = 0; c < 32; c += 2) {
          a[(c0 + c) >> 1].x += __uint_as_float(d[c]);
          a[(c0 + c) >> 1].y += __uint_as_float(d[c + 1]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // ordered before the next row's MMAs by its barriers
    } else {
      // ---- Gram + rhs: a[c] += (w_j x_j[r]) x_j[c],  b_r += c_j x_j[r] ----------------------------------------------
  #pragma unroll 2
      for (int j = 0; j < n; j++) {
        const float xr = S.tile[j * K + r];
        const float cj = s_cs[j];
        const float wx = implicit ? xr * (cj - 1.0f) : xr;
        br = fmaf(cj, xr, br);
        const float2 w2 = make_float2(wx, wx);
  #pragma unroll
        for (int c4 = 0; c4 < NB4; c4++) {
          if (4 * c4 > wmax) break;   // warp-uniform
          const float4 v = *reinterpret_cast<const float4*>(&S.tile[j * K + 4 * c4]);
          a[2 * c4] = __ffma2_rn(w2, make_float2(v.x, v.y), a[2 * c4]);
          a[2 * c4 + 1] = __ffma2_rn(w2, make_float2(v.z, v.w), a[2 * c4 + 1]);
        }
      }
    }
    // ---- right-looking Cholesky, 4 columns per pair of barriers ----------------------------------------------------
    bool failed = false;
    for (int p = 0; p < NB4; p++) {
      const int j0 = 4 * p;
      // P1: the diagonal block's four threads publish their rows (columns j0..j0+3 = registers 0, 1) and rhs entries
      if (r >= j0 && r < j0 + 4) {
        *reinterpret_cast<float4*>(&S.D[r - j0][0]) = make_float4(a[0].x, a[0].y, a[1].x, a[1].y);
        S.D[r - j0][4] = br;
      }
      __syncthreads();
      // P2 (warps that still own rows at or below the panel): factor the 4 x 4 block redundantly per thread
      const bool active = (wmax >= j0);   // warp-uniform: warps entirely above the panel only keep the barriers
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      if (active) {
        const float4 d0 = *reinterpret_cast<const float4*>(&S.D[0][0]);
        const float4 d1 = *reinterpret_cast<const float4*>(&S.D[1][0]);
        const float4 d2 = *reinterpret_cast<const float4*>(&S.D[2][0]);
        const float4 d3 = *reinterpret_cast<const float4*>(&S.D[3][0]);
        const float b0 = S.D[0][4], b1 = S.D[1][4], b2 = S.D[2][4], b3 = S.D[3][4];
        const float p0 = d0.x;
        const float i0 = rsqrtf(p0);
        const float L10 = d1.x * i0, L20 = d2.x * i0, L30 = d3.x * i0;
        const float p1v = fmaf(-L10, L10, d1.y);
        const float i1 = rsqrtf(p1v);
        const float L21 = fmaf(-L20, L10, d2.y) * i1, L31 = fmaf(-L30, L10, d3.y) * i1;
        const float p2v = fmaf(-L21, L21, fmaf(-L20, L20, d2.z));
        const float i2 = rsqrtf(p2v);
        const float L32 = fmaf(-L31, L21, fmaf(-L30, L20, d3.z)) * i2;
        const float p3v = fmaf(-L32, L32, fmaf(-L31, L31, fmaf(-L30, L30, d3.w)));
        const float i3 = rsqrtf(p3v);
        // a non-positive (or NaN) pivot is remembered, not branched on: every warp keeps the same barrier count
        if (!(fminf(fminf(p0, p1v), fminf(p2v, p3v)) > 0.0f)) failed = true;
        // forward substitution through the block: z = L^-1 rhs
        const float z0 = b0 * i0;
        const float z1 = fmaf(-L10, z0, b1) * i1;
        const float z2 = fmaf(-L21, z1, fmaf(-L20, z0, b2)) * i2;
        const float z3 = fmaf(-L32, z2, fmaf(-L31, z1, fmaf(-L30, z0, b3))) * i3;
        if (r == j0) {
          *reinterpret_cast<float4*>(&S.zz[j0]) = make_float4(z0, z1, z2, z3);
          *reinterpret_cast<float4*>(&S.rs[j0]) = make_float4(i0, i1, i2, i3);
        }
        // my four entries of the panel: l = a L_D^-T  (for the block's own threads this reproduces the block factor)
        l0 = a[0].x * i0;
        l1 = fmaf(-l0, L10, a[0].y) * i1;
        l2 = fmaf(-l1, L21, fmaf(-l0, L20, a[1].x)) * i2;
        l3 = fmaf(-l2, L32, fmaf(-l1, L31, fmaf(-l0, L30, a[1].y))) * i3;
        S.Lt[(j0 + 0) * LDT + r] = l0;   // entries with r < j0 + q are never read (c < j)
        S.Lt[(j0 + 1) * LDT + r] = l1;
        S.Lt[(j0 + 2) * LDT + r] = l2;
        S.Lt[(j0 + 3) * LDT + r] = l3;
        br = fmaf(-l3, z3, fmaf(-l2, z2, fmaf(-l1, z1, fmaf(-l0, z0, br))));   // meaningful for r > j0 + 3
      }
      __syncthreads();
      if (active) {
        // P3: a[c] -= sum_q l_q L[c][j0+q] for the column blocks right of the panel, written one block to the left
        const float2 n0 = make_float2(-l0, -l0), n1 = make_float2(-l1, -l1), n2 = make_float2(-l2, -l2),
                     n3 = make_float2(-l3, -l3);
        const float* lt = &S.Lt[j0 * LDT + j0 + 4];
#pragma unroll
        for (int ib = 0; ib < NB4 - 1; ib++) {
          if (j0 + 4 + 4 * ib > wmax) break;   // warp-uniform
          const float4 v0 = *reinterpret_cast<const float4*>(lt + 0 * LDT + 4 * ib);
          const float4 v1 = *reinterpret_cast<const float4*>(lt + 1 * LDT + 4 * ib);
          const float4 v2 = *reinterpret_cast<const float4*>(lt + 2 * LDT + 4 * ib);
          const float4 v3 = *reinterpret_cast<const float4*>(lt + 3 * LDT + 4 * ib);
          float2 lo = __ffma2_rn(n0, make_float2(v0.x, v0.y), a[2 * ib + 2]);
          float2 hi = __ffma2_rn(n0, make_float2(v0.z, v0.w), a[2 * ib + 3]);
          lo = __ffma2_rn(n1, make_float2(v1.x, v1.y), lo);
          hi = __ffma2_rn(n1, make_float2(v1.z, v1.w), hi);
          lo = __ffma2_rn(n2, make_float2(v2.x, v2.y), lo);
          hi = __ffma2_rn(n2, make_float2(v2.z, v2.w), hi);
          a[2 * ib] = __ffma2_rn(n3, make_float2(v3.x, v3.y), lo);
          a[2 * ib + 1] = __ffma2_rn(n3, make_float2(v3.z, v3.w), hi);
        }
      }
    }
    // row pointers of the row after next: issued here (the system's registers are dead), consumed by advance()
    // after the loss loop's own L2 round trip
    int rowC = -1, pC0 = 0, pC1 = 0;
    if (t + 2 * (int)gridDim.x < total) {
      rowC = P.row_list ? ld_pinned_i32(P.row_list + t + 2 * gridDim.x) : t + 2 * (int)gridDim.x + P.row_begin;
      pC0 = ld_pinned_i32(P.ptr + rowC);
      pC1 = ld_pinned_i32(P.ptr + rowC + 1);
    }
    auto advance = [&]() {
      rowA = rowB; nA = nB;
      rowB = rowC; pB = pC0 - P.ptr_base; nB = pC1 - pC0;
    };
    if (failed) {   // seen by every thread of the warps that were active at that panel
      S.fail = 1;
      if (lane == 0) atomicExch(P.status, 1);
    }
    __syncthreads();
    if (S.fail) { advance(); continue; }   // Y row untouched; status reports B200ALS_ENOTSPD
    // ---- blocked back substitution  L' y = z :  y_r = (z_r - sum_{c > r} Lt[r][c] y_c) * rs_r ------------------------
    for (int b0 = K - 32; b0 >= 0; b0 -= 32) {
      if (b0 + 32 < K) {
        // the solved part l >= b0 + 32 is split over the warps; lane <-> row b0 + lane, contiguous 16-byte reads of Lt
        const int nl = (K - (b0 + 32)) / NW, l0s = b0 + 32 + warp * nl;
        const float* lrow = &S.Lt[(b0 + lane) * LDT + l0s];
        float ps = 0.f;
        for (int l = 0; l < nl; l += 4) {
          const float4 lv = *reinterpret_cast<const float4*>(lrow + l);
          const float4 yv = *reinterpret_cast<const float4*>(&S.zz[l0s + l]);
          ps = fmaf(lv.x, yv.x, fmaf(lv.y, yv.y, fmaf(lv.z, yv.z, fmaf(lv.w, yv.w, ps))));
        }
        S.part[warp][lane] = ps;
        __syncthreads();
      }
      if (warp == 0) {
        const int i = b0 + lane;
        const float ri = S.rs[i];
        float zi = S.zz[i];
        if (b0 + 32 < K) {
#pragma unroll
          for (int w2 = 0; w2 < NW; w2++) zi -= S.part[w2][lane];
        }
#pragma unroll 8
        for (int sidx = 31; sidx >= 0; sidx--) {
          const float ys = __shfl_sync(kFull, zi * ri, sidx);   // y_{b0+s}, final once step s is reached
          if (lane < sidx) zi = fmaf(-S.Lt[i * LDT + b0 + sidx], ys, zi);
        }
        S.zz[i] = zi * ri;
      }
      __syncthreads();
    }
    float* y = P.Y + (size_t)row * K;
    if (tid < K / 4) *reinterpret_cast<float4*>(y + tid * 4) = *reinterpret_cast<const float4*>(&S.zz[tid * 4]);
    // ---- loss (wrmf_implicit.hpp:259-261 / wrmf_explicit.hpp:131-132); the gathered rows come back through L2 -------
    float l = 0.0f;
    {
      const float4 yv = (lane * 4 < K) ? *reinterpret_cast<const float4*>(&S.zz[lane * 4]) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j0 = warp * 4; j0 < n; j0 += NW * 4) {   // four gathered rows per warp in flight
        float4 xv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j0 + u < n && lane * 4 < K) xv[u] = ldg_f4(P.X + (size_t)s_idx[j0 + u] * K + lane * 4);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          float dsum = fmaf(xv[u].x, yv.x, fmaf(xv[u].y, yv.y, fmaf(xv[u].z, yv.z, xv[u].w * yv.w)));
          dsum = warp_sum(dsum);
          if (j0 + u < n && lane == 0) {
            const float c = s_cs[j0 + u];
            const float tt = implicit ? (1.0f - dsum) : (c - dsum);
            l += implicit ? tt * tt * c : tt * tt;
          }
        }
      }
    }
    if (warp == 0) {
      float yy = 0.0f;
      for (int f = lane; f < K; f += 32) yy = fmaf(S.zz[f], S.zz[f], yy);
      yy = warp_sum(yy);
      if (lane == 0) l = fmaf(implicit ? (float)P.lambda : lam_use, yy, l);
    }
    cta_loss += block_sum_double((double)l, S.red);
    advance();
  }
  if (tid == 0) P.loss_partials[blockIdx.x] = cta_loss;
  if constexpr (kTc) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
  }
}

}  // namespace b200als
