// als_chol_rows_split.cuh -- rank-128 variant of als_chol_rows_kernel with every row of the lower half SPLIT OVER TWO
// THREADS, for occupancy: at rank 128 the row-per-thread kernel needs 128 accumulator registers per thread and runs
// 12 warps per SM (ncu: 56 % of the samples are barrier / fixed-latency stalls with 2-3 warps per scheduler).  Here a CTA
// has 192 threads and every thread holds 64 accumulators:
//   warp 0: rows 96..127, columns 0..63      warp 4: rows 96..127, columns 64..127
//   warp 1: rows 64..95,  columns 0..63      warp 5: rows 64..95,  columns 64..95
//   warp 2: rows 32..63,  columns 0..63      warp 3: rows 0..31, columns 0..31
// (rows in this order so that warps 0/4 and 1/5 sit on the same TMEM lane quadrant -- 32 (warp % 4) -- which is what a
// tensor-core Gram will need when it is added here).  ~110 registers => 3 CTAs x 6 warps = 18 warps per SM.
// The factorisation is the one of als_chol_rows.cuh (4-column panels, P1 publish / P2 solve / P3 update, left-shifting
// register windows).  While the panel lies left of column 64 the `right` threads do not own the panel's entries of
// their row: after the second barrier they read l (4 scalars) and z (one float4) from shared memory, keep their own copy
// of the rhs entry, and update their window in place; from column 64 on they are the panel owners and the `left`
// warps only keep the barriers.  Index logic: scripts/emulate_chol_rows.py::emulate_split128 (tests/
// test_chol_rows_emulation.py).  EXPERIMENTAL (kernel = 8): written after the round's GPU budget was spent, compiled
// and emulated but not yet run on hardware.
#pragma once
#include "als_chol_rows.cuh"

namespace b200als {

constexpr int kSplitThreads = 192;

__global__ void __launch_bounds__(kSplitThreads, 3) als_chol_rows_split_kernel(SolveParams<float> P) {
  constexpr int K = 128;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using SM = CholRowsSmem<K>;
  SM& S = *reinterpret_cast<SM*>(smem_raw);
  constexpr int LDT = SM::LDT;
  constexpr int NT = kSplitThreads, NW = NT / 32;
  constexpr int NB4 = K / 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rbase = (warp == 0 || warp == 4) ? 96 : (warp == 1 || warp == 5) ? 64 : (warp == 2) ? 32 : 0;
  const bool left = (warp < 4);
  const int r = rbase + lane;                        // the matrix row this thread works on
  const int cb = left ? 0 : 64;                      // first column of its register window
  const int rmax = rbase + 31;                       // last row of the warp
  const int cmax = left ? min(rmax, 63) : rmax;      // last column the warp ever needs
  const bool implicit = (P.feedback == 0);
  const int total = P.n_list_dev ? __ldg(P.n_list_dev) : P.n_list;
  double cta_loss = 0.0;
  auto row_id = [&](int tt) -> int { return P.row_list ? __ldg(P.row_list + tt) : tt + P.row_begin; };
  auto fetch_meta = [&](int buf, int p, int cnt) {
    for (int j = tid; j < cnt; j += NT) {
      cp_async_4(&S.idx[buf][j], P.idx + p + j);
      cp_async_4(&S.cs[buf][j], P.val + p + j);
    }
  };
  int rowA = -1, nA = 0;
  int rowB = -1, pB = 0, nB = 0;
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
    __syncthreads();   // previous row fully consumed; this row's indices / values visible
    fetch_meta(buf ^ 1, pB, nB);
    if (tid == 0) S.fail = 0;
    for (int e = tid; e < n * NB4; e += NT) {
      const int j = e / NB4, c4 = e - j * NB4;
      cp_async_16(&S.tile[j * K + c4 * 4], P.X + (size_t)s_idx[j] * K + c4 * 4);
    }
    // ---- while the tile is in flight: my window of column r of XtX (symmetric), or lambda_u on the diagonal --------------
    const float lam_use = implicit ? 0.0f : (float)(P.lambda * (P.dynamic_lambda ? (double)(float)n : 1.));
    float2 a[32];   // a[i] = columns (cb + 2i, cb + 2i + 1) of row r; shifted left by one block per owned panel
#pragma unroll
    for (int c4 = 0; c4 < 16; c4++) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      const int c = cb + 4 * c4;
      if (implicit) {
        if (c <= cmax) {
          g.x = __ldg(P.G + (size_t)(c + 0) * K + r);
          g.y = __ldg(P.G + (size_t)(c + 1) * K + r);
          g.z = __ldg(P.G + (size_t)(c + 2) * K + r);
          g.w = __ldg(P.G + (size_t)(c + 3) * K + r);
        }
      } else {
        if (c + 0 == r) g.x = lam_use;
        if (c + 1 == r) g.y = lam_use;
        if (c + 2 == r) g.z = lam_use;
        if (c + 3 == r) g.w = lam_use;
      }
      a[2 * c4] = make_float2(g.x, g.y);
      a[2 * c4 + 1] = make_float2(g.z, g.w);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // ---- Gram + rhs (every thread keeps its own copy of the rhs entry of its row) -----------------------------------------
    float br = 0.0f;
#pragma unroll 2
    for (int j = 0; j < n; j++) {
      const float xr = S.tile[j * K + r];
      const float cj = s_cs[j];
      const float wx = implicit ? xr * (cj - 1.0f) : xr;
      br = fmaf(cj, xr, br);
      const float2 w2 = make_float2(wx, wx);
#pragma unroll
      for (int c4 = 0; c4 < 16; c4++) {
        if (cb + 4 * c4 > cmax) break;   // warp-uniform
        const float4 v = *reinterpret_cast<const float4*>(&S.tile[j * K + cb + 4 * c4]);
        a[2 * c4] = __ffma2_rn(w2, make_float2(v.x, v.y), a[2 * c4]);
        a[2 * c4 + 1] = __ffma2_rn(w2, make_float2(v.z, v.w), a[2 * c4 + 1]);
      }
    }
    // ---- right-looking Cholesky, 4 columns per pair of barriers ----------------------------------------------------------
    bool failed = false;
    for (int p = 0; p < NB4; p++) {
      const int j0 = 4 * p;
      const bool phase_a = (j0 < 64);                              // the panel lies in the left half
      const bool owner = (left == phase_a) && (rmax >= j0);       // warp-uniform: my window holds the panel's columns
      // P1: the diagonal block's four owner threads publish their rows (window registers 0, 1) and rhs entries
      if (owner && r >= j0 && r < j0 + 4) {
        *reinterpret_cast<float4*>(&S.D[r - j0][0]) = make_float4(a[0].x, a[0].y, a[1].x, a[1].y);
        S.D[r - j0][4] = br;
      }
      __syncthreads();
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      if (owner) {
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
        if (!(fminf(fminf(p0, p1v), fminf(p2v, p3v)) > 0.0f)) failed = true;
        const float z0 = b0 * i0;
        const float z1 = fmaf(-L10, z0, b1) * i1;
        const float z2 = fmaf(-L21, z1, fmaf(-L20, z0, b2)) * i2;
        const float z3 = fmaf(-L32, z2, fmaf(-L31, z1, fmaf(-L30, z0, b3))) * i3;
        if (r == j0) {
          *reinterpret_cast<float4*>(&S.zz[j0]) = make_float4(z0, z1, z2, z3);
          *reinterpret_cast<float4*>(&S.rs[j0]) = make_float4(i0, i1, i2, i3);
        }
        l0 = a[0].x * i0;
        l1 = fmaf(-l0, L10, a[0].y) * i1;
        l2 = fmaf(-l1, L21, fmaf(-l0, L20, a[1].x)) * i2;
        l3 = fmaf(-l2, L32, fmaf(-l1, L31, fmaf(-l0, L30, a[1].y))) * i3;
        S.Lt[(j0 + 0) * LDT + r] = l0;
        S.Lt[(j0 + 1) * LDT + r] = l1;
        S.Lt[(j0 + 2) * LDT + r] = l2;
        S.Lt[(j0 + 3) * LDT + r] = l3;
        br = fmaf(-l3, z3, fmaf(-l2, z2, fmaf(-l1, z1, fmaf(-l0, z0, br))));
      }
      __syncthreads();
      if (owner) {
        // P3 of the panel's owners: window shifts one block to the left
        const float2 n0 = make_float2(-l0, -l0), n1 = make_float2(-l1, -l1), n2 = make_float2(-l2, -l2),
                     n3 = make_float2(-l3, -l3);
        const float* lt = &S.Lt[j0 * LDT + j0 + 4];
#pragma unroll
        for (int ib = 0; ib < 15; ib++) {
          if (j0 + 4 + 4 * ib > cmax) break;   // warp-uniform
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
      } else if (phase_a && !left) {
        // right halves while the panel is still in the left half: l and z come from shared memory, update in place
        l0 = S.Lt[(j0 + 0) * LDT + r];
        l1 = S.Lt[(j0 + 1) * LDT + r];
        l2 = S.Lt[(j0 + 2) * LDT + r];
        l3 = S.Lt[(j0 + 3) * LDT + r];
        const float4 z4 = *reinterpret_cast<const float4*>(&S.zz[j0]);
        br = fmaf(-l3, z4.w, fmaf(-l2, z4.z, fmaf(-l1, z4.y, fmaf(-l0, z4.x, br))));
        const float2 n0 = make_float2(-l0, -l0), n1 = make_float2(-l1, -l1), n2 = make_float2(-l2, -l2),
                     n3 = make_float2(-l3, -l3);
        const float* lt = &S.Lt[j0 * LDT + 64];
#pragma unroll
        for (int ib = 0; ib < 16; ib++) {
          if (64 + 4 * ib > cmax) break;   // warp-uniform
          const float4 v0 = *reinterpret_cast<const float4*>(lt + 0 * LDT + 4 * ib);
          const float4 v1 = *reinterpret_cast<const float4*>(lt + 1 * LDT + 4 * ib);
          const float4 v2 = *reinterpret_cast<const float4*>(lt + 2 * LDT + 4 * ib);
          const float4 v3 = *reinterpret_cast<const float4*>(lt + 3 * LDT + 4 * ib);
          float2 lo = __ffma2_rn(n0, make_float2(v0.x, v0.y), a[2 * ib]);
          float2 hi = __ffma2_rn(n0, make_float2(v0.z, v0.w), a[2 * ib + 1]);
          lo = __ffma2_rn(n1, make_float2(v1.x, v1.y), lo);
          hi = __ffma2_rn(n1, make_float2(v1.z, v1.w), hi);
          lo = __ffma2_rn(n2, make_float2(v2.x, v2.y), lo);
          hi = __ffma2_rn(n2, make_float2(v2.z, v2.w), hi);
          a[2 * ib] = __ffma2_rn(n3, make_float2(v3.x, v3.y), lo);
          a[2 * ib + 1] = __ffma2_rn(n3, make_float2(v3.z, v3.w), hi);
        }
      }
    }
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
    if (failed) {
      S.fail = 1;
      if (lane == 0) atomicExch(P.status, 1);
    }
    __syncthreads();
    if (S.fail) { advance(); continue; }
    // ---- blocked back substitution (warps 0..3 subtract the solved part, warp 0 solves each 32 x 32 triangle) -----------
    for (int b0 = K - 32; b0 >= 0; b0 -= 32) {
      if (b0 + 32 < K) {
        if (warp < 4) {
          const int nl = (K - (b0 + 32)) / 4, l0s = b0 + 32 + warp * nl;
          const float* lrow = &S.Lt[(b0 + lane) * LDT + l0s];
          float ps = 0.f;
          for (int l = 0; l < nl; l += 4) {
            const float4 lv = *reinterpret_cast<const float4*>(lrow + l);
            const float4 yv = *reinterpret_cast<const float4*>(&S.zz[l0s + l]);
            ps = fmaf(lv.x, yv.x, fmaf(lv.y, yv.y, fmaf(lv.z, yv.z, fmaf(lv.w, yv.w, ps))));
          }
          S.part[warp][lane] = ps;
        }
        __syncthreads();
      }
      if (warp == 0) {
        const int i = b0 + lane;
        const float ri = S.rs[i];
        float zi = S.zz[i];
        if (b0 + 32 < K) {
#pragma unroll
          for (int w2 = 0; w2 < 4; w2++) zi -= S.part[w2][lane];
        }
#pragma unroll 8
        for (int sidx = 31; sidx >= 0; sidx--) {
          const float ys = __shfl_sync(kFull, zi * ri, sidx);
          if (lane < sidx) zi = fmaf(-S.Lt[i * LDT + b0 + sidx], ys, zi);
        }
        S.zz[i] = zi * ri;
      }
      __syncthreads();
    }
    float* y = P.Y + (size_t)row * K;
    if (tid < K / 4) *reinterpret_cast<float4*>(y + tid * 4) = *reinterpret_cast<const float4*>(&S.zz[tid * 4]);
    // ---- loss ---------------------------------------------------------------------------------------------------------
    float l = 0.0f;
    {
      const float4 yv = *reinterpret_cast<const float4*>(&S.zz[lane * 4]);
      for (int jb = warp * 4; jb < n; jb += NW * 4) {
        float4 xv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (jb + u < n) xv[u] = ldg_f4(P.X + (size_t)s_idx[jb + u] * K + lane * 4);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          float dsum = fmaf(xv[u].x, yv.x, fmaf(xv[u].y, yv.y, fmaf(xv[u].z, yv.z, xv[u].w * yv.w)));
          dsum = warp_sum(dsum);
          if (jb + u < n && lane == 0) {
            const float c = s_cs[jb + u];
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
}

}  // namespace b200als
