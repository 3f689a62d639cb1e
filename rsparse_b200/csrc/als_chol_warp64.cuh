// als_chol_warp64.cuh -- rank-64 variant of als_chol_rows_kernel with ONE WARP PER SYSTEM: lane l works on rows l
// (columns 0..31, 32 accumulators) and l + 32 (columns 0..63, 64 accumulators).  Same algorithm, same panels, same
// shared-memory layout as als_chol_rows.cuh; what changes is who shares what:
//   * every broadcast LDS.128 of the Gram and of the rank-4 update feeds BOTH rows of the lane (ncu on the two-warp
//     kernel: the l1tex data pipe is 82 % busy and a broadcast LDS.128 costs two wavefronts -- here a third of the
//     Gram's and a seventh of the update's operand loads disappear);
//   * the 4 x 4 diagonal block is factored once per panel instead of once per warp;
//   * the CTA is a single warp, so the two barriers per panel are __syncwarp(), not __syncthreads() (21 % of the
//     two-warp kernel's stall samples were barrier stalls).
// Cost: 96 accumulators + temporaries = ~165 registers and 22 KB of shared memory per warp => 10 warps per SM instead of
// 16, each with twice the instruction-level parallelism.
// Default at rank 64 since round 2: parity-green on first run (profiles/r2/), C2 22.8 ms per launch vs 25.2 ms for the
// two-warp kernel (which stays selectable as kernel = 4).  Its per-row index logic is the one emulated by
// scripts/emulate_chol_rows.py::emulate(K = 64).
#pragma once
#include "als_chol_rows.cuh"

namespace b200als {

__global__ void __launch_bounds__(32, 12) als_chol_warp64_kernel(SolveParams<float> P) {
  constexpr int K = 64;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using SM = CholRowsSmem<K>;
  SM& S = *reinterpret_cast<SM*>(smem_raw);
  constexpr int LDT = SM::LDT;
  constexpr int NB4 = K / 4;
  const int lane = threadIdx.x;
  const int r0 = lane, r1 = lane + 32;
  const bool implicit = (P.feedback == 0);
  const int total = P.n_list_dev ? __ldg(P.n_list_dev) : P.n_list;
  double cta_loss = 0.0;
  auto row_id = [&](int tt) -> int { return P.row_list ? __ldg(P.row_list + tt) : tt + P.row_begin; };
  auto fetch_meta = [&](int buf, int p, int cnt) {
    for (int j = lane; j < cnt; j += 32) {
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
    __syncwarp();   // previous row fully consumed; this row's indices / values (landed last iteration) visible
    fetch_meta(buf ^ 1, pB, nB);
    for (int e = lane; e < n * NB4; e += 32) {
      const int j = e / NB4, c4 = e - j * NB4;
      cp_async_16(&S.tile[j * K + c4 * 4], P.X + (size_t)s_idx[j] * K + c4 * 4);
    }
    // ---- while the tile is in flight: columns r0 / r1 of XtX (symmetric), or lambda_u on the diagonal ------------------
    const float lam_use = implicit ? 0.0f : (float)(P.lambda * (P.dynamic_lambda ? (double)(float)n : 1.));
    float2 a0[16];   // row r0, columns (2i, 2i+1); shifted left by one 4-column block per panel
    float2 a1[32];   // row r1
#pragma unroll
    for (int c4 = 0; c4 < NB4; c4++) {
      float4 g1 = make_float4(0.f, 0.f, 0.f, 0.f), g0 = g1;
      if (implicit) {
        g1.x = __ldg(P.G + (size_t)(4 * c4 + 0) * K + r1);
        g1.y = __ldg(P.G + (size_t)(4 * c4 + 1) * K + r1);
        g1.z = __ldg(P.G + (size_t)(4 * c4 + 2) * K + r1);
        g1.w = __ldg(P.G + (size_t)(4 * c4 + 3) * K + r1);
        if (c4 < 8) {
          g0.x = __ldg(P.G + (size_t)(4 * c4 + 0) * K + r0);
          g0.y = __ldg(P.G + (size_t)(4 * c4 + 1) * K + r0);
          g0.z = __ldg(P.G + (size_t)(4 * c4 + 2) * K + r0);
          g0.w = __ldg(P.G + (size_t)(4 * c4 + 3) * K + r0);
        }
      } else {
        if (4 * c4 + 0 == r1) g1.x = lam_use;
        if (4 * c4 + 1 == r1) g1.y = lam_use;
        if (4 * c4 + 2 == r1) g1.z = lam_use;
        if (4 * c4 + 3 == r1) g1.w = lam_use;
        if (4 * c4 + 0 == r0) g0.x = lam_use;
        if (4 * c4 + 1 == r0) g0.y = lam_use;
        if (4 * c4 + 2 == r0) g0.z = lam_use;
        if (4 * c4 + 3 == r0) g0.w = lam_use;
      }
      a1[2 * c4] = make_float2(g1.x, g1.y);
      a1[2 * c4 + 1] = make_float2(g1.z, g1.w);
      if (c4 < 8) {
        a0[2 * c4] = make_float2(g0.x, g0.y);
        a0[2 * c4 + 1] = make_float2(g0.z, g0.w);
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    // ---- Gram + rhs: every broadcast load of x_j feeds both rows -----------------------------------------------------------
    float br0 = 0.0f, br1 = 0.0f;
#pragma unroll 2
    for (int j = 0; j < n; j++) {
      const float x0 = S.tile[j * K + r0], x1 = S.tile[j * K + r1];
      const float cj = s_cs[j];
      const float wj = implicit ? (cj - 1.0f) : 1.0f;
      br0 = fmaf(cj, x0, br0);
      br1 = fmaf(cj, x1, br1);
      const float2 w0 = make_float2(x0 * wj, x0 * wj), w1 = make_float2(x1 * wj, x1 * wj);
#pragma unroll
      for (int c4 = 0; c4 < NB4; c4++) {
        const float4 v = *reinterpret_cast<const float4*>(&S.tile[j * K + 4 * c4]);
        a1[2 * c4] = __ffma2_rn(w1, make_float2(v.x, v.y), a1[2 * c4]);
        a1[2 * c4 + 1] = __ffma2_rn(w1, make_float2(v.z, v.w), a1[2 * c4 + 1]);
        if (c4 < 8) {
          a0[2 * c4] = __ffma2_rn(w0, make_float2(v.x, v.y), a0[2 * c4]);
          a0[2 * c4 + 1] = __ffma2_rn(w0, make_float2(v.z, v.w), a0[2 * c4 + 1]);
        }
      }
    }
    __syncwarp();   // the tile is dead from here on: Lt re-uses its shared memory
    // ---- right-looking Cholesky, 4 columns per pair of warp barriers -----------------------------------------------------
    bool failed = false;
    for (int p = 0; p < NB4; p++) {
      const int j0 = 4 * p;
      const bool phase_a = (j0 < 32);   // the diagonal block belongs to the lanes' first rows
      // P1: the diagonal block's four lanes publish their rows (window registers 0, 1) and rhs entries
      if (phase_a) {
        if (lane >= j0 && lane < j0 + 4) {
          *reinterpret_cast<float4*>(&S.D[lane - j0][0]) = make_float4(a0[0].x, a0[0].y, a0[1].x, a0[1].y);
          S.D[lane - j0][4] = br0;
        }
      } else {
        if (r1 >= j0 && r1 < j0 + 4) {
          *reinterpret_cast<float4*>(&S.D[r1 - j0][0]) = make_float4(a1[0].x, a1[0].y, a1[1].x, a1[1].y);
          S.D[r1 - j0][4] = br1;
        }
      }
      __syncwarp();
      // P2: factor the 4 x 4 block once per lane, solve both rows' panel entries against it
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
      if (lane == 0) {
        *reinterpret_cast<float4*>(&S.zz[j0]) = make_float4(z0, z1, z2, z3);
        *reinterpret_cast<float4*>(&S.rs[j0]) = make_float4(i0, i1, i2, i3);
      }
      // row r1 always holds the panel's columns in its window registers 0, 1
      const float m0 = a1[0].x * i0;
      const float m1 = fmaf(-m0, L10, a1[0].y) * i1;
      const float m2 = fmaf(-m1, L21, fmaf(-m0, L20, a1[1].x)) * i2;
      const float m3 = fmaf(-m2, L32, fmaf(-m1, L31, fmaf(-m0, L30, a1[1].y))) * i3;
      S.Lt[(j0 + 0) * LDT + r1] = m0;
      S.Lt[(j0 + 1) * LDT + r1] = m1;
      S.Lt[(j0 + 2) * LDT + r1] = m2;
      S.Lt[(j0 + 3) * LDT + r1] = m3;
      br1 = fmaf(-m3, z3, fmaf(-m2, z2, fmaf(-m1, z1, fmaf(-m0, z0, br1))));
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      if (phase_a) {   // row r0 takes part while the panel is left of column 32
        l0 = a0[0].x * i0;
        l1 = fmaf(-l0, L10, a0[0].y) * i1;
        l2 = fmaf(-l1, L21, fmaf(-l0, L20, a0[1].x)) * i2;
        l3 = fmaf(-l2, L32, fmaf(-l1, L31, fmaf(-l0, L30, a0[1].y))) * i3;
        S.Lt[(j0 + 0) * LDT + r0] = l0;
        S.Lt[(j0 + 1) * LDT + r0] = l1;
        S.Lt[(j0 + 2) * LDT + r0] = l2;
        S.Lt[(j0 + 3) * LDT + r0] = l3;
        br0 = fmaf(-l3, z3, fmaf(-l2, z2, fmaf(-l1, z1, fmaf(-l0, z0, br0))));
      }
      __syncwarp();
      // P3: rank-4 update of both rows from the published panel; windows shift one block to the left
      {
        const float2 n0 = make_float2(-m0, -m0), n1 = make_float2(-m1, -m1), n2 = make_float2(-m2, -m2),
                     n3 = make_float2(-m3, -m3);
        const float2 k0 = make_float2(-l0, -l0), k1 = make_float2(-l1, -l1), k2 = make_float2(-l2, -l2),
                     k3 = make_float2(-l3, -l3);
        const float* lt = &S.Lt[j0 * LDT + j0 + 4];
#pragma unroll
        for (int ib = 0; ib < NB4 - 1; ib++) {
          if (j0 + 4 + 4 * ib > 63) break;   // warp-uniform
          const float4 v0 = *reinterpret_cast<const float4*>(lt + 0 * LDT + 4 * ib);
          const float4 v1 = *reinterpret_cast<const float4*>(lt + 1 * LDT + 4 * ib);
          const float4 v2 = *reinterpret_cast<const float4*>(lt + 2 * LDT + 4 * ib);
          const float4 v3 = *reinterpret_cast<const float4*>(lt + 3 * LDT + 4 * ib);
          float2 lo = __ffma2_rn(n0, make_float2(v0.x, v0.y), a1[2 * ib + 2]);
          float2 hi = __ffma2_rn(n0, make_float2(v0.z, v0.w), a1[2 * ib + 3]);
          lo = __ffma2_rn(n1, make_float2(v1.x, v1.y), lo);
          hi = __ffma2_rn(n1, make_float2(v1.z, v1.w), hi);
          lo = __ffma2_rn(n2, make_float2(v2.x, v2.y), lo);
          hi = __ffma2_rn(n2, make_float2(v2.z, v2.w), hi);
          a1[2 * ib] = __ffma2_rn(n3, make_float2(v3.x, v3.y), lo);
          a1[2 * ib + 1] = __ffma2_rn(n3, make_float2(v3.z, v3.w), hi);
          if (ib < 7 && j0 + 4 + 4 * ib <= 31) {   // row r0: columns up to 31 only (warp-uniform)
            float2 lo0 = __ffma2_rn(k0, make_float2(v0.x, v0.y), a0[2 * ib + 2]);
            float2 hi0 = __ffma2_rn(k0, make_float2(v0.z, v0.w), a0[2 * ib + 3]);
            lo0 = __ffma2_rn(k1, make_float2(v1.x, v1.y), lo0);
            hi0 = __ffma2_rn(k1, make_float2(v1.z, v1.w), hi0);
            lo0 = __ffma2_rn(k2, make_float2(v2.x, v2.y), lo0);
            hi0 = __ffma2_rn(k2, make_float2(v2.z, v2.w), hi0);
            a0[2 * ib] = __ffma2_rn(k3, make_float2(v3.x, v3.y), lo0);
            a0[2 * ib + 1] = __ffma2_rn(k3, make_float2(v3.z, v3.w), hi0);
          }
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
    __syncwarp();
    if (failed) {   // every lane factored the same blocks: warp-uniform
      if (lane == 0) atomicExch(P.status, 1);
      advance();
      continue;     // Y row untouched; status reports B200ALS_ENOTSPD
    }
    // ---- back substitution  L' y = z  (rows 32..63 first, then rows 0..31 after subtracting the solved part) --------------
    for (int b0 = 32; b0 >= 0; b0 -= 32) {
      const int i = b0 + lane;
      const float ri = S.rs[i];
      float zi = S.zz[i];
      if (b0 == 0) {
        const float* lrow = &S.Lt[lane * LDT + 32];
        float ps = 0.f;
#pragma unroll
        for (int l = 0; l < 32; l += 4) {
          const float4 lv = *reinterpret_cast<const float4*>(lrow + l);
          const float4 yv = *reinterpret_cast<const float4*>(&S.zz[32 + l]);
          ps = fmaf(lv.x, yv.x, fmaf(lv.y, yv.y, fmaf(lv.z, yv.z, fmaf(lv.w, yv.w, ps))));
        }
        zi -= ps;
      }
#pragma unroll 8
      for (int sidx = 31; sidx >= 0; sidx--) {
        const float ys = __shfl_sync(kFull, zi * ri, sidx);
        if (lane < sidx) zi = fmaf(-S.Lt[i * LDT + b0 + sidx], ys, zi);
      }
      S.zz[i] = zi * ri;
      __syncwarp();
    }
    float* y = P.Y + (size_t)row * K;
    if (lane < K / 4) *reinterpret_cast<float4*>(y + lane * 4) = *reinterpret_cast<const float4*>(&S.zz[lane * 4]);
    // ---- loss (the gathered rows come back through L2, four in flight) ------------------------------------------------------
    float l = 0.0f;
    {
      const float4 yv = (lane < 16) ? *reinterpret_cast<const float4*>(&S.zz[lane * 4]) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int jb = 0; jb < n; jb += 4) {
        float4 xv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (jb + u < n && lane < 16) xv[u] = ldg_f4(P.X + (size_t)s_idx[jb + u] * K + lane * 4);
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
    {
      float yy = fmaf(S.zz[lane], S.zz[lane], S.zz[lane + 32] * S.zz[lane + 32]);
      yy = warp_sum(yy);
      if (lane == 0) l = fmaf(implicit ? (float)P.lambda : lam_use, yy, l);
    }
    if (lane == 0) cta_loss += (double)l;
    advance();
  }
  if (lane == 0) P.loss_partials[blockIdx.x] = cta_loss;
}

}  // namespace b200als
