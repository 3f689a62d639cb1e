// als_resident.cuh -- the hot kernel: fixed-step CG half-iteration with the row's gathered factor
// tile RESIDENT IN REGISTERS for the whole solve (rank 128, fp32, rows with 1..80 non-zeros).
//
// Reference semantics: cg_solver_implicit (inst/include/wrmf_implicit.hpp:8-32), cg_solver_explicit
// (inst/include/wrmf_explicit.hpp:8-31) and the surrounding column loop (wrmf_implicit.hpp:175-282,
// wrmf_explicit.hpp:71-146).  What is different from the reference's shape:
//   * one CTA of 4 warps owns a row; warp w holds gathered rows j = w, w+4, ... (<= 20 of them) as 10 register
//     GROUPS x 2 half-warps: lane (g = lane >> 4, l = lane & 15) keeps, for each group, 8 features of ONE row
//     (two float4) => the 40 KB tile X_nnz is read from HBM exactly once per row and then lives in 80
//     registers/thread for all 2(s+1) mat-vecs;
//   * the two mat-vecs of a CG step are fused into one sweep over the registers:
//       u_j = x_j . p   (8 FMA/lane per row, then a transposing halving reduction over the 16 lanes of the
//                        half-warp: 11 SHFL for its 10 rows -- the other half-warp reduces its own 10 at the same time)
//       w_j = (c_j - 1) u_j ;  acc += w_j x_j   (80 FMA/lane), one exchange between the two half-warps (4 SHFL),
//       then a 4-way cross-warp sum in smem;
//   * the NEXT row's tile (and its warm-start y) is prefetched while this row computes: every warp stages
//     its own 20 gathered rows with one 512-byte cp.async.bulk (TMA engine, SASS UBLKCP) per owner lane
//     into its own 10.5 KB shared-memory slot, completion by a per-warp mbarrier (complete_tx); the CSR
//     indices / confidences of the row after that and the row pointers of the one after that travel
//     through registers => no thread ever waits on a dependent HBM load chain, the four warps do
//     identical work and meet only at the one cross-warp sum per sweep;
//   * XtX p: either the session has rotated both factor matrices into the eigenbasis of XtX (kDiag:
//     XtX p = d (.) p, 4 FMA/lane -- see eig.cuh), or (kFullG) each warp multiplies a 32-column slab
//     of XtX from L1/L2 and the slabs are summed by the same cross-warp reduction;
//   * the loss term X_nnz' y is obtained from the u vectors already computed
//     (X_nnz' y = X_nnz' x0 + sum_k alpha_k X_nnz' p_k) instead of a fifth sweep;
//   * alpha = rsold / p'Ap and beta are formed by the fp32 hardware reciprocal, <= 2 ulp (the reference divides
//     in double and rounds to T, wrmf_implicit.hpp:18,23,28); rounding-level difference, covered by the fp32
//     tolerance;
//   * p'XtX p rides in the per-warp scalar of the cross-warp exchange and the loss terms are summed per lane in
//     fp64, so a CG step has ONE warp-wide scalar reduction that is not overlapped with a sweep (|r|^2).
// Algorithmic HBM bytes per row: 4nk + 8n + 4 + 4k + 4k  (42,628 B at n = 80, k = 128).
#pragma once
#include "common.cuh"

namespace b200als {

constexpr int kResK = 128;       // rank handled by this kernel
constexpr int kResWarps = 4;
constexpr int kResThreads = kResWarps * 32;
constexpr int kResIPW = 20;      // gathered rows per warp (= float4 registers per lane)
constexpr int kResGroups = 10;   // register groups per lane: group q of half-warp g holds tile slot 2 qn + g
constexpr int kResMaxN = kResWarps * kResIPW;  // 80
constexpr int kResRowBytes = kResK * 4;        // 512

struct ResidentParams {
  const int32_t* ptr;
  const int32_t* idx;
  const float* val;
  const float* X;      // 128 x n_src
  float* Y;            // 128 x n_targets
  const float* diag;   // [128]  eigenvalues of XtX (+lambda)          (kDiag)
  const float* G;      // 128 x 128 XtX + lambda I                      (kFullG)
  int feedback;
  int cg_steps;
  int dynamic_lambda;
  float lambda;
  const int32_t* row_list;  // rows with 1 <= nnz <= kResMaxN (nullptr: all rows 0..n_list-1 qualify)
  int n_list;
  const int* n_list_dev;    // when non-null the list length is read from device memory (pipelined calls)
  int ptr_base;             // ptr[] values are offsets into a buffer that starts at this absolute offset
  int row_begin;            // without a row list: solve rows [row_begin, row_begin + n_list)
  double* loss_partials;    // [gridDim.x]
};

struct __align__(128) ResidentSmem {
  // per-warp tile slot: rows 0..19 = this warp's gathered rows (item j = w + 4q), row 20 = the warm-start y
  float tile[kResWarps][(kResIPW + 1) * kResK];
  float vbuf[2][kResWarps][kResK];         // cross-warp partial sums (double buffered per sweep)
  float sbuf[2][kResWarps];                // per-warp scalar partials riding the same exchange (p'Ap terms)
  float wbuf[kResWarps][32];               // per-warp w_j broadcast: [half-warp][block of 5, padded to 8]
  uint64_t bar[kResWarps];                 // one mbarrier per warp slot
  double red[32];
};

// Transposing halving reduction: on entry lane L holds t[0..N) partial dot products, on exit the lane
// whose bits select slot q holds the full 32-lane sum of t[q]; returns it (0 for padding lanes).
// Levels with kSwapFree = true assume the registers of lanes whose bit M is set were LOADED with their two
// halves swapped (see the layout note above resident_owner_group), so "keep the low registers, send the high registers" is the
// same instruction stream for every lane -- no selects.  Odd-sized levels fall back to selects.
template <int N>
struct Halver {
  template <int M, int kSwapFreeLevels>
  static __device__ __forceinline__ float run(float (&t)[N], int lane) {
    constexpr int H = (N + 1) / 2;
    float o[H];
    if constexpr (kSwapFreeLevels > 0 && (N % 2 == 0)) {
#pragma unroll
      for (int v = 0; v < H; v++) o[v] = t[v] + __shfl_xor_sync(kFull, t[v + H], M);
    } else {
      const bool upper = (lane & M) != 0;
#pragma unroll
      for (int v = 0; v < H; v++) {
        const float lo = t[v];
        const float hi = (v + H < N) ? t[v + H] : 0.0f;
        const float send = upper ? lo : hi;
        const float keep = upper ? hi : lo;
        o[v] = keep + __shfl_xor_sync(kFull, send, M);
      }
    }
    if constexpr (M == 1) {
      return o[0];
    } else {
      return Halver<H>::template run<M / 2, (kSwapFreeLevels > 0 ? kSwapFreeLevels - 1 : 0)>(o, lane);
    }
  }
};
// Layout of a warp's 20 tile slots (slot s = gathered row w + 4 s of the CSR row, 512 B each in shared memory):
// half-warp g = lane >> 4 holds the slots of its parity, s = 2 qn + g, qn = 0..9.  Register group q = 5 h + r of a
// lane holds NATURAL group qn = 5 (h ^ b3) + r (b3 = lane bit 3: the first halving level 10 -> 5 is select-free
// because lanes with bit 3 set keep their two blocks of five swapped), as two float4: register c of the group is
// the 16-byte chunk (c ^ g) * 16 + l of the row (l = lane & 15), so that "keep register 0, send register 1" in the
// exchange between the half-warps leaves lane L with chunk L of the sum -- the natural layout of x, r, p.
//
// group (within its half-warp) owned by a lane after Halver<10>::run<8, 1>: 10 -> 5 -> 3 -> 2 -> 1 over lane bits 3..0
__device__ __forceinline__ int resident_owner_group(int lane, int& in5_out) {
  const int b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1, b0 = lane & 1;
  const int in3 = 2 * b1 + b0;          // index within the group of 3 (valid < 3)
  const int in5 = 3 * b2 + in3;         // index within the group of 5 (valid < 5)
  in5_out = in5;
  if (in3 >= 3 || in5 >= 5) return -1;
  return 5 * b3 + in5;
}
// lane (within a half-warp) that owns natural group qn (inverse of resident_owner_group)
__host__ __device__ constexpr int resident_group_owner_lane(int qn) {
  const int b3 = qn / 5, in5 = qn % 5, b2 = in5 / 3, in3 = in5 % 3;
  return 8 * b3 + 4 * b2 + 2 * (in3 / 2) + (in3 % 2);
}

// Packed fp32 math (Blackwell FFMA2 / FMUL2 / FADD2: one instruction per register PAIR): the float4 held per
// gathered row is two aligned pairs, so the dot products and the axpy take half the issue slots.
__device__ __forceinline__ float2 lo2(const float4& a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ float2 hi2(const float4& a) { return make_float2(a.z, a.w); }
__device__ __forceinline__ float4 join4(const float2& l, const float2& h) { return make_float4(l.x, l.y, h.x, h.y); }
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  float2 m = __fmul2_rn(lo2(a), lo2(b));
  m = __ffma2_rn(hi2(a), hi2(b), m);
  return m.x + m.y;
}
// a * s + c  /  a (.) b + c on float4 operands
__device__ __forceinline__ float4 axpy4(float s, const float4& a, const float4& c) {
  const float2 ss = make_float2(s, s);
  return join4(__ffma2_rn(ss, lo2(a), lo2(c)), __ffma2_rn(ss, hi2(a), hi2(c)));
}
__device__ __forceinline__ float4 fma4(const float4& a, const float4& b, const float4& c) {
  return join4(__ffma2_rn(lo2(a), lo2(b), lo2(c)), __ffma2_rn(hi2(a), hi2(b), hi2(c)));
}
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) {
  return join4(__fadd2_rn(lo2(a), lo2(b)), __fadd2_rn(hi2(a), hi2(b)));
}
__device__ __forceinline__ float4 neg4(const float4& a) { return make_float4(-a.x, -a.y, -a.z, -a.w); }

// ---- one CG sweep, in two halves -------------------------------------------------------------------------
// (A) u_j = x_j . vec for this warp's 20 rows; the owner lane of a slot returns the full sum (0 on padding lanes).
//     vec is in the natural layout (lane L holds chunk L); register 1 of a group needs the chunk of lane L ^ 16.
__device__ __forceinline__ float sweep_dots(const float4 (&xt)[kResIPW], const float4& vec) {
  const float4 vo = make_float4(__shfl_xor_sync(kFull, vec.x, 16), __shfl_xor_sync(kFull, vec.y, 16),
                                __shfl_xor_sync(kFull, vec.z, 16), __shfl_xor_sync(kFull, vec.w, 16));
  float t[kResGroups];
#pragma unroll
  for (int q = 0; q < kResGroups; q++) {
    float2 m = __fmul2_rn(lo2(xt[2 * q]), lo2(vec));
    m = __ffma2_rn(hi2(xt[2 * q]), hi2(vec), m);
    m = __ffma2_rn(lo2(xt[2 * q + 1]), lo2(vo), m);
    m = __ffma2_rn(hi2(xt[2 * q + 1]), hi2(vo), m);
    t[q] = m.x + m.y;
  }
  return Halver<kResGroups>::template run<8, 1>(t, lane_id());
}
// (B) acc = sum_j wq_j x_j over the warp's rows (+ this warp's slab of XtX * vec when kFullG), 4-way cross-warp
//     sum through shared memory; `spart` (a per-warp scalar, identical in all lanes) is summed across the four
//     warps on the way and returned in `ssum`.
template <bool kFullG>
__device__ __forceinline__ float4 sweep_apply(const float4 (&xt)[kResIPW], float wq, int wpos, const float4& vec,
                                              int gmode /* 0: none, 1: acc - G vec, 2: acc + G vec */,
                                              ResidentSmem& S, int sweep, const float* __restrict__ G, float spart,
                                              float& ssum) {
  const int lane = lane_id(), w = warp_id();
  if (wpos >= 0) S.wbuf[w][wpos] = wq;
  __syncwarp();
  // four independent accumulator chains: (register 0 | register 1 of a group) x (low | high pair)
  float2 a0l = make_float2(0.f, 0.f), a0h = a0l, a1l = a0l, a1h = a0l;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const float* wp = &S.wbuf[w][(lane & 24) ^ (h * 8)];   // my half-warp's block of natural groups 5 (h ^ b3) ..
    const float4 w03 = *reinterpret_cast<const float4*>(wp);
    const float ws[5] = {w03.x, w03.y, w03.z, w03.w, wp[4]};
#pragma unroll
    for (int r = 0; r < 5; r++) {
      const int q = 5 * h + r;
      const float2 s2 = make_float2(ws[r], ws[r]);
      a0l = __ffma2_rn(s2, lo2(xt[2 * q]), a0l);
      a0h = __ffma2_rn(s2, hi2(xt[2 * q]), a0h);
      a1l = __ffma2_rn(s2, lo2(xt[2 * q + 1]), a1l);
      a1h = __ffma2_rn(s2, hi2(xt[2 * q + 1]), a1h);
    }
  }
  // the other half-warp's rows: keep register 0 (chunk L), send register 1 (chunk L ^ 16)
  float4 acc = join4(__fadd2_rn(a0l, make_float2(__shfl_xor_sync(kFull, a1l.x, 16), __shfl_xor_sync(kFull, a1l.y, 16))),
                     __fadd2_rn(a0h, make_float2(__shfl_xor_sync(kFull, a1h.x, 16), __shfl_xor_sync(kFull, a1h.y, 16))));
  __syncwarp();
  if constexpr (kFullG) {
    // this warp's slab of XtX * vec: columns j in [32w, 32w+32); vec_j lives in lane j/4 (component j%4)
    if (gmode != 0) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
      for (int g4 = 0; g4 < 8; g4++) {
        const int src = w * 8 + g4;
        const float vj[4] = {__shfl_sync(kFull, vec.x, src), __shfl_sync(kFull, vec.y, src),
                             __shfl_sync(kFull, vec.z, src), __shfl_sync(kFull, vec.w, src)};
#pragma unroll
        for (int c = 0; c < 4; c++) g = axpy4(vj[c], ldg_f4(G + (size_t)(src * 4 + c) * kResK + lane * 4), g);
      }
      acc = axpy4((gmode == 1) ? -1.0f : 1.0f, g, acc);
      spart += warp_sum(dot4(vec, g));   // vec' (G_slab vec): this warp's share of vec' XtX vec
    }
  }
  float* vb = &S.vbuf[sweep & 1][0][0];
  *reinterpret_cast<float4*>(vb + w * kResK + lane * 4) = acc;
  if (lane == 0) S.sbuf[sweep & 1][w] = spart;
  __syncthreads();
  const float4 v0 = *reinterpret_cast<const float4*>(vb + lane * 4);
  const float4 v1 = *reinterpret_cast<const float4*>(vb + 1 * kResK + lane * 4);
  const float4 v2 = *reinterpret_cast<const float4*>(vb + 2 * kResK + lane * 4);
  const float4 v3 = *reinterpret_cast<const float4*>(vb + 3 * kResK + lane * 4);
  const float4 sv = *reinterpret_cast<const float4*>(&S.sbuf[sweep & 1][0]);
  ssum = ((sv.x + sv.y) + sv.z) + sv.w;
  return add4(add4(add4(v0, v1), v2), v3);   // fixed order: identical in all four warps
}

// kCtas: resident CTAs per SM the register allocation is sized for (3: 168 registers, no spills; 4: 128
// registers with ~170 B of spills per thread that stay in L1).
template <bool kFullG, int kStage, int kCtas>
__global__ void __launch_bounds__(kResThreads, kCtas) als_cg_resident_kernel(ResidentParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ResidentSmem& S = *reinterpret_cast<ResidentSmem*>(smem_raw);
  const int lane = lane_id(), w = warp_id(), tid = threadIdx.x;
  const bool implicit = (P.feedback == 0);
  const int stride = gridDim.x;
  int in5;
  const int og = resident_owner_group(lane, in5);  // natural group (within my half-warp) this lane owns after the halving reduce
  const int slot = (og >= 0) ? 2 * og + (lane >> 4) : -1;             // ... = this tile slot of the warp
  const int wpos = (og >= 0) ? (lane & 24) + in5 : -1;                // where its w_j goes in S.wbuf[w]
  const int my_j = (slot >= 0) ? (w + kResWarps * slot) : (1 << 30);  // its position within the CSR row
  float* my_tile = &S.tile[w][0];
  uint64_t* my_bar = &S.bar[w];

  const long long n_list = P.n_list_dev ? (long long)__ldg(P.n_list_dev) : (long long)P.n_list;
  auto valid = [&](int i) -> bool { return (long long)blockIdx.x + (long long)i * stride < n_list; };
  auto row_of = [&](int i) -> int {  // i-th row of this CTA (caller checks valid(i))
    const long long t = (long long)blockIdx.x + (long long)i * stride;
    return P.row_list ? __ldg(P.row_list + t) : (int)t + P.row_begin;
  };
  // Every warp stages ITS OWN 20 gathered rows (+ its own copy of the warm-start y) with one 512-byte bulk
  // copy per owner lane, so the four warps do identical work and never wait for a producer warp.
  auto issue_tile = [&](int row, int n, int my_idx) {
    if constexpr (kStage == 0) {
      const int nw = (n > w) ? (n - w + kResWarps - 1) / kResWarps : 0;
      if (lane == 0) mbar_expect_tx(my_bar, (uint32_t)(nw + 1) * kResRowBytes);
      __syncwarp();
      if (my_j < n) bulk_g2s(my_tile + slot * kResK, P.X + (size_t)my_idx * kResK, kResRowBytes, my_bar);
      if (lane == 0) bulk_g2s(my_tile + kResIPW * kResK, P.Y + (size_t)row * kResK, kResRowBytes, my_bar);
    } else {
      // half-warp g copies the slots of its parity: 16 lanes x 16 B = one half of a 512-byte row per instruction
      const float* xl = P.X + (lane & 15) * 4;
      float* tl = my_tile + (lane >> 4) * kResK + (lane & 15) * 4;
      if (n == kResMaxN) {
#pragma unroll
        for (int qn = 0; qn < kResGroups; qn++) {
          const int src = __shfl_sync(kFull, my_idx, resident_group_owner_lane(qn) | (lane & 16));
          cp_async_16(tl + qn * 2 * kResK, xl + (size_t)src * kResK);
          cp_async_16(tl + qn * 2 * kResK + 64, xl + (size_t)src * kResK + 64);
        }
      } else {
        const int nw = (n > w) ? (n - w + kResWarps - 1) / kResWarps : 0;
        const int nwh = (nw - (lane >> 4) + 1) >> 1;   // slots 2 qn + g < nw
#pragma unroll
        for (int qn = 0; qn < kResGroups; qn++) {
          const int src = __shfl_sync(kFull, my_idx, resident_group_owner_lane(qn) | (lane & 16));
          if (qn < nwh) {
            cp_async_16(tl + qn * 2 * kResK, xl + (size_t)src * kResK);
            cp_async_16(tl + qn * 2 * kResK + 64, xl + (size_t)src * kResK + 64);
          }
        }
      }
      cp_async_16(my_tile + kResIPW * kResK + lane * 4, P.Y + (size_t)row * kResK + lane * 4);
      cp_async_mbar_arrive_noinc(my_bar);
    }
  };

  if (tid < kResWarps) mbar_init(&S.bar[tid], kStage == 0 ? 1 : 32);
  if (tid == 0) mbar_fence_init();
  __syncthreads();

  // ---- software pipeline state (per warp, in registers; every global load is consumed one row later) ----
  //   row i   : n0, cq (confidence of my slot)                         -- tile landing / landed
  //   row i+1 : rid1, n1, idx1, val1                                    -- bulk copies issued at the start of row i
  //   row i+2 : rid2, p2, n2 known; idx2/val2 loaded during row i
  //   row i+3 : rid3 known; CSR range loaded during row i
  //   row i+4 : id loaded during row i (only when a row list is used)
  int rid0 = -1, rid1 = -1, rid2 = -1, rid3 = -1;
  int n0 = 0, n1 = 0, n2 = 0, p2 = 0, idx1 = 0;
  float cq = 0.f, val1 = 0.f;
  if (valid(0)) {
    rid0 = row_of(0);
    const int p = __ldg(P.ptr + rid0) - P.ptr_base;
    n0 = __ldg(P.ptr + rid0 + 1) - P.ptr_base - p;
    int idx0 = 0;
    if (my_j < n0) { idx0 = __ldg(P.idx + p + my_j); cq = __ldg(P.val + p + my_j); }
    issue_tile(rid0, n0, idx0);
  }
  if (valid(1)) {
    rid1 = row_of(1);
    const int p = __ldg(P.ptr + rid1) - P.ptr_base;
    n1 = __ldg(P.ptr + rid1 + 1) - P.ptr_base - p;
    if (my_j < n1) { idx1 = __ldg(P.idx + p + my_j); val1 = __ldg(P.val + p + my_j); }
  }
  if (valid(2)) { rid2 = row_of(2); p2 = __ldg(P.ptr + rid2) - P.ptr_base; n2 = __ldg(P.ptr + rid2 + 1) - P.ptr_base - p2; }
  if (valid(3)) rid3 = row_of(3);

  float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!kFullG && implicit) dg = ldg_f4(P.diag + lane * 4);
  const float4 ndg = neg4(dg);
  double lane_loss = 0.0;   // summed over this lane's rows in fp64; one block reduction at the end
  int sweep = 0;

  for (int i = 0; valid(i); i++) {
    const int n = n0;
    // ---- my 20 rows: shared memory -> registers ----------------------------------------------------------
    mbar_wait(my_bar, (uint32_t)(i & 1));
    // register c of group q = 5 h + r: chunk (c ^ g) * 16 + l of tile slot 2 (5 (h ^ b3) + r) + g  (see the layout
    // note above resident_owner_group): four lane-dependent bases, compile-time offsets from there
    float4 xt[kResIPW];
    {
      const int g = lane >> 4, l = lane & 15, ob = ((lane >> 3) & 1) * 5;
      const float* tb = my_tile + g * kResK;
      const float* cb[2] = {tb + (g * 16 + l) * 4, tb + ((g ^ 1) * 16 + l) * 4};
      const int nb[2] = {ob, 5 - ob};   // first natural group of block h
      if (n == kResMaxN) {   // full tile: no padding anywhere
#pragma unroll
        for (int q = 0; q < kResGroups; q++)
#pragma unroll
          for (int c = 0; c < 2; c++)
            xt[2 * q + c] = *reinterpret_cast<const float4*>(cb[c] + (nb[q / 5] + q % 5) * 2 * kResK);
      } else {
        const int nw = (n > w) ? (n - w + kResWarps - 1) / kResWarps : 0;   // slots of this warp that hold a row
        const int nwh = (nw - g + 1) >> 1;                                  // ... of my half-warp's parity
#pragma unroll
        for (int q = 0; q < kResGroups; q++) {
          const bool ok = (q % 5 < nwh - nb[q / 5]);
#pragma unroll
          for (int c = 0; c < 2; c++)
            xt[2 * q + c] = ok ? *reinterpret_cast<const float4*>(cb[c] + (nb[q / 5] + q % 5) * 2 * kResK)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    float4 x = *reinterpret_cast<const float4*>(my_tile + kResIPW * kResK + lane * 4);
    __syncwarp();  // my slot is free again
    // ---- prefetch (nothing here waits on memory) ----------------------------------------------------------
    if (valid(i + 1)) issue_tile(rid1, n1, idx1);
    int idx2 = 0, p3 = 0, p3e = 0, rid4 = -1;
    float val2 = 0.f;
    if (valid(i + 2) && my_j < n2) { idx2 = ld_pinned_i32(P.idx + p2 + my_j); val2 = ld_pinned_f32(P.val + p2 + my_j); }
    if (valid(i + 3)) { p3 = ld_pinned_i32(P.ptr + rid3); p3e = ld_pinned_i32(P.ptr + rid3 + 1); }
    if (valid(i + 4)) rid4 = P.row_list ? ld_pinned_i32(P.row_list + ((long long)blockIdx.x + (long long)(i + 4) * stride))
                                        : row_of(i + 4);
    // ---- CG ---------------------------------------------------------------------------------------------
    // Same iterates as cg_solver_implicit / cg_solver_explicit, with the dependent chain shortened:
    //   * p'Ap = p'(XtX p) + sum_j (c_j-1) u_j^2   (u = X_nnz' p is already reduced per row; the XtX term is
    //     d.p^2 in the eigenbasis, lambda p.p for explicit feedback) -- no dot(p, Ap) after the exchange;
    //   * X_nnz' p_new = X_nnz' r_new + beta X_nnz' p_old, so the next sweep's row dots start from r_new while the
    //     |r_new|^2 reduction and the division for beta are still in flight.
    const float lam_use = implicit ? P.lambda : (P.lambda * (P.dynamic_lambda ? (float)n : 1.0f));
    const bool mine = (my_j < n);
    float dummy;
    float u_p = sweep_dots(xt, x);   // u_x = X_nnz' x0 (owner lanes)
    float uy = u_p;                  // running X_nnz' y for the loss
    float4 v = sweep_apply<kFullG>(xt, implicit ? (cq - (cq - 1.0f) * u_p) : (cq - u_p), wpos, x,
                                   (kFullG && implicit) ? 1 : 0, S, sweep++, P.G, 0.0f, dummy);
    float4 r;
    if (implicit) {
      if (kFullG) r = v;
      else r = fma4(ndg, x, v);          // v - d (.) x
    } else {
      r = axpy4(-lam_use, x, v);         // v - lambda_use x
    }
    float4 p = r;
    float rsold = warp_sum(dot4(r, r));
    u_p = sweep_dots(xt, p);   // issued with the |r|^2 reduction, not after the branch on its result
    // Guard the reference does not have (wrmf_implicit.hpp:23 computes rsold / (p'Ap) = 0/0 -> NaN once a row
    // has converged exactly, e.g. when the same half-iteration is repeated): a zero residual skips the loop.
    const int n_steps = (rsold > 0.0f) ? P.cg_steps : 0;
#pragma unroll 1
    for (int it = 0; it < n_steps; it++) {
      // p' XtX p without the tile part: every warp holds the whole p, so each adds a QUARTER of its lanes' terms to
      // the per-warp scalar that the cross-warp exchange sums anyway -- one warp reduction per step instead of two,
      // interleaved with the sweep below instead of ahead of it.
      float gterm;
      if (implicit) gterm = kFullG ? 0.0f : 0.25f * dot4(p, fma4(dg, p, make_float4(0.f, 0.f, 0.f, 0.f)));
      else gterm = (0.25f * lam_use) * dot4(p, p);
      const float cw = implicit ? (cq - 1.0f) : 1.0f;
      const float spart = warp_sum((mine ? cw * u_p * u_p : 0.0f) + gterm);
      float ssum;
      v = sweep_apply<kFullG>(xt, cw * u_p, wpos, p, (kFullG && implicit) ? 2 : 0, S, sweep++, P.G, spart, ssum);
      float4 Ap;
      if (implicit) {
        if (kFullG) Ap = v;
        else Ap = fma4(dg, p, v);
      } else {
        Ap = axpy4(lam_use, p, v);
      }
      const float pAp = ssum;
      // rsold / pAp and rsnew / rsold by the hardware reciprocal (div.approx, <= 2 ulp; the reference divides in
      // double and rounds, wrmf_implicit.hpp:18,23,28): two instructions instead of the exact-division chain with
      // its slow-path branch on the critical path of every step
      const float a = (pAp != 0.0f) ? __fdividef(rsold, pAp) : 0.0f;
      x = axpy4(a, p, x);
      r = axpy4(-a, Ap, r);
      uy = fmaf(a, u_p, uy);
      if (it + 1 == n_steps) break;                       // nothing after the last step needs |r|^2 (cf. :26-28)
      const float rsnew = warp_sum(dot4(r, r));           // in flight ...
      const float u_r = sweep_dots(xt, r);                // ... while the next sweep's row dots run
      if (rsnew < (float)B200ALS_CG_TOL) break;           // identical in all four warps (same data, same order)
      const float bt = __fdividef(rsnew, rsold);
      p = axpy4(bt, p, r);
      u_p = fmaf(bt, u_p, u_r);
      rsold = rsnew;
    }
    if (w == 0) *reinterpret_cast<float4*>(P.Y + (size_t)rid0 * kResK + lane * 4) = x;
    // ---- loss ---------------------------------------------------------------------------------------------
    {   // per-lane terms: owner lanes add their rating's term, warp 0 adds its features' share of lambda |y|^2
      float l = 0.0f;
      if (my_j < n) {
        const float d = implicit ? (1.0f - uy) : (cq - uy);
        l = implicit ? d * d * cq : d * d;
      }
      if (w == 0) l = fmaf(lam_use, dot4(x, x), l);
      lane_loss += (double)l;
    }
    // ---- advance the pipeline ------------------------------------------------------------------------------
    n0 = n1; cq = val1; rid0 = rid1;
    n1 = n2; idx1 = idx2; val1 = val2; rid1 = rid2;
    p2 = p3 - P.ptr_base; n2 = p3e - p3; rid2 = rid3;
    rid3 = rid4;
  }
  const double tot = block_sum_double(lane_loss, S.red);
  if (tid == 0) P.loss_partials[blockIdx.x] = tot;
}

}  // namespace b200als
