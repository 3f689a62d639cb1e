// topk.cuh -- top-k of the user x item score matrix with per-user and global exclusions.
// Reference: `top_product` src/matrix_top_product.cpp:20-102 (called by find_top_product R/utils.R:31-59 from
// MatrixFactorizationRecommender$predict, R/MatrixFactorizationRecommender.R:24-78):
//   for every user row j: yvec = x.row(j) * y (double), walk the items in increasing index, skip items listed in the
//   user's `not_recommend` row (sorted) or in the global exclude set, keep the k best in a min-heap of (score, index)
//   pairs with the STRICT replacement rule `q.top().first < val`, emit indices (1-based) by decreasing (score, index).
// Here: one CTA = 32 users, items streamed in tiles of 64.  Scores are accumulated in DOUBLE like the reference
// (float x float products are exact in double), so the ranking does not depend on fp32 rounding:
//   * user block and item tile are staged transposed in shared memory as doubles (xs[f][u], ys[f][i]); ranks above
//     128 (up to 256) stage the item tile in two passes of 128 features into the same accumulators;
//   * 256 threads, thread tile 2 users x 4 items, per feature 3 x LDS.128 + 8 DFMA;
//   * selection: warp w owns users 4w..4w+3; a score becomes a candidate only if it beats the user's current k-th
//     score (strictly) -- only candidates pay for the exclusion tests (bitmap; binary search in the user's sorted
//     not_recommend row); the per-user list is kept sorted by decreasing (score, index) in shared memory, which
//     reproduces the heap's tie behaviour exactly (evict the smallest pair, later equal scores never replace).
#pragma once
#include "common.cuh"

namespace b200als {

constexpr int kTopUB = 32;      // users per CTA
constexpr int kTopIT = 64;      // items per tile
constexpr int kTopMaxK = 128;   // top_k limit (list storage)
constexpr int kTopMaxRank = 256;   // the user block is staged whole
constexpr int kTopFC = 128;        // features of an item tile staged at a time (rank > 128: two passes, same accumulators)

struct TopkSmem {
  alignas(16) double xs[kTopMaxRank][kTopUB];       // 64 KB
  alignas(16) double ys[kTopFC][kTopIT];            // 64 KB
  alignas(16) double sc[kTopUB][kTopIT + 2];        // scores of the current tile
  double lscore[kTopUB][kTopMaxK];                  // per-user sorted lists
  int lidx[kTopUB][kTopMaxK];
  int lsize[kTopUB];
};

struct TopkParams {
  const float* x;        // rank x n_user (user embeddings, column-major: user u at x + u*rank)
  const float* y;        // rank x n_item
  long long n_user;
  int n_item, rank, top_k;
  const int32_t* nr_ptr;   // [n_user+1] not_recommend CSR (may be null)
  const int32_t* nr_idx;   // ascending column indices per user
  const uint32_t* exclude_bits;  // n_item bits (may be null)
  double glob_mean;
  int32_t* idx_out;      // n_user x top_k column-major, 1-based, INT_MIN = NA
  double* score_out;     // n_user x top_k column-major, NA_real_ when missing
};

__device__ __forceinline__ bool topk_excluded(const TopkParams& P, long long user, int item) {
  if (P.exclude_bits && ((P.exclude_bits[item >> 5] >> (item & 31)) & 1u)) return true;
  if (P.nr_ptr) {
    int lo = P.nr_ptr[user], hi = P.nr_ptr[user + 1];
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const int v = __ldg(P.nr_idx + mid);
      if (v == item) return true;
      if (v < item) lo = mid + 1; else hi = mid;
    }
  }
  return false;
}

__global__ void __launch_bounds__(256) topk_kernel(TopkParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TopkSmem& S = *reinterpret_cast<TopkSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long u0 = (long long)blockIdx.x * kTopUB;
  const int rank = P.rank, K = P.top_k;
  // ---- user block -> xs[f][u] (double) -----------------------------------------------------------------------
  for (int e = tid; e < kTopUB * rank; e += 256) {
    const int u = e / rank, f = e - u * rank;
    S.xs[f][u] = (u0 + u < P.n_user) ? (double)__ldg(P.x + (size_t)(u0 + u) * rank + f) : 0.0;
  }
  if (tid < kTopUB) S.lsize[tid] = 0;
  const int tu = (tid >> 4) * 2;      // this thread's 2 users  (16 thread-rows x 2)
  const int ti = (tid & 15) * 4;      // this thread's 4 items  (16 thread-cols x 4)
  for (int i0 = 0; i0 < P.n_item; i0 += kTopIT) {
    double acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    for (int f0 = 0; f0 < rank; f0 += kTopFC) {   // features in ascending order, as one pass would take them
      const int fc = min(kTopFC, rank - f0);
      __syncthreads();   // previous tile's scores (first pass) / previous feature chunk consumed, xs ready
      for (int e = tid; e < kTopIT * fc; e += 256) {
        const int i = e / fc, f = e - i * fc;
        S.ys[f][i] = (i0 + i < P.n_item) ? (double)__ldg(P.y + (size_t)(i0 + i) * rank + f0 + f) : 0.0;
      }
      __syncthreads();
#pragma unroll 4
      for (int f = 0; f < fc; f++) {
        const double2 xv = *reinterpret_cast<const double2*>(&S.xs[f0 + f][tu]);
        const double2 y01 = *reinterpret_cast<const double2*>(&S.ys[f][ti]);
        const double2 y23 = *reinterpret_cast<const double2*>(&S.ys[f][ti + 2]);
        acc[0][0] = fma(xv.x, y01.x, acc[0][0]); acc[0][1] = fma(xv.x, y01.y, acc[0][1]);
        acc[0][2] = fma(xv.x, y23.x, acc[0][2]); acc[0][3] = fma(xv.x, y23.y, acc[0][3]);
        acc[1][0] = fma(xv.y, y01.x, acc[1][0]); acc[1][1] = fma(xv.y, y01.y, acc[1][1]);
        acc[1][2] = fma(xv.y, y23.x, acc[1][2]); acc[1][3] = fma(xv.y, y23.y, acc[1][3]);
      }
    }
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) S.sc[tu + a][ti + b] = acc[a][b];
    __syncthreads();
    // ---- selection: warp w handles users 4w .. 4w+3, items in increasing index -------------------------------
    for (int uu = 0; uu < 4; uu++) {
      const int u = warp * 4 + uu;
      const long long user = u0 + u;
      if (user >= P.n_user) break;
      for (int half = 0; half < 2; half++) {
        const int il = half * 32 + lane;
        const int item = i0 + il;
        const double val = S.sc[u][il];
        int size = S.lsize[u];
        const double thr = (size == K) ? S.lscore[u][K - 1] : 0.0;
        bool cand = (item < P.n_item) && (size < K || val > thr);
        unsigned mask = __ballot_sync(kFull, cand);
        while (mask) {
          const int src = __ffs(mask) - 1;
          mask &= mask - 1;
          const double v = __shfl_sync(kFull, val, src);
          const int it = i0 + half * 32 + src;
          size = S.lsize[u];
          // re-test against the (possibly updated) list, then the exclusion lists (lane 0 decides, broadcast)
          int ok = 0;
          if (lane == 0) ok = (size < K || v > S.lscore[u][K - 1]) && !topk_excluded(P, user, it);
          ok = __shfl_sync(kFull, ok, 0);
          if (!ok) continue;
          // position = number of entries with a strictly larger score (equal scores have smaller indices: after us)
          int pos = 0;
          for (int j = lane; j < size; j += 32) pos += (S.lscore[u][j] > v) ? 1 : 0;
          pos = warp_sum(pos);
          const int new_size = min(size + 1, K);
          // make room at `pos`: shift [pos, new_size-1) down by one, in place, 32 entries at a time from the tail
          for (int top = new_size - 1; top > pos; top -= 32) {
            const int j = top - lane;            // destination index
            const bool mv = (j > pos);
            double sj = 0.0;
            int ij = 0;
            if (mv) { sj = S.lscore[u][j - 1]; ij = S.lidx[u][j - 1]; }
            __syncwarp();
            if (mv) { S.lscore[u][j] = sj; S.lidx[u][j] = ij; }
            __syncwarp();
          }
          if (lane == 0) {
            S.lscore[u][pos] = v;
            S.lidx[u][pos] = it;
            S.lsize[u] = new_size;
          }
          __syncwarp();
        }
      }
    }
  }
  __syncthreads();
  // ---- output (column-major n_user x top_k, 1-based, NA beyond the list) -------------------------------------------
  for (int e = tid; e < kTopUB * K; e += 256) {
    const int u = e / K, r = e - u * K;
    const long long user = u0 + u;
    if (user >= P.n_user) continue;
    const bool have = r < S.lsize[u];
    P.idx_out[(size_t)r * P.n_user + user] = have ? (S.lidx[u][r] + 1) : (int32_t)0x80000000;
    P.score_out[(size_t)r * P.n_user + user] =
        have ? (S.lscore[u][r] + P.glob_mean) : __longlong_as_double(0x7FF00000000007A2ll);   // R's NA_real_
  }
}

}  // namespace b200als
