// rotate_tc.cuh -- Z = X * R (n x 128 times 128 x 128, in place allowed) on tcgen05 tensor cores with the 3xTF32
// split (fp32-grade result): the change of basis of eig.cuh for large factor matrices (10 M x 128 at C3).
//   D[128 rows x 128] = A[128 rows x K=128] * B,  A = row tile of X (row-major = K-major, staged as it lies),
//   B = R as N x K K-major = R' (passed pre-transposed), resident in shared memory for the whole kernel (hi + lo).
// Per row tile: four K-quarters of 32 features are staged (hi/lo) into a double-buffered A slot, one thread issues
// 3 x 4 `tcgen05.mma.kind::tf32` (M128 N128 K8) per quarter; hi*hi accumulates in one TMEM accumulator, the two cross
// terms in a second one (their truncation then happens at a 1000x smaller magnitude, see gram_tc.cuh).  TMEM holds
// two such accumulator pairs (all 512 columns): while the tensor core works on tile i the CUDA cores drain tile i-1
// (`tcgen05.ld`), add the pair and store the rows.  One persistent CTA per SM; HBM-bound by design (1 KB moved per
// row for 98 kflop).
#pragma once
#include "gram_tc.cuh"

namespace b200als {

constexpr int kRtLBOA = 144;                     // A: K-adjacent core matrices 144 B apart (bank spread for the stores)
constexpr int kRtSBOA = 8 * kRtLBOA;             // 8 core matrices (32 features) per 8-row group
constexpr int kRtATile = 16 * kRtSBOA;           // 128 rows x 32 features, hi or lo
constexpr int kRtLBOB = 128;
constexpr int kRtSBOB = 32 * 128 + 16;           // B: 32 core matrices (128 features of K) per 8-column group
constexpr int kRtBTile = 16 * kRtSBOB;           // 128 (n) x 128 (k), hi or lo

struct RotTcSmem {
  alignas(128) unsigned char B[2][kRtBTile];        // [hi/lo]
  alignas(128) unsigned char A[2][2][kRtATile];     // [buffer][hi/lo]
  uint64_t mma_done[2];
  uint64_t acc_ready[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t rt_desc(const void* p, int lbo, int sbo) {
  return (uint64_t)((smem_u32(p) & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Rt[n][k] = R[k][n]
__global__ void transpose_128_kernel(const float* __restrict__ R, float* __restrict__ Rt) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < kTcK * kTcK) Rt[(e % kTcK) * kTcK + (e / kTcK)] = R[e];
}

// X and Z may be the SAME buffer (in-place rotation, how the engine calls it): neither is __restrict__; every row is read
// (one stage ahead) before it is written, by the CTA that writes it.
__global__ void __launch_bounds__(128) rotate_tc_kernel(const float* X, float* Z, const float* __restrict__ Rt, long long n) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  RotTcSmem& S = *reinterpret_cast<RotTcSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&S.mma_done[0], 1);
    mbar_init(&S.mma_done[1], 1);
    mbar_init(&S.acc_ready[0], 1);
    mbar_init(&S.acc_ready[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- B = R' (rows n, K contiguous), hi/lo, resident -------------------------------------------------------------
  for (int e = tid; e < kTcK * (kTcK / 4); e += 128) {
    const int nn = e / (kTcK / 4), c4 = e % (kTcK / 4);
    const float4 v = __ldg(reinterpret_cast<const float4*>(Rt + (size_t)nn * kTcK) + c4);
    const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
    const int off = (nn >> 3) * kRtSBOB + c4 * kRtLBOB + (nn & 7) * 16;
    *reinterpret_cast<float4*>(&S.B[0][off]) = hi;
    *reinterpret_cast<float4*>(&S.B[1][off]) = lo;
  }
  fence_proxy_async();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;

  const long long n_tiles = (n + 127) / 128;
  uint32_t ph_buf[2] = {0, 0}, ph_acc[2] = {0, 0};
  long long stage = 0;   // A-buffer use counter of this CTA
  long long prev_tile = -1;
  int it = 0;            // local tile counter (TMEM pair = it & 1)

  // drain TMEM pair `tb` (hi'hi + cross), store the 128 rows of tile `tile`
  auto drain = [&](int tb, long long tile) {
    mbar_wait(&S.acc_ready[tb], ph_acc[tb]);
    ph_acc[tb] ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long row = tile * 128 + warp * 32 + lane;
#pragma unroll
    for (int c0 = 0; c0 < kTcK; c0 += 32) {
      uint32_t r0[32], r1[32];
      const uint32_t t0 = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tb * 256 + c0);
      const uint32_t t1 = t0 + kTcK;
#define B200ALS_TMEM_LD32(R, ADDR)                                                                                        \
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                             \
                   "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                              \
                   "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"              \
                   : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]),      \
                     "=r"(R[8]), "=r"(R[9]), "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15]), \
                     "=r"(R[16]), "=r"(R[17]), "=r"(R[18]), "=r"(R[19]), "=r"(R[20]), "=r"(R[21]), "=r"(R[22]),           \
                     "=r"(R[23]), "=r"(R[24]), "=r"(R[25]), "=r"(R[26]), "=r"(R[27]), "=r"(R[28]), "=r"(R[29]),           \
                     "=r"(R[30]), "=r"(R[31])                                                                            \
                   : "r"(ADDR)                                                                                           \
                   : "memory")
      B200ALS_TMEM_LD32(r0, t0);
      B200ALS_TMEM_LD32(r1, t1);
#undef B200ALS_TMEM_LD32
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < n) {
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          float4 o;
          o.x = __uint_as_float(r0[c + 0]) + __uint_as_float(r1[c + 0]);
          o.y = __uint_as_float(r0[c + 1]) + __uint_as_float(r1[c + 1]);
          o.z = __uint_as_float(r0[c + 2]) + __uint_as_float(r1[c + 2]);
          o.w = __uint_as_float(r0[c + 3]) + __uint_as_float(r1[c + 3]);
          *reinterpret_cast<float4*>(Z + (size_t)row * kTcK + c0 + c) = o;
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };

  // global loads of a quarter (8 x 16 B per thread: 8 lanes cover the 128 contiguous bytes of a row); issued one
  // stage ahead so their latency hides behind the staging / MMA issue / drain of the current stage
  auto load_quarter = [&](long long tile, int q, float4 (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int rr = warp * 32 + i * 4 + (lane >> 3), c4 = lane & 7;
      const long long row = tile * 128 + rr;
      v[i] = (tile < n_tiles && row < n) ? *(reinterpret_cast<const float4*>(X + (size_t)row * kTcK + q * 32) + c4)   /* plain load: Z may alias X */
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  // The loads run TWO quarters ahead of the staging (32 KB in flight per SM instead of 16 KB: with one CTA of 4 warps per
  // SM the depth of this pipeline, not HBM, set the pace -- 2.2 TB/s in round 1).
  float4 nxt[8], nxt2[8];
  load_quarter(blockIdx.x, 0, nxt);
  load_quarter(blockIdx.x, 1, nxt2);
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
    const int tb = it & 1;
    for (int q = 0; q < 4; q++, stage++) {
      const int b = (int)(stage & 1);
      float4 cur[8];
#pragma unroll
      for (int i = 0; i < 8; i++) { cur[i] = nxt[i]; nxt[i] = nxt2[i]; }
      // quarter q + 2 of this tile, or quarter q - 2 of the CTA's next tile (its rows are not written by anyone before we
      // read them: only this CTA writes them, after this read)
      if (q < 2) load_quarter(tile, q + 2, nxt2);
      else load_quarter(tile + gridDim.x, q - 2, nxt2);
      if (stage >= 2) {
        mbar_wait(&S.mma_done[b], ph_buf[b]);
        ph_buf[b] ^= 1;
      }
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int rr = warp * 32 + i * 4 + (lane >> 3), c4 = lane & 7;
        const float4 v = cur[i];
        const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
        const int off = (rr >> 3) * kRtSBOA + c4 * kRtLBOA + (rr & 7) * 16;
        *reinterpret_cast<float4*>(&S.A[b][0][off]) = hi;
        *reinterpret_cast<float4*>(&S.A[b][1][off]) = lo;
      }
      fence_proxy_async();
      __syncthreads();   // also orders the previous drain's TMEM reads before this tile's first (overwriting) MMA
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d0 = tmem + (uint32_t)(tb * 256), d1 = d0 + kTcK;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
          const uint64_t ah = rt_desc(&S.A[b][0][ks * 2 * kRtLBOA], kRtLBOA, kRtSBOA);
          const uint64_t al = rt_desc(&S.A[b][1][ks * 2 * kRtLBOA], kRtLBOA, kRtSBOA);
          const int kofs = (q * 8 + ks * 2) * kRtLBOB;
          const uint64_t bh = rt_desc(&S.B[0][kofs], kRtLBOB, kRtSBOB);
          const uint64_t bl = rt_desc(&S.B[1][kofs], kRtLBOB, kRtSBOB);
          const uint32_t first = (q == 0 && ks == 0) ? 0u : 1u;
          tc_mma_tf32(d0, ah, bh, first);   // hi * hi
          tc_mma_tf32(d1, ah, bl, first);   // hi * lo
          tc_mma_tf32(d1, al, bh, 1u);      // lo * hi
        }
        tc_commit(&S.mma_done[b]);
        if (q == 3) tc_commit(&S.acc_ready[tb]);
      }
    }
    if (prev_tile >= 0) drain((it - 1) & 1, prev_tile);   // overlaps with the MMAs just issued
    prev_tile = tile;
  }
  if (prev_tile >= 0) drain((it - 1) & 1, prev_tile);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace b200als
