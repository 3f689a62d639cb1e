// gram_tc.cuh -- XtX = X'X on the 5th-generation tensor cores (tcgen05, accumulators in TMEM), rank 128 and 256
// (128 x 128 blocks of the lower triangle), 3xTF32 (fp32-grade, default) or bf16 operands (gram_tc_blocks_kernel).
// Reference: `tcrossprod(X) + lambda*I`, R/model_WRMF.R:474-486 / :347-353 (BLAS syrk/gemm in R) -- the one
// genuinely dense contraction of the half-iteration (2*n*k^2 flop).
//
// fp32-grade result from TF32 tensor cores by the 3xTF32 split: x = hi + lo with hi = x truncated to TF32 (low 13
// mantissa bits cleared), lo = x - hi (exact); G += hi'hi + hi'lo + lo'hi (the dropped lo'lo term is ~2^-22).
// D = A * B with A = X' (M = 128 features, K = rows) and B = X (K = rows, N = 128 features): both operands are the
// SAME shared-memory tile, stored K-major -- i.e. the row block is transposed while it is staged:
//   element (feature m, row k) at  (m/8)*SBO + (k/4)*LBO + (m%8)*16 + (k%4)*4   (canonical no-swizzle K-major
//   layout ((8,n),2):((1,SBO),LBO) in 16-byte units, cute/atom/mma_traits_sm100.hpp)
// One CTA of 128 threads per slice of rows: stage 32 rows (4 x LDG.128 per lane -> 8 x STS.128 hi/lo), one thread
// issues 3 x 4 `tcgen05.mma.cta_group::1.kind::tf32` (M128 N128 K8) per 32 rows, double-buffered tiles released by
// `tcgen05.commit` -> mbarrier.  The tensor core adds into its fp32 accumulator with truncation, a bias of ~3e-8
// per add that grows with the length of the chain (measured: 5.4e-6 relative after 96 adds).  Two measures keep
// it at ~1e-6: the large hi'hi products and the small cross terms use SEPARATE TMEM accumulators (the cross terms
// then truncate at their own, 1000x smaller magnitude), and every 256 rows both accumulators are drained
// (`tcgen05.ld.32x32b`) into per-thread registers with round-to-nearest adds.  Per-CTA partials go out in double
// and are summed in a fixed order by gram_reduce_kernel (bit-reproducible).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace b200als {

constexpr int kTcK = 128;          // rank
constexpr int kTcRows = 32;        // rows staged per tile (K extent of a tile)
constexpr int kTcDrainTiles = 8;   // tiles (256 rows) accumulated in TMEM between drains
constexpr int kTcLBO = 128;                               // bytes between core matrices adjacent in K
constexpr int kTcSBO = (kTcRows / 4) * 128 + 16;          // bytes between 8-feature groups (+16: bank spread)
constexpr int kTcTileBytes = (kTcK / 8) * kTcSBO;         // one hi (or lo) tile


__device__ __forceinline__ uint64_t tc_smem_desc(const void* p) {
  // SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout NONE [61,64)
  return (uint64_t)((smem_u32(p) & 0x3FFFF) >> 4) | ((uint64_t)(kTcLBO >> 4) << 16) | ((uint64_t)(kTcSBO >> 4) << 32) |
         (1ull << 46);
}
// InstrDescriptor: c_format F32 (1<<4) | a_format TF32 (2<<7) | b_format TF32 (2<<10) | K-major A,B | N>>3 [17,23) | M>>4 [24,29)
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(kTcIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// kind::f16 with bf16 operands: a_format = b_format = BF16 (1), fp32 accumulate; one MMA covers K = 16 rows
constexpr uint32_t kTcIdescBf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(kTcIdescBf16), "r"(accumulate)
      : "memory");
}
// bf16 tiles: core matrix = 8 features x 16 bytes = 8 rows of K; a 32-row tile is 4 core matrices deep
constexpr int kTcLBO16 = 128;
constexpr int kTcSBO16 = (kTcRows / 8) * 128 + 16;
constexpr int kTcTileBytes16 = (kTcK / 8) * kTcSBO16;
__device__ __forceinline__ uint64_t tc_smem_desc16(const void* p) {
  return (uint64_t)((smem_u32(p) & 0x3FFFF) >> 4) | ((uint64_t)(kTcLBO16 >> 4) << 16) | ((uint64_t)(kTcSBO16 >> 4) << 32) |
         (1ull << 46);
}
struct GramTcSmem2 {   // two operand slices (rank 256: off-diagonal 128 x 128 blocks of XtX)
  alignas(128) unsigned char tile[2][2][2][kTcTileBytes];   // [buffer][slice a / b][hi / lo]   (bf16 mode: [..][..][0] only)
  uint64_t mma_done[2];
  uint64_t acc_ready;
  uint32_t tmem_base;
};

// XtX in 128 x 128 blocks: blockIdx.y = lower-triangular tile (ta, tb), D = X[:, 128 ta ..]' X[:, 128 tb ..] over this CTA's
// rows; X has `ld` floats per row (128 or 256).  kBf16 = false: 3xTF32 split (fp32-grade, two accumulators, see above);
// kBf16 = true: operands rounded to bf16 (RN), ONE tcgen05.mma.kind::f16 pass, fp32 accumulate -- the "tensor-core bf16
// Gram" of BASELINE configs[4], ~2^-9 relative operand rounding.
// partials: [gridDim.x][n_tiles][128*128] doubles (row-major [a][b]) -- the layout gram_reduce_kernel reads.
template <bool kBf16>
__global__ void __launch_bounds__(128) gram_tc_blocks_kernel(const float* __restrict__ X, int ld, long long n, long long rows_per_cta,
                                                             double* __restrict__ partials) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  GramTcSmem2& S = *reinterpret_cast<GramTcSmem2*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int ta = 0, acc_t = 0;
  while (acc_t + ta + 1 <= (int)blockIdx.y) { acc_t += ta + 1; ta++; }
  const int tb = (int)blockIdx.y - acc_t;
  const bool two = (ta != tb);
  const float* Xa = X + ta * kTcK;
  const float* Xb = X + tb * kTcK;
  const long long r_begin = (long long)blockIdx.x * rows_per_cta;
  const long long r_end = min(n, r_begin + rows_per_cta);
  if (tid == 0) {
    mbar_init(&S.mma_done[0], 1);
    mbar_init(&S.mma_done[1], 1);
    mbar_init(&S.acc_ready, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;

  float acc[kTcK];   // thread (warp w, lane l) owns output row a = 32 w + l of the block, all 128 columns
#pragma unroll
  for (int c = 0; c < kTcK; c++) acc[c] = 0.f;

  const long long n_tiles = (r_end > r_begin) ? (r_end - r_begin + kTcRows - 1) / kTcRows : 0;
  uint32_t phase_buf[2] = {0, 0}, phase_acc = 0;
  for (long long t = 0; t < n_tiles; t++) {
    const int b = (int)(t & 1);
    if (t >= 2) {
      mbar_wait(&S.mma_done[b], phase_buf[b]);
      phase_buf[b] ^= 1;
    }
    const long long r0 = r_begin + t * kTcRows;
    // ---- stage 32 rows of one or two 128-feature slices, transposed to K-major -------------------------------------
#pragma unroll
    for (int sl = 0; sl < 2; sl++) {
      if (sl == 1 && !two) break;
      const float* Xs = sl ? Xb : Xa;
      if constexpr (!kBf16) {
#pragma unroll
        for (int g = 0; g < kTcRows / 16; g++) {
          const int kb = warp + 4 * g;   // k4-block inside the tile
          float4 v[4];
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            const long long r = r0 + kb * 4 + kk;
            v[kk] = (r < r_end) ? __ldg(reinterpret_cast<const float4*>(Xs + (size_t)r * ld) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          const float col[4][4] = {{v[0].x, v[1].x, v[2].x, v[3].x}, {v[0].y, v[1].y, v[2].y, v[3].y},
                                   {v[0].z, v[1].z, v[2].z, v[3].z}, {v[0].w, v[1].w, v[2].w, v[3].w}};
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int m = 4 * lane + j;   // feature
            float4 hi, lo;
            hi.x = __uint_as_float(__float_as_uint(col[j][0]) & 0xFFFFE000u); lo.x = col[j][0] - hi.x;
            hi.y = __uint_as_float(__float_as_uint(col[j][1]) & 0xFFFFE000u); lo.y = col[j][1] - hi.y;
            hi.z = __uint_as_float(__float_as_uint(col[j][2]) & 0xFFFFE000u); lo.z = col[j][2] - hi.z;
            hi.w = __uint_as_float(__float_as_uint(col[j][3]) & 0xFFFFE000u); lo.w = col[j][3] - hi.w;
            const int off = (m >> 3) * kTcSBO + kb * kTcLBO + (m & 7) * 16;
            *reinterpret_cast<float4*>(&S.tile[b][sl][0][off]) = hi;
            *reinterpret_cast<float4*>(&S.tile[b][sl][1][off]) = lo;
          }
        }
      } else {
        // bf16: warp w packs the 8 rows of k8-block w: 8 x LDG.128 per lane -> 4 x STS.128 (8 bf16 each)
        const int kb = warp;
        float4 v[8];
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
          const long long r = r0 + kb * 8 + kk;
          v[kk] = (r < r_end) ? __ldg(reinterpret_cast<const float4*>(Xs + (size_t)r * ld) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int m = 4 * lane + j;
          float e[8];
#pragma unroll
          for (int kk = 0; kk < 8; kk++) e[kk] = (j == 0) ? v[kk].x : (j == 1) ? v[kk].y : (j == 2) ? v[kk].z : v[kk].w;
          uint4 pk;
          pk.x = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(e[0])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(e[1])) << 16);
          pk.y = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(e[2])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(e[3])) << 16);
          pk.z = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(e[4])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(e[5])) << 16);
          pk.w = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(e[6])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(e[7])) << 16);
          const int off = (m >> 3) * kTcSBO16 + kb * kTcLBO16 + (m & 7) * 16;
          *reinterpret_cast<uint4*>(&S.tile[b][sl][0][off]) = pk;
        }
      }
    }
    fence_proxy_async();
    __syncthreads();
    const bool last_of_window = ((t % kTcDrainTiles) == kTcDrainTiles - 1) || (t == n_tiles - 1);
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int sb = two ? 1 : 0;
      if constexpr (!kBf16) {
#pragma unroll
        for (int ks = 0; ks < kTcRows / 8; ks++) {
          const uint64_t ah = tc_smem_desc(&S.tile[b][0][0][ks * 2 * kTcLBO]);
          const uint64_t al = tc_smem_desc(&S.tile[b][0][1][ks * 2 * kTcLBO]);
          const uint64_t bh = tc_smem_desc(&S.tile[b][sb][0][ks * 2 * kTcLBO]);
          const uint64_t bl = tc_smem_desc(&S.tile[b][sb][1][ks * 2 * kTcLBO]);
          const uint32_t first = ((t % kTcDrainTiles) == 0 && ks == 0) ? 0u : 1u;
          tc_mma_tf32(tmem, ah, bh, first);            // hi' hi          -> accumulator 0
          tc_mma_tf32(tmem + kTcK, ah, bl, first);     // hi' lo          -> accumulator 1
          tc_mma_tf32(tmem + kTcK, al, bh, 1u);        // lo' hi          -> accumulator 1
        }
      } else {
#pragma unroll
        for (int ks = 0; ks < kTcRows / 16; ks++) {
          const uint64_t a16 = tc_smem_desc16(&S.tile[b][0][0][ks * 2 * kTcLBO16]);
          const uint64_t b16 = tc_smem_desc16(&S.tile[b][sb][0][ks * 2 * kTcLBO16]);
          const uint32_t first = ((t % kTcDrainTiles) == 0 && ks == 0) ? 0u : 1u;
          tc_mma_bf16(tmem, a16, b16, first);
        }
      }
      tc_commit(&S.mma_done[b]);
      if (last_of_window) tc_commit(&S.acc_ready);
    }
    if (last_of_window) {
      mbar_wait(&S.acc_ready, phase_acc);
      phase_acc ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < (kBf16 ? kTcK : 2 * kTcK); c0 += 32) {   // columns [0,128): main ; [128,256): cross terms (3xTF32)
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 32; c++) acc[(c0 + c) & (kTcK - 1)] += __uint_as_float(r[c]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
    }
  }
  double* out = partials + ((size_t)blockIdx.x * gridDim.y + blockIdx.y) * (kTcK * kTcK) + (size_t)(warp * 32 + lane) * kTcK;
#pragma unroll
  for (int c = 0; c < kTcK; c++) out[c] = (double)acc[c];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

}  // namespace b200als
