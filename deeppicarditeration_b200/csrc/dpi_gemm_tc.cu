// TF32 tcgen05 GEMM with the fused epilogues of dpi_gemm_epi.cuh, for the 512-wide layers of PISGradNet (HJB):
//   C[R,N] = epilogue(A[R,K] . B),  B given as W[N][K] (b_kn = 0: y = x W^T) or W[K][N] (b_kn = 1: c W),
// fp32 operands in HBM as the fp32 pipeline leaves them (nothing is repacked or converted), read by the tensor
// cores as TF32 (10-bit mantissa), fp32 accumulation in TMEM.
//
//   * one 128 x 128 output tile per CTA, two CTAs per SM (the epilogue of one overlaps the main loop of the other);
//     tiles are numbered with the N tiles fastest, so the CTAs that share a row block run together and its A
//     operand comes from L2 for all but the first of them;
//   * operands go global -> shared with cp.async (16-byte copies with the SWIZZLE_128B XOR applied to the
//     destination; 4-byte copies that transpose a [K][N] weight on the fly), three stages of 32 K-columns
//     (one 128-byte swizzle row of fp32) in flight;
//   * one elected thread issues four tcgen05.mma.kind::tf32 (K = 8) per stage and commits the stage's "empty"
//     barrier; after the last stage it commits the accumulator barrier;
//   * all eight warps run the epilogue (TMEM lane quarter x column half, thread = output row), through the same
//     gemm_epilogue4 as the fp32 kernels.
#include <cstdio>

#include "dpi_common.cuh"
#include "dpi_gemm_epi.cuh"
#include "dpi_kernels.h"
#include "dpi_tc_common.cuh"

namespace dpi {
namespace {
using namespace tc;

constexpr int kTile = 128, kStageK = 32, kStages = 3, kThreads = 256;
constexpr uint32_t kOpBytes = kTile * kStageK * 4;                    // 16 KB: 128 rows x 128 B
constexpr uint32_t kSmemBytes = 2 * kStages * kOpBytes + 64 + 1024;   // operands + barriers + alignment slack

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(kThreads, 2) tf32_gemm_kernel(const GemmArgs g, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms are 1024-byte aligned
  uint8_t* const smem = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t sA = sbase, sB = sbase + kStages * kOpBytes;
  const uint32_t bars = sbase + 2 * kStages * kOpBytes;  // empty[kStages], accfull, tmem pointer
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int64_t rt = blockIdx.x / n_tiles;
  const int nt = blockIdx.x - (int)(rt * n_tiles);
  const int64_t r0 = rt * kTile;
  const int n0 = nt * kTile;
  const int K = g.K, N = g.N;
  const int nk = (K + kStageK - 1) / kStageK;  // the last chunk may be partial (K % 4 == 0): zero-filled

  if (t == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bars + 8u * s, 1);
    mbar_init(bars + 8u * kStages, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(bars + 8u * (kStages + 1), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 2 * kStages * kOpBytes + 8 * (kStages + 1));

  // ---- operand loads of one K chunk (all threads) ----
  auto load_chunk = [&](int kc, int st) {
    const int k0 = kc * kStageK;
    const uint32_t a_dst = sA + (uint32_t)st * kOpBytes, b_dst = sB + (uint32_t)st * kOpBytes;
    // A: 128 rows x 8 chunks of 16 B; thread handles chunk c4 = t & 7 of rows (t >> 3) + 32 i
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = (t >> 3) + 32 * i, c4 = t & 7;
      const bool ok = r0 + row < g.R && k0 + 4 * c4 < K;
      const float* src = g.A + (ok ? (r0 + row) * (int64_t)g.lda + k0 + 4 * c4 : 0);
      cp_async16(a_dst + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((c4 ^ (row & 7)) << 4), src,
                 ok ? 16u : 0u);
    }
    if (!g.b_kn) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = (t >> 3) + 32 * i, c4 = t & 7;
        const bool ok = n0 + row < N && k0 + 4 * c4 < K;
        const float* src = g.B + (ok ? (int64_t)(n0 + row) * g.ldb + k0 + 4 * c4 : 0);
        cp_async16(b_dst + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((c4 ^ (row & 7)) << 4), src,
                   ok ? 16u : 0u);
      }
    } else {
      // W[K][N]: element (k, n) goes to row n, K column k of the K-major image; a warp reads 32 consecutive n of one k
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int e = t + kThreads * i;  // 0 .. 4095
        const int k = e >> 7, n = e & 127;
        const bool ok = n0 + n < N && k0 + k < K;
        const float* src = g.B + (ok ? (int64_t)(k0 + k) * g.ldb + n0 + n : 0);
        cp_async4(b_dst + (uint32_t)(n >> 3) * 1024u + (uint32_t)(n & 7) * 128u + (uint32_t)(((k >> 2) ^ (n & 7)) << 4) +
                      (uint32_t)(k & 3) * 4u,
                  src, ok ? 4u : 0u);
      }
    }
  };

  // ---- main loop ----
  const uint32_t leader = elect_one();
  constexpr uint32_t idesc = make_idesc_tf32(kTile, kTile);
  for (int kc = 0; kc < kStages - 1; ++kc) {
    if (kc < nk) load_chunk(kc, kc);
    cp_async_commit();
  }
  for (int kc = 0; kc < nk; ++kc) {
    const int st = kc % kStages;
    // refill the stage that chunk kc-1 used (free once its MMAs have completed) with chunk kc + kStages - 1
    const int kn = kc + kStages - 1;
    if (kn < nk) {
      const int sn = kn % kStages;
      if (kn >= kStages) mbar_wait(bars + 8u * sn, (uint32_t)((kn / kStages - 1) & 1), 60);
      load_chunk(kn, sn);
    }
    cp_async_commit();
    cp_async_wait<kStages - 1>();  // this thread's copies of chunk kc have landed
    fence_proxy_async();           // ... and are visible to the tensor core's (async proxy) reads
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      const uint64_t dA = make_desc_sw128(sA + (uint32_t)st * kOpBytes), dB = make_desc_sw128(sB + (uint32_t)st * kOpBytes);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_tf32_p(tmem, dA + (uint64_t)(2 * ks), dB + (uint64_t)(2 * ks), idesc, (uint32_t)((kc | ks) != 0), leader);
      tc_commit_p(bars + 8u * st, leader);
      if (kc == nk - 1) tc_commit_p(bars + 8u * kStages, leader);
    }
  }
  mbar_wait(bars + 8u * kStages, 0, 61);
  tc_fence_after();

  // ---- epilogue: warp = (lane quarter q, column half h), thread = output row ----
  const int q = warp & 3, h = warp >> 2;
  const int64_t row = r0 + q * 32 + lane;
  const bool c_vec = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
  const bool d_vec = g.D && ((g.ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.D) & 15) == 0);
  const bool e_vec = g.E && ((g.lde & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.E) & 15) == 0);
  const uint32_t tacc = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(64 * h);
#pragma unroll 1
  for (int ch = 0; ch < 4; ++ch) {
    uint32_t v[16];
    tmem_ld16(tacc + (uint32_t)(16 * ch), v);
    tmem_ld_wait();
    if (row < g.R) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const int col = n0 + 64 * h + 16 * ch + j;
        if (col < N) {
          float w[4] = {__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                        __uint_as_float(v[j + 3])};
          gemm_epilogue4(g, row, col, w, c_vec, d_vec, e_vec);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

__global__ void transpose_kernel(const float* __restrict__ W, int rows, int cols, float* __restrict__ WT) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? W[(size_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) WT[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

}  // namespace

int gemm_transpose(const float* W, int rows, int cols, float* WT, cudaStream_t st) {
  transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, st>>>(W, rows, cols, WT);
  DPI_CUDA_CHECK(cudaGetLastError());
  return DPI_OK;
}

bool gemm_tf32_eligible(const GemmArgs& g) {
  return !g.a_km && g.K % 4 == 0 && g.K >= kStageK && (g.lda & 3) == 0 && (g.b_kn || (g.ldb & 3) == 0) &&
         (reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.B) & 15) == 0 && g.R >= kTile &&
         g.N >= 64;
}

int gemm_tf32(const GemmArgs& g, cudaStream_t st, int64_t* launches) {
  const int n_tiles = (g.N + kTile - 1) / kTile;
  const int64_t r_tiles = (g.R + kTile - 1) / kTile;
  if (r_tiles * n_tiles > 0x7fffffffLL) return DPI_ERR_INVALID;
  static bool attr_set = false;
  if (!attr_set) {
    DPI_CUDA_CHECK(cudaFuncSetAttribute(tf32_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set = true;
  }
  tf32_gemm_kernel<<<(unsigned)(r_tiles * n_tiles), kThreads, kSmemBytes, st>>>(g, n_tiles);
  DPI_CUDA_CHECK(cudaGetLastError());
  if (launches) *launches += 1;
  return DPI_OK;
}

}  // namespace dpi
