// fp32 integral estimator (and u_k / grad u_k evaluation) for the HJB equation `OUProcessEquation` with u_k = PISGradNet.
//
// Restates picard/data.py:471-527 (estimator), :350-366 (path points), :1226-1325 (get_f: u_k and its full input
// gradient), picard/solution.py:211-289 (PISGradNet: sin/cos time embedding, t_encoder, smooth_net, nn_module,
// output S <n, x> + (1 - S) g(gamma x)), picard/equations.py:657-666 (ff) and picard/utils.py:852-915 (GMM log-density
// and its gradient).
//
// The reference back-propagates through the whole network to get grad_x u_k.  Here the gradient is the closed form
// 
//     grad_x u = S (n + J_n^T x) + (1 - S) gamma g_x(gamma x),      J_n^T x = one VJP through nn_module, cotangent x,
// S and t_emb depending on the time only.  nn_module is 512 wide (1 MB of fp32 per layer), so unlike the Burgers
// kernels the weights cannot live in shared memory: the rows (point x MC sample, plus one baseline row per point)
// of a chunk of points go through the layers as batched GEMMs - a register-tiled 128x128x16 fp32 GEMM with the
// bias / ELU / ELU' / VJP-mask epilogues fused - with the activations of the chunk in handle-owned scratch.  The
// weights are read in torch layout in both directions (y = x W^T forward, c W backward), nothing is repacked.
#include <algorithm>
#include <cmath>

#include "dpi_common.cuh"
#include "dpi_kernels.h"
#include "dpi_gemm_epi.cuh"

namespace dpi {
namespace {

// ------------------------------------------------------------------------------------------------------------
// C[R,N] = epilogue(A[R,K] . B), B given as W[N][K] (b_kn = 0: y = x W^T) or W[K][N] (b_kn = 1: c W)
// ------------------------------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 16, kGemmThreads = 256;
constexpr int kLds = BM + 4;  // 528-byte rows: 16-byte aligned, the +4 skews the k rows across banks

__device__ __forceinline__ void load8(const float* __restrict__ p, int i0, int limit, bool vec, float (&v)[8]) {
  // elements [i0, i0+8) of a contiguous run of `limit` floats starting at p; zero beyond
  if (vec && i0 + 7 < limit) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p + i0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p + i0 + 4));
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (i0 + i < limit) ? __ldg(p + i0 + i) : 0.f;
  }
}

__global__ void __launch_bounds__(kGemmThreads, 2) sgemm_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][kLds];
  __shared__ __align__(16) float Bs[2][BK][kLds];
  const int t = threadIdx.x;
  const int M_ = (int)g.R;  // output rows (when a_km: columns of the stored A)
  const int64_t r0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int K = g.K, N = g.N;
  const bool a_vec = ((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
  const bool b_vec = ((g.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0);

  // global -> register staging.  A (and B when b_kn == 0): thread owns 8 consecutive k of one row.
  const int lrow = t & 127, lkq = (t >> 7) * 8;
  // B when b_kn == 1: thread owns 8 consecutive n of one k row
  const int bk_k = t >> 4, bk_n = (t & 15) * 8;
  const float* arow = (!g.a_km && r0 + lrow < g.R) ? g.A + (r0 + lrow) * (int64_t)g.lda : nullptr;
  const float* brow = nullptr;
  if (!g.b_kn && n0 + lrow < N) brow = g.B + (int64_t)(n0 + lrow) * g.ldb;

  float ra[8], rb[8];
  auto fetch = [&](int k0) {
    if (g.a_km) {  // A stored [K][M]: thread owns 8 consecutive m of one k row (weight-gradient GEMMs)
      if (k0 + bk_k < K) {
        load8(g.A + (int64_t)(k0 + bk_k) * g.lda, (int)r0 + bk_n, M_, a_vec, ra);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) ra[i] = 0.f;
      }
    } else if (arow) {
      load8(arow, k0 + lkq, K, a_vec, ra);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) ra[i] = 0.f;
    }
    if (!g.b_kn) {
      if (brow) {
        load8(brow, k0 + lkq, K, b_vec, rb);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) rb[i] = 0.f;
      }
    } else {
      if (k0 + bk_k < K) {
        load8(g.B + (int64_t)(k0 + bk_k) * g.ldb, n0 + bk_n, N, b_vec, rb);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) rb[i] = 0.f;
      }
    }
  };
  auto stash = [&](int buf) {
    if (g.a_km) {
      *reinterpret_cast<float4*>(&As[buf][bk_k][bk_n]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
      *reinterpret_cast<float4*>(&As[buf][bk_k][bk_n + 4]) = make_float4(ra[4], ra[5], ra[6], ra[7]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) As[buf][lkq + i][lrow] = ra[i];
    }
    if (!g.b_kn) {
#pragma unroll
      for (int i = 0; i < 8; ++i) Bs[buf][lkq + i][lrow] = rb[i];
    } else {
      *reinterpret_cast<float4*>(&Bs[buf][bk_k][bk_n]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
      *reinterpret_cast<float4*>(&Bs[buf][bk_k][bk_n + 4]) = make_float4(rb[4], rb[5], rb[6], rb[7]);
    }
  };

  const int ty = t >> 4, tx = t & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  fetch(0);
  stash(0);
  __syncthreads();
  const int nk = (K + BK - 1) / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) fetch((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      stash(buf ^ 1);  // the other buffer was last read before the previous barrier
      __syncthreads();
    }
  }

  // ---- epilogue ----
  const bool c_vec = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
  const bool d_vec = g.D && ((g.ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.D) & 15) == 0);
  const bool e_vec = g.E && ((g.lde & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.E) & 15) == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = r0 + ((i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= g.R) continue;
#pragma unroll
    for (int jg = 0; jg < 2; ++jg) {
      const int col = n0 + jg * 64 + tx * 4;
      if (col >= N) continue;
      float v[4] = {acc[i][jg * 4 + 0], acc[i][jg * 4 + 1], acc[i][jg * 4 + 2], acc[i][jg * 4 + 3]};
      gemm_epilogue4(g, row, col, v, c_vec, d_vec, e_vec);
    }
  }
}

// Small problems (the regression step: a few hundred rows, 128-wide layers): 64x64x16 tiles, 4x4 per thread, optional
// split-K over gridDim.z with partials [z][R][N] in g.ws (epilogue kEpiNone only; summed by splitk_reduce_kernel in a
// fixed order, so the result does not depend on scheduling).
// K step 128: these launches are latency bound (a 512 x 128 x 128 problem is 16 CTAs), so a CTA fetches its whole K
// panel in one round trip (32 independent 16-byte loads per thread) instead of one trip per 16 or 32 columns
constexpr int SM_T = 64, SM_K = 128, kLdsS = SM_T + 4;
constexpr size_t kSmallSmemBytes = 2 * 2 * (size_t)SM_K * kLdsS * sizeof(float);  // As[2], Bs[2]
constexpr int SM_KH = SM_K / 16;
__global__ void __launch_bounds__(256) sgemm_small_kernel(const GemmArgs g, int k_per_split) {
  extern __shared__ __align__(16) float sm_small[];
  float (*As)[SM_K][kLdsS] = reinterpret_cast<float (*)[SM_K][kLdsS]>(sm_small);
  float (*Bs)[SM_K][kLdsS] = As + 2;
  const int t = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * SM_T;
  const int n0 = blockIdx.y * SM_T;
  const int N = g.N, M_ = (int)g.R;
  const int kbeg = blockIdx.z * k_per_split, kend = min(g.K, kbeg + k_per_split);
  const bool a_vec = ((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
  const bool b_vec = ((g.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0);
  // row-major operand ([rows][K]): thread owns 4 consecutive k of one row; transposed operand ([K][cols]): 4 consecutive cols
  const int lrow = t >> 2, lk4 = (t & 3) * 4;
  const int tk = t >> 4, tn4 = (t & 15) * 4;
  auto ld4 = [&](const float* p, int i0, int limit, bool vec, float (&v)[4]) {
    if (vec && i0 + 3 < limit) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(p + i0));
      v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = (i0 + i < limit) ? __ldg(p + i0 + i) : 0.f;
    }
  };
  float ra[SM_KH][4], rb[SM_KH][4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < SM_KH; ++q) {
#pragma unroll
      for (int i = 0; i < 4; ++i) ra[q][i] = rb[q][i] = 0.f;
      const int kq = k0 + 16 * q;
      if (g.a_km) {
        if (kq + tk < kend) ld4(g.A + (int64_t)(kq + tk) * g.lda, (int)r0 + tn4, M_, a_vec, ra[q]);
      } else if (r0 + lrow < g.R) {
        ld4(g.A + (r0 + lrow) * (int64_t)g.lda, kq + lk4, kend, a_vec, ra[q]);
      }
      if (g.b_kn) {
        if (kq + tk < kend) ld4(g.B + (int64_t)(kq + tk) * g.ldb, n0 + tn4, N, b_vec, rb[q]);
      } else if (n0 + lrow < N) {
        ld4(g.B + (int64_t)(n0 + lrow) * g.ldb, kq + lk4, kend, b_vec, rb[q]);
      }
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int q = 0; q < SM_KH; ++q) {
      if (g.a_km) {
        *reinterpret_cast<float4*>(&As[buf][16 * q + tk][tn4]) = make_float4(ra[q][0], ra[q][1], ra[q][2], ra[q][3]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) As[buf][16 * q + lk4 + i][lrow] = ra[q][i];
      }
      if (g.b_kn) {
        *reinterpret_cast<float4*>(&Bs[buf][16 * q + tk][tn4]) = make_float4(rb[q][0], rb[q][1], rb[q][2], rb[q][3]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) Bs[buf][16 * q + lk4 + i][lrow] = rb[q][i];
      }
    }
  };
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int nk = (kend - kbeg + SM_K - 1) / SM_K;
  if (nk > 0) {
    fetch(kbeg);
    stash(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) fetch(kbeg + (kt + 1) * SM_K);
#pragma unroll 16
    for (int kk = 0; kk < SM_K; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      stash(buf ^ 1);
      __syncthreads();
    }
  }
  const int col = n0 + tx * 4;
  if (col >= N) return;
  if (gridDim.z > 1) {  // split-K partial
    float* wsz = g.ws + (int64_t)blockIdx.z * g.R * N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t row = r0 + ty * 4 + i;
      if (row >= g.R) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col + j < N) wsz[row * N + col + j] = acc[i][j];
    }
    return;
  }
  const bool c_vec = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
  const bool d_vec = g.D && ((g.ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.D) & 15) == 0);
  const bool e_vec = g.E && ((g.lde & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.E) & 15) == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = r0 + ty * 4 + i;
    if (row >= g.R) continue;
    float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
    gemm_epilogue4(g, row, col, v, c_vec, d_vec, e_vec);
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, int64_t R, int N, float* __restrict__ C,
                                     int ldc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * N) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += ws[(int64_t)z * R * N + i];
  C[(i / N) * ldc + (i % N)] = s;
}

int gemm_impl(const GemmArgs& g, cudaStream_t st, int64_t* launches) {
  if (g.R <= 0 || g.N <= 0) return DPI_OK;
  if (g.tf32 && gemm_tf32_eligible(g)) return gemm_tf32(g, st, launches);
  const int64_t big_tiles = ((g.R + BM - 1) / BM) * ((g.N + BN - 1) / BN);
  if (big_tiles < 96) {
    // too few 128x128 tiles to fill the 148 SMs: small tiles, and split K when even those are few
    const int64_t tiles = ((g.R + SM_T - 1) / SM_T) * ((g.N + SM_T - 1) / SM_T);
    int splits = 1;
    if (g.ws && g.epi == kEpiNone && tiles < 64 && g.K >= 256) {
      splits = (int)std::min<int64_t>({(int64_t)16, (int64_t)g.K / 64, (int64_t)g.ws_splits});
      if (splits < 1) splits = 1;
    }
    const int kps = ((g.K + splits - 1) / splits + SM_K - 1) / SM_K * SM_K;
    splits = (g.K + kps - 1) / kps;
    dim3 grid((unsigned)((g.R + SM_T - 1) / SM_T), (unsigned)((g.N + SM_T - 1) / SM_T), (unsigned)splits);
    static bool attr_set = false;
    if (!attr_set) {
      DPI_CUDA_CHECK(cudaFuncSetAttribute(sgemm_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSmemBytes));
      attr_set = true;
    }
    sgemm_small_kernel<<<grid, 256, kSmallSmemBytes, st>>>(g, kps);
    DPI_CUDA_CHECK(cudaGetLastError());
    if (launches) *launches += 1;
    if (splits > 1) {
      splitk_reduce_kernel<<<(unsigned)((g.R * g.N + 255) / 256), 256, 0, st>>>(g.ws, splits, g.R, g.N, g.C, g.ldc);
      DPI_CUDA_CHECK(cudaGetLastError());
      if (launches) *launches += 1;
    }
    return DPI_OK;
  }
  dim3 grid((unsigned)((g.R + BM - 1) / BM), (unsigned)((g.N + BN - 1) / BN));
  sgemm_kernel<<<grid, kGemmThreads, 0, st>>>(g);
  DPI_CUDA_CHECK(cudaGetLastError());
  if (launches) *launches += 1;
  return DPI_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Row construction.  Row layout of a chunk of `np` points: r = p * rpp + m with rpp = M_I + 1; m < M_I is the MC
// sample m (data.py:350-366), m == M_I the baseline row (t, x) (data.py:511-512).  One extra row at the very end
// (index np * rpp) carries lambda = 0 for smooth_net(emb(0)) (solution.py:247-252).  In `points` mode (net_eval)
// rpp == 1 and every row is (t, x) itself.  One warp per row, lane <-> 4 dims per 128-dim slot.
// ------------------------------------------------------------------------------------------------------------
struct RowBufs {
  float* emb;   // [R+1][128]   sin | cos time embedding
  float* in;    // [R+1][64+d]  nn_module input: t_emb (written by the t_encoder GEMM) | X
  float* z;     // [R][d]       increments (paths mode)
  float* sa;    // [R][2]       (s, a) with a = sqrt(s - t) sigma
};

__global__ void __launch_bounds__(256)
hjb_rows_kernel(const ProblemDev P, const float* __restrict__ tx, int64_t np, int rpp, int M_I, uint64_t seed,
                int64_t point_base, const float* __restrict__ u_s, const float* __restrict__ zS,
                const float* __restrict__ phase, const float* __restrict__ coeff, RowBufs B) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t R = np * rpp;
  if (r > R) return;
  const int d = P.nx, ldin = 64 + d;
  float lam = 0.f;
  if (r < R) {
    const int64_t p = r / rpp;
    const int m = (int)(r - p * rpp);
    const float* txp = tx + p * (int64_t)(1 + d);
    const float t = __ldg(txp);
    const bool path = (rpp > 1) && (m < M_I);
    float s = t, a = 0.f;
    if (path) {
      const uint64_t gp = (uint64_t)(point_base + p);
      const float uu = zS ? __ldg(u_s + p * M_I + m) : u01(philox_at(seed, gp, kStreamS, (uint32_t)m, 0u).x);
      s = uu * horizon(P, t) + t;
      a = sqrtf(uu * horizon(P, t)) * P.sigma;
      const bool vec_ok = (d & 3) == 0;
      for (int j0 = 4 * lane; j0 < d; j0 += 128) {
        float4 z;
        if (zS)
          z = load_row4(zS + (p * M_I + m) * (int64_t)d, j0, d, vec_ok);
        else
          z = normal4(seed, gp, kStreamIntegral, (uint32_t)m, (uint32_t)(j0 >> 2));
        const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (j0 + q < d) {
            B.z[r * d + j0 + q] = zz[q];
            B.in[r * ldin + 64 + j0 + q] = fmaf(a, zz[q], __ldg(txp + 1 + j0 + q));
          }
      }
    } else {
      for (int j = lane; j < d; j += 32) {
        B.in[r * ldin + 64 + j] = __ldg(txp + 1 + j);
        if (rpp > 1) B.z[r * d + j] = 0.f;
      }
    }
    if (lane == 0) {
      B.sa[2 * r] = s;
      B.sa[2 * r + 1] = a;
    }
    lam = P.T - s;
  } else {
    for (int j = lane; j < d; j += 32) B.in[r * ldin + 64 + j] = 0.f;
  }
  // solution.py:211-229: arg up to ~100 rad -> evaluate in double (fast-math sinf is not accurate out there)
  for (int c = lane; c < 64; c += 32) {
    const double arg = (double)__ldg(coeff + c) * (double)lam + (double)__ldg(phase + c);
    double sn, cs;
    sincos(arg, &sn, &cs);
    B.emb[r * 128 + c] = (float)sn;
    B.emb[r * 128 + 64 + c] = (float)cs;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Per-row combination: u, grad_x u, and f = ff(s, X, u, grad) (equations.py:657-666).  One warp per row.
// ------------------------------------------------------------------------------------------------------------
struct FinalArgs {
  const float* in;     // [R+1][64+d]: X at column 64
  const float* nout;   // [R][d]  nn_module output n
  const float* jtx;    // [R][d]  J_n^T x
  const float* sraw;   // [R+1]   smooth_net(emb(lambda))[0]; entry R is the lambda = 0 row
  const float* sa;     // [R][2]
  float* f;            // [R] or null
  float* u;            // [R] or null
  float* du;           // [R][d] or null
  int64_t R;
};

__global__ void __launch_bounds__(256) hjb_final_kernel(const ProblemDev P, const FinalArgs A) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= A.R) return;
  const int d = P.nx, ldin = 64 + d, K = P.ou_K;
  const float S = __ldg(A.sraw + r) - __ldg(A.sraw + A.R);
  const float s = __ldg(A.sa + 2 * r);
  const float gamma = __expf(-0.5f * (P.T - s));
  const float* X = A.in + r * (int64_t)ldin + 64;
  // GMM at y = gamma X (utils.py:852-880): l_k = lognorm_k - 1/2 sum_j (y_j - m_kj)^2 / var_kj
  float q[kMaxOuK];
#pragma unroll
  for (int k = 0; k < kMaxOuK; ++k) q[k] = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float yj = gamma * __ldg(X + j);
#pragma unroll
    for (int k = 0; k < kMaxOuK; ++k)
      if (k < K) {
        const float e = yj - __ldg(P.ou_mean + k * d + j);
        q[k] = fmaf(e * e, __ldg(P.ou_ivar + k * d + j), q[k]);
      }
  }
  float l[kMaxOuK], mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < kMaxOuK; ++k) {
    l[k] = (k < K) ? P.ou_lognorm[k] - 0.5f * warp_sum(q[k]) : -INFINITY;
    mx = fmaxf(mx, l[k]);
  }
  float se = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxOuK; ++k) {
    l[k] = __expf(l[k] - mx);
    se += l[k];
  }
  const float g = -(mx + __logf(se));
  const float inv_se = 1.f / se;
  // grad = S (n + J^T x) + (1 - S) gamma g_x(gamma X),   g_x(y) = sum_k resp_k (y - m_k) / var_k  (utils.py:882-915)
  float dot_nx = 0.f, dot_fg = 0.f, gg = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float xj = __ldg(X + j), yj = gamma * xj;
    float gx = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxOuK; ++k)
      if (k < K) gx = fmaf(l[k] * inv_se, (yj - __ldg(P.ou_mean + k * d + j)) * __ldg(P.ou_ivar + k * d + j), gx);
    const float nj = __ldg(A.nout + r * d + j);
    const float grad = S * (nj + __ldg(A.jtx + r * d + j)) + (1.f - S) * gamma * gx;
    dot_nx = fmaf(nj, xj, dot_nx);
    dot_fg = fmaf(P.ou_theta * (P.ou_mu - xj), grad, dot_fg);
    gg = fmaf(grad, grad, gg);
    if (A.du) A.du[r * d + j] = grad;
  }
  dot_nx = warp_sum(dot_nx);
  dot_fg = warp_sum(dot_fg);
  gg = warp_sum(gg);
  if (lane == 0) {
    if (A.u) A.u[r] = S * dot_nx + (1.f - S) * g;
    if (A.f) A.f[r] = -dot_fg - 0.5f * P.alpha * gg - (float)d * P.ou_theta;
  }
}

// ------------------------------------------------------------------------------------------------------------
// MC reduction (data.py:514-527): y[p] += mean_m dT (f_m - f_b) (1, Z_m / (a_m)) + (dT f_b, 0); clip (data.py:222).
// One CTA per point, thread j owns gradient component j; fixed summation order.
// ------------------------------------------------------------------------------------------------------------
// blockIdx.y slices the samples (kRedSlices slices per point, so that a chunk of a few dozen points still fills the
// GPU); inside a CTA four groups of 128 threads take the slice's samples in tiles of 128: a group first writes the
// tile's 128 coefficients to shared memory (one division per sample instead of one per (sample, component)), then
// thread j of the group adds coef_m Z_mj over the tile.  Slice partials [point][slice][1 + d] are added in a fixed
// order by hjb_reduce_finish_kernel.
constexpr int kRedGroups = 4, kRedSlices = 8;
__global__ void __launch_bounds__(128 * kRedGroups)
hjb_reduce_kernel(const ProblemDev P, const float* __restrict__ tx, int M_I, const float* __restrict__ f,
                  const float* __restrict__ z, const float* __restrict__ sa, float* __restrict__ part) {
  __shared__ float coef_s[kRedGroups][128];
  __shared__ float gpart[kRedGroups][1024 + 1];
  __shared__ float fpart[kRedGroups * 4];
  const int64_t p = blockIdx.x;
  const int slice = blockIdx.y;
  const int d = P.nx, rpp = M_I + 1;
  const int gq = threadIdx.x >> 7, tj = threadIdx.x & 127;
  const float dT = horizon(P, __ldg(tx + p * (int64_t)(1 + d)));
  const int64_t rb = p * rpp;
  const float fb = __ldg(f + rb + M_I);
  float G[8], fs = 0.f;  // components tj + 128 k (d <= 1024)
#pragma unroll
  for (int k = 0; k < 8; ++k) G[k] = 0.f;
  const int n_tiles = (M_I + 127) / 128;
  const int t_per = (n_tiles + kRedSlices - 1) / kRedSlices;
  const int t_beg = slice * t_per, t_end = min(n_tiles, t_beg + t_per);
  for (int tile = t_beg + gq; tile - gq < t_end; tile += kRedGroups) {  // same trip count for every group
    const int m0 = tile * 128;
    const int m = m0 + tj;
    float cf = 0.f;
    if (tile < t_end && m < M_I) {
      const float fm = __ldg(f + rb + m);
      cf = dT * (fm - fb) / __ldg(sa + 2 * (rb + m) + 1);
      fs += fm;
    }
    __syncthreads();  // the previous tile's coefficients have been consumed
    coef_s[gq][tj] = cf;
    __syncthreads();
    const int cnt = (tile < t_end) ? min(128, M_I - m0) : 0;
    for (int i = 0; i < cnt; ++i) {
      const float c = coef_s[gq][i];
      const float* zr = z + (rb + m0 + i) * d;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int j = tj + 128 * k;
        if (j < d) G[k] = fmaf(c, __ldg(zr + j), G[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) gpart[gq][tj + 128 * k] = G[k];
  fs = warp_sum(fs);
  if ((threadIdx.x & 31) == 0) fpart[threadIdx.x >> 5] = fs;
  __syncthreads();
  float* out = part + (p * kRedSlices + slice) * (int64_t)(1 + d);
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    float Gs = 0.f;
    for (int q = 0; q < kRedGroups; ++q) Gs += gpart[q][j];
    out[1 + j] = Gs;
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < kRedGroups * 4; ++i) s += fpart[i];
    out[0] = s;
  }
}
__global__ void __launch_bounds__(128)
hjb_reduce_finish_kernel(const ProblemDev P, const float* __restrict__ tx, int M_I, const float* __restrict__ part,
                         float bound, float* __restrict__ y) {
  const int64_t p = blockIdx.x;
  const int d = P.nx;
  const float dT = horizon(P, __ldg(tx + p * (int64_t)(1 + d)));
  const float invM = 1.f / (float)M_I;
  const bool clip = bound > 0.f && isfinite(bound);
  float* yp = y + p * (int64_t)(1 + d);
  for (int j = threadIdx.x; j < 1 + d; j += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < kRedSlices; ++q) s += part[(p * kRedSlices + q) * (int64_t)(1 + d) + j];
    float v = yp[j] + (j == 0 ? dT * s : s) * invM;
    if (clip) v = fminf(fmaxf(v, -bound), bound);
    yp[j] = v;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Chunk scratch carve-up
// ------------------------------------------------------------------------------------------------------------
struct Scratch {
  RowBufs rows;
  float *t1, *sA, *sB, *sraw, *hA, *hB, *E[DPI_MAX_LINEAR], *nout, *jtx, *f;
};

int64_t floats_per_row(const PisDev& N, int d) {
  const int nh = N.n_main - 1;
  return 128 + (64 + d) + d + 2 + 64 + 64 + 64 + 1 + 2 * (int64_t)N.hmax + (int64_t)nh * N.hmax + d + d + 1 + 8;
}

Scratch carve(float* base, int64_t rows1 /* rows incl. the lambda = 0 row */, const PisDev& N, int d) {
  Scratch S{};
  float* p = base;
  auto take = [&](int64_t per_row) {
    float* q = p;
    p += ((rows1 * per_row + 3) / 4) * 4;  // keep every buffer 16-byte aligned
    return q;
  };
  S.rows.emb = take(128);
  S.rows.in = take(64 + d);
  S.rows.z = take(d);
  S.rows.sa = take(2);
  S.t1 = take(64);
  S.sA = take(64);
  S.sB = take(64);
  S.sraw = take(1);
  S.hA = take(N.hmax);
  S.hB = take(N.hmax);
  for (int l = 0; l < N.n_main - 1; ++l) S.E[l] = take(N.hmax);
  S.nout = take(d);
  S.jtx = take(d);
  S.f = take(1);
  return S;
}

// time nets + nn_module forward + VJP for `R` rows (+1 lambda = 0 row) whose emb / X are already in S.rows
int eval_rows(const PisDev& N, int d, const Scratch& S, int64_t R, cudaStream_t st, int64_t* launches) {
  int rc;
  const int64_t R1 = R + 1;
  auto G = [&](const float* A, int lda, const dpi_linear& L, int N_, float* C, int ldc, int epi, float* D = nullptr,
               int ldd = 0) {
    GemmArgs g{};
    g.A = A, g.lda = lda, g.B = L.W, g.ldb = L.n_in, g.b_kn = 0, g.bias = L.b, g.C = C, g.ldc = ldc, g.D = D, g.ldd = ldd;
    g.R = R1, g.N = N_, g.K = L.n_in, g.epi = epi, g.tf32 = N.tf32;
    return gemm_impl(g, st, launches);
  };
  // t_encoder (solution.py:170-174): Linear(128,64) ELU Linear(64,64) -> columns [0,64) of the nn_module input
  if ((rc = G(S.rows.emb, 128, N.t_encoder[0], 64, S.t1, 64, kEpiBiasElu))) return rc;
  if ((rc = G(S.t1, 64, N.t_encoder[1], 64, S.rows.in, 64 + d, kEpiBias))) return rc;
  // smooth_net (solution.py:180-190): only output 0 of the last Linear is used (solution.py:254)
  {
    const float* cur = S.rows.emb;
    int ld = 128;
    float* bufs[2] = {S.sA, S.sB};
    for (int l = 0; l < N.n_smooth - 1; ++l) {
      float* out = bufs[l & 1];
      if ((rc = G(cur, ld, N.smooth_net[l], N.smooth_net[l].n_out, out, 64, kEpiBiasElu))) return rc;
      cur = out, ld = 64;
    }
    if ((rc = G(cur, ld, N.smooth_net[N.n_smooth - 1], 1, S.sraw, 1, kEpiBias))) return rc;
  }
  // nn_module forward (solution.py:195-203), keeping ELU' of every hidden layer for the VJP
  const int L = N.n_main;
  {
    // every hidden activation h_l = ELU(z_l) is kept in S.E[l]: the VJP recovers ELU'(z_l) from it (h > 0 ? 1 : h + 1),
    // so a layer writes one tensor, not two
    const float* cur = S.rows.in;
    int ld = 64 + d;
    for (int l = 0; l < L - 1; ++l) {
      const int w = N.main[l].n_out;
      if ((rc = G(cur, ld, N.main[l], w, S.E[l], w, kEpiBiasElu))) return rc;
      cur = S.E[l], ld = w;
    }
    if ((rc = G(cur, ld, N.main[L - 1], d, S.nout, d, kEpiBias))) return rc;
  }
  // VJP with cotangent x: c = x W_last; c = (c . ELU'(z_l)) W_l for l = L-2 .. 0; keep the x columns of the result
  {
    // Bt (optional): the same weight block transposed, [N_][ldbt] K-major -- used by the tensor-core path
    auto GB = [&](const float* A, int lda, int K_, const float* Bp, int ldb, int N_, const float* E, int lde, float* C,
                  int ldc, const float* Bt = nullptr, int ldbt = 0) {
      GemmArgs g{};
      g.A = A, g.lda = lda, g.B = Bp, g.ldb = ldb, g.b_kn = 1, g.E = E, g.lde = lde, g.C = C, g.ldc = ldc;
      g.R = R, g.N = N_, g.K = K_, g.epi = E ? kEpiMulDElu : kEpiNone, g.tf32 = N.tf32;
      if (N.tf32 && Bt) g.B = Bt, g.ldb = ldbt, g.b_kn = 0;
      return gemm(g, st, launches);
    };
    float* bufs[2] = {S.hA, S.hB};
    const int wl = N.main[L - 1].n_in;
    // main_t[l] = W_l^T is [n_in][n_out]: row n (an input of layer l) holds the K = n_out weights the VJP contracts
    if ((rc = GB(S.rows.in + 64, 64 + d, d, N.main[L - 1].W, wl, wl, S.E[L - 2], wl, bufs[0], wl, N.main_t[L - 1], d)))
      return rc;
    const float* cur = bufs[0];
    int ld = wl, b = 1;
    for (int l = L - 2; l >= 1; --l) {
      const int win = N.main[l].n_in, wout = N.main[l].n_out;
      if ((rc = GB(cur, ld, wout, N.main[l].W, win, win, S.E[l - 1], win, bufs[b], win, N.main_t[l], wout))) return rc;
      cur = bufs[b], ld = win, b ^= 1;
    }
    const int w0 = N.main[0].n_out;
    if ((rc = GB(cur, ld, w0, N.main[0].W + 64, 64 + d, d, nullptr, 0, S.jtx, d,
                 N.main_t[0] ? N.main_t[0] + (size_t)64 * w0 : nullptr, w0)))
      return rc;
  }
  return DPI_OK;
}

}  // namespace

int gemm(const GemmArgs& g, cudaStream_t st, int64_t* launches) { return gemm_impl(g, st, launches); }

// rows of activations (+ the lambda = 0 row, + alignment slack of carve) and, behind them, the slice partials of the MC
// reduction: kRedSlices x (1 + d) floats per point.  (The partials used to alias the hA activation buffer, which is
// too small for narrow networks with few samples: 8 (1 + d) > (M_I + 1) hmax.)
int64_t hjb_scratch_floats(const PisDev& N, int d, int64_t rows, int64_t points) {
  return (rows + 1) * floats_per_row(N, d) + 64 + points * (int64_t)kRedSlices * (1 + d);
}

int64_t hjb_points_per_chunk(const PisDev& N, int d, int M_I, int64_t budget_bytes) {
  const int64_t per_point =
      ((int64_t)(M_I + 1) * floats_per_row(N, d) + (int64_t)kRedSlices * (1 + d)) * (int64_t)sizeof(float);
  int64_t pc = budget_bytes / per_point;
  return pc < 1 ? 1 : pc;
}

int launch_hjb_paths(const ProblemDev& P, const PisDev& N, const float* tx, int64_t n, int M_I, uint64_t seed,
                     int64_t base, const float* u_s, const float* zS, float bound, float* scratch,
                     int64_t points_per_chunk, float* y, cudaStream_t st, int64_t* launches) {
  const int d = P.nx, rpp = M_I + 1;
  for (int64_t p0 = 0; p0 < n; p0 += points_per_chunk) {
    const int64_t np = std::min<int64_t>(points_per_chunk, n - p0);
    const int64_t R = np * rpp;
    const Scratch S = carve(scratch, R + 1, N, d);
    hjb_rows_kernel<<<(unsigned)((R + 1 + 7) / 8), 256, 0, st>>>(
        P, tx + p0 * (1 + d), np, rpp, M_I, seed, base + p0, u_s ? u_s + p0 * M_I : nullptr,
        zS ? zS + p0 * (int64_t)M_I * d : nullptr, N.phase, N.coeff, S.rows);
    DPI_CUDA_CHECK(cudaGetLastError());
    *launches += 1;
    int rc = eval_rows(N, d, S, R, st, launches);
    if (rc) return rc;
    FinalArgs F{};
    F.in = S.rows.in, F.nout = S.nout, F.jtx = S.jtx, F.sraw = S.sraw, F.sa = S.rows.sa, F.f = S.f, F.R = R;
    hjb_final_kernel<<<(unsigned)((R + 7) / 8), 256, 0, st>>>(P, F);
    DPI_CUDA_CHECK(cudaGetLastError());
    // slice partials: np x kRedSlices x (1 + d) floats behind the carved activation buffers of this chunk
    float* part = scratch + hjb_scratch_floats(N, d, R, 0);
    hjb_reduce_kernel<<<dim3((unsigned)np, kRedSlices), 128 * kRedGroups, 0, st>>>(P, tx + p0 * (1 + d), M_I, S.f, S.rows.z,
                                                                                S.rows.sa, part);
    hjb_reduce_finish_kernel<<<(unsigned)np, 128, 0, st>>>(P, tx + p0 * (1 + d), M_I, part, bound, y + p0 * (1 + d));
    DPI_CUDA_CHECK(cudaGetLastError());
    *launches += 3;
  }
  return DPI_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Plain MLP u_k: value and full input gradient at the rows of tx (picard/data.py:1230,1245-1251 as used by the
// evaluation callback, picard/utils.py:362-458): forward chain keeping ELU', then one VJP with cotangent 1.
// scratch: ones [R] | act A [R][hmax] | act B [R][hmax] | E_l [R][hmax] per hidden layer.
// ------------------------------------------------------------------------------------------------------------
__global__ void fill_ones_kernel(float* __restrict__ p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 1.f;
}

int64_t mlp_eval_scratch_floats(const dpi_linear* L, int n_lin, int64_t rows) {
  int hmax = 1;
  for (int l = 0; l < n_lin - 1; ++l) hmax = std::max(hmax, L[l].n_out);
  return rows * (int64_t)(1 + (2 + (n_lin - 1)) * hmax) + 64;
}

int launch_mlp_eval(const dpi_linear* L, int n_lin, int d, const float* tx, int64_t n, float* scratch,
                    int64_t rows_per_chunk, float* u, float* du, cudaStream_t st, int64_t* launches) {
  int hmax = 1;
  for (int l = 0; l < n_lin - 1; ++l) hmax = std::max(hmax, L[l].n_out);
  for (int64_t p0 = 0; p0 < n; p0 += rows_per_chunk) {
    const int64_t R = std::min<int64_t>(rows_per_chunk, n - p0);
    float* ones = scratch;
    float* bufs[2] = {scratch + ((R + 3) / 4) * 4, scratch + ((R + 3) / 4) * 4 + R * (int64_t)hmax};
    float* E0 = bufs[1] + R * (int64_t)hmax;
    auto Eof = [&](int l) { return E0 + (int64_t)l * R * hmax; };
    int rc;
    // forward
    const float* cur = tx + p0 * (1 + d);
    int ld = 1 + d;
    for (int l = 0; l < n_lin - 1; ++l) {
      GemmArgs g{};
      g.A = cur, g.lda = ld, g.B = L[l].W, g.ldb = L[l].n_in, g.b_kn = 0, g.bias = L[l].b;
      g.C = bufs[l & 1], g.ldc = L[l].n_out, g.D = Eof(l), g.ldd = L[l].n_out;
      g.R = R, g.N = L[l].n_out, g.K = L[l].n_in, g.epi = kEpiBiasEluD;
      if ((rc = gemm(g, st, launches))) return rc;
      cur = bufs[l & 1], ld = L[l].n_out;
    }
    if (u) {
      GemmArgs g{};
      g.A = cur, g.lda = ld, g.B = L[n_lin - 1].W, g.ldb = L[n_lin - 1].n_in, g.b_kn = 0, g.bias = L[n_lin - 1].b;
      g.C = u + p0, g.ldc = 1, g.R = R, g.N = 1, g.K = L[n_lin - 1].n_in, g.epi = kEpiBias;
      if ((rc = gemm(g, st, launches))) return rc;
    }
    if (du) {
      fill_ones_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(ones, R);
      DPI_CUDA_CHECK(cudaGetLastError());
      *launches += 1;
      // c = 1 . W_last (K = 1), masked by ELU' of the last hidden layer; then back through the hidden layers
      const int wl = L[n_lin - 1].n_in;
      GemmArgs g{};
      g.A = ones, g.lda = 1, g.B = L[n_lin - 1].W, g.ldb = wl, g.b_kn = 1, g.E = Eof(n_lin - 2), g.lde = wl;
      g.C = bufs[0], g.ldc = wl, g.R = R, g.N = wl, g.K = 1, g.epi = kEpiMulE;
      if ((rc = gemm(g, st, launches))) return rc;
      const float* c = bufs[0];
      int cld = wl, b = 1;
      for (int l = n_lin - 2; l >= 1; --l) {
        GemmArgs h{};
        h.A = c, h.lda = cld, h.B = L[l].W, h.ldb = L[l].n_in, h.b_kn = 1, h.E = Eof(l - 1), h.lde = L[l].n_in;
        h.C = bufs[b], h.ldc = L[l].n_in, h.R = R, h.N = L[l].n_in, h.K = L[l].n_out, h.epi = kEpiMulE;
        if ((rc = gemm(h, st, launches))) return rc;
        c = bufs[b], cld = L[l].n_in, b ^= 1;
      }
      GemmArgs h{};
      h.A = c, h.lda = cld, h.B = L[0].W + 1, h.ldb = 1 + d, h.b_kn = 1;
      h.C = du + p0 * d, h.ldc = d, h.R = R, h.N = d, h.K = L[0].n_out, h.epi = kEpiNone;
      if ((rc = gemm(h, st, launches))) return rc;
    }
  }
  return DPI_OK;
}

int launch_hjb_eval(const ProblemDev& P, const PisDev& N, const float* tx, int64_t n, float* scratch,
                    int64_t rows_per_chunk, float* u, float* du, cudaStream_t st, int64_t* launches) {
  const int d = P.nx;
  for (int64_t p0 = 0; p0 < n; p0 += rows_per_chunk) {
    const int64_t R = std::min<int64_t>(rows_per_chunk, n - p0);
    const Scratch S = carve(scratch, R + 1, N, d);
    hjb_rows_kernel<<<(unsigned)((R + 1 + 7) / 8), 256, 0, st>>>(P, tx + p0 * (1 + d), R, 1, 0, 0, 0, nullptr, nullptr,
                                                               N.phase, N.coeff, S.rows);
    DPI_CUDA_CHECK(cudaGetLastError());
    *launches += 1;
    int rc = eval_rows(N, d, S, R, st, launches);
    if (rc) return rc;
    FinalArgs F{};
    F.in = S.rows.in, F.nout = S.nout, F.jtx = S.jtx, F.sraw = S.sraw, F.sa = S.rows.sa, F.R = R;
    F.u = u ? u + p0 : nullptr;
    F.du = du ? du + p0 * d : nullptr;
    hjb_final_kernel<<<(unsigned)((R + 7) / 8), 256, 0, st>>>(P, F);
    DPI_CUDA_CHECK(cudaGetLastError());
    *launches += 1;
  }
  return DPI_OK;
}

}  // namespace dpi
