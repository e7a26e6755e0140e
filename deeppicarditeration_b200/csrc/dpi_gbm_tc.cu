// 16-bit tcgen05 integral estimator for the fully nonlinear equation `GBMEquationComplexExact` with SDGD Hessian
// sampling: same mathematics as dpi_gbm_simt.cu (picard/data.py:471-527, :497-502, :1226-1325; picard/equations.py:457-466),
// with the hidden layers of the 2-jet propagation on the 5th-generation tensor cores.
//
// One CTA owns one collocation point (two CTAs per SM).  A tile is ONE path row with its v <= 127 sampled directions.
// The jets of all directions form two A operands of 128 rows x 64 hidden units (dz rows and ddz rows, fp16, K-major
// SWIZZLE_128B); row 127 of the dz operand carries the VALUE activations of the path row, so that the same MMA also
// produces the next layer's pre-activations and with them the ELU' / ELU'' factors of that layer.  Per hidden layer:
//     D_dz [128 x 64] = A_dz . W^T ,   D_ddz [128 x 64] = A_ddz . W^T      (8 tcgen05.mma of 128 x 64 x 16, TMEM columns 0..127)
//     row 127: z = D_dz + b  ->  e1 = ELU'(z), e2 = ELU''(z), h = ELU(z)      (published through shared memory)
//     rows i < v:  dh = e1 dz ,  ddh = e2 dz^2 + e1 ddz                        -> next A operands, or u_jj = w_out . ddh
// The first layer (gather of the sampled W1 columns), the source terms of the exact solution, the MC reduction and the
// baseline gather are the SIMT code of dpi_gbm_simt.cu.  Deliberately bulk-synchronous (no warp specialisation): the
// work per tile is small and latency-bound, and two CTAs per SM overlap each other's barriers.
#include <cstdio>

#include "dpi_common.cuh"
#include "dpi_kernels.h"
#include "dpi_tc_common.cuh"

namespace dpi {
namespace {
using namespace tc;

constexpr int kH = 64;
constexpr int kThreads = 256;
constexpr int kRows = 128;      // A-operand rows: directions 0..126, row 127 = value activations
constexpr int kValRow = 127;
constexpr int kLDW1 = kH + 1;
constexpr int kMaxHidden = 4;

struct GbmTcSmem {
  alignas(1024) uint8_t Adz[kRows * 128];           // [row][64 fp16], SWIZZLE_128B
  alignas(1024) uint8_t Addz[kRows * 128];
  alignas(1024) uint8_t Wimg[kMaxHidden][kH * 128];  // hidden weights W_l[n][k] fp16, SWIZZLE_128B (N = 64 rows)
  float Wt1[128 * kLDW1];                            // [j][h] = W1[h][1+j] (fp32: first layer stays exact)
  float bh[kMaxHidden][kH];
  float e1[kH], e2[kH], hval[kH];                    // ELU'(z), ELU''(z), ELU(z) of the current layer's value row
  float c0[kH], wt[kH], wout[kH];
  float red[2][kRows];                               // output-layer partial dot products of the two column halves
  float xs[128], zs[128], diag[128];
  float gw[kMaxGbmNodes][128];
  float part[4][kH];
  int idx[kRows];
  float scal[16];
  uint64_t mbar;
  uint32_t tmem_ptr;
};

__device__ __forceinline__ uint32_t a_off(int row, int k) {  // byte offset of element (row, k) in a [rows x 64] fp16 image
  return (uint32_t)row * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
}

// four block sums behind one pair of barriers (the row pipeline is barrier bound)
__device__ void block_sum4(GbmTcSmem& S, float& a, float& b, float& c, float& e) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a = warp_sum(a), b = warp_sum(b), c = warp_sum(c), e = warp_sum(e);
  __syncthreads();
  if (lane == 0) S.part[0][warp] = a, S.part[1][warp] = b, S.part[2][warp] = c, S.part[3][warp] = e;
  __syncthreads();
  a = b = c = e = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) a += S.part[0][w], b += S.part[1][w], c += S.part[2][w], e += S.part[3][w];
}
__device__ void block_sum3(GbmTcSmem& S, float& a, float& b, float& c) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a = warp_sum(a), b = warp_sum(b), c = warp_sum(c);
  __syncthreads();
  if (lane == 0) S.part[0][warp] = a, S.part[1][warp] = b, S.part[2][warp] = c;
  __syncthreads();
  a = b = c = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) a += S.part[0][w], b += S.part[1][w], c += S.part[2][w];
}

// u_k-independent part of ffi at (s, X = x + a Z)  (equations.py:457-466; same as dpi_gbm_simt.cu)
__device__ float gbm_source(GbmTcSmem& S, const ProblemDev& P, float s, float a, bool use_z) {
  const int tid = threadIdx.x, d = P.nx;
  float arg[kMaxGbmNodes], sn[kMaxGbmNodes], dummy = 0.f;
  {
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
    if (tid < d) {
      const float X = S.xs[tid] + (use_z ? a * S.zs[tid] : 0.f);
      p0 = S.gw[0][tid] * X;
      if (P.gbm_nodes > 1) p1 = S.gw[1][tid] * X;
      if (P.gbm_nodes > 2) p2 = S.gw[2][tid] * X;
      if (P.gbm_nodes > 3) p3 = S.gw[3][tid] * X;
    }
    block_sum4(S, p0, p1, p2, p3);
    arg[0] = p0, arg[1] = p1, arg[2] = p2, arg[3] = p3;
  }
  float ut = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxGbmNodes; ++k) {
    sn[k] = 0.f;
    if (k < P.gbm_nodes) {
      const float ak = arg[k] + S.scal[8 + k] * s;
      sn[k] = sinf(ak);
      ut += P.gbm_v[k] * S.scal[8 + k] * cosf(ak);
    }
  }
  float lap = 0.f, ab = 0.f, z = 0.f;
  if (tid < d) {
    float h = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxGbmNodes; ++k)
      if (k < P.gbm_nodes) h -= P.gbm_v[k] * S.gw[k][tid] * S.gw[k][tid] * sn[k];
    lap = h;
    ab = fabsf(h);
  }
  block_sum3(S, lap, ab, z);
  return -ut - 0.5f * lap - 0.25f * ab;
}

// first layer of the value row: z1 = c0 + s w_t + a (W_x Z)  ->  e1, e2, hval
__device__ void first_layer(GbmTcSmem& S, int d, float s, float a, bool use_z) {
  const int tid = threadIdx.x, col = tid & 63, part = tid >> 6;
  float acc = 0.f;
  if (use_z)
    for (int j = part; j < d; j += 4) acc = fmaf(S.Wt1[j * kLDW1 + col], S.zs[j], acc);
  S.part[part][col] = acc;
  __syncthreads();
  if (tid < kH) {
    const float z = S.c0[tid] + s * S.wt[tid] + a * (S.part[0][tid] + S.part[1][tid] + S.part[2][tid] + S.part[3][tid]);
    const float e = __expf(fminf(z, 0.f));
    const bool pos = z > 0.f;
    S.hval[tid] = pos ? z : e - 1.f;
    S.e1[tid] = pos ? 1.f : e;
    S.e2[tid] = pos ? 0.f : e;
  }
  __syncthreads();
}

// second derivatives u_jj for the directions idx[0..ndir): result in S.red[0][i] (i < ndir)
__device__ void jet_tile_tc(GbmTcSmem& S, uint32_t sbase, uint32_t tmem, const MlpDev& N, int ndir, uint32_t leader,
                            uint32_t& phase) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_hidden = N.n_linear - 2;
  const int row = (warp & 3) * 32 + lane, hf = warp >> 2;  // TMEM lane = operand row; this warp's 32 columns
  // ---- first-layer jets: dh = ELU'(z1) w, ddh = ELU''(z1) w^2 with w = W1[:,1+j]; value row <- h1 ----
  {
    const int j = (row < ndir) ? S.idx[row] : -1;
    const float* w1 = (j >= 0) ? &S.Wt1[j * kLDW1] : nullptr;
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {  // this thread's 32 columns = 4 chunks of 8
      const int k0 = hf * 32 + c8 * 8;
      uint32_t pd[4], pdd[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float d0 = 0.f, d1 = 0.f, dd0 = 0.f, dd1 = 0.f;
        const int k = k0 + 2 * q;
        if (row == kValRow) {
          d0 = S.hval[k], d1 = S.hval[k + 1];
        } else if (w1) {
          const float wa = w1[k], wb = w1[k + 1];
          d0 = S.e1[k] * wa, d1 = S.e1[k + 1] * wb;
          dd0 = S.e2[k] * wa * wa, dd1 = S.e2[k + 1] * wb * wb;
        }
        pd[q] = pack_f16x2(d0, d1);
        pdd[q] = pack_f16x2(dd0, dd1);
      }
      *reinterpret_cast<uint4*>(S.Adz + a_off(row, k0)) = make_uint4(pd[0], pd[1], pd[2], pd[3]);
      *reinterpret_cast<uint4*>(S.Addz + a_off(row, k0)) = make_uint4(pdd[0], pdd[1], pdd[2], pdd[3]);
    }
  }
  const uint32_t mbar = sbase + (uint32_t)offsetof(GbmTcSmem, mbar);
  const uint64_t dA = make_desc_sw128(sbase + (uint32_t)offsetof(GbmTcSmem, Adz));
  const uint64_t dB = make_desc_sw128(sbase + (uint32_t)offsetof(GbmTcSmem, Addz));
  constexpr uint32_t idesc = make_idesc_f16(128, 64);
  for (int l = 0; l < n_hidden; ++l) {
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {  // warp-convergent issue, elected lane (uniform descriptors)
      tc_fence_after();
      const uint64_t dW = make_desc_sw128(sbase + (uint32_t)offsetof(GbmTcSmem, Wimg) + (uint32_t)l * (kH * 128));
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem, 0);  // lane-0 broadcast: marks the TMEM base as warp-uniform
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        umma_bf16_p(tb, dA + (uint64_t)(2 * ks), dW + (uint64_t)(2 * ks), idesc, ks > 0, leader);
        umma_bf16_p(tb + 64u, dB + (uint64_t)(2 * ks), dW + (uint64_t)(2 * ks), idesc, ks > 0, leader);
      }
      tc_commit_p(mbar, leader);
    }
    mbar_wait(mbar, phase, 50 + l);
    phase ^= 1u;
    tc_fence_after();
    // ---- value row: z = D_dz[127] + b -> e1, e2, h of this layer ----
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(hf * 32);
    uint32_t dz[32], ddz[32];
    tmem_ld32(taddr, dz);
    tmem_ld32(taddr + 64u, ddz);
    tmem_ld_wait();
    if (row == kValRow) {
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int k = hf * 32 + c;
        const float z = __uint_as_float(dz[c]) + S.bh[l][k];
        const float e = __expf(fminf(z, 0.f));
        const bool pos = z > 0.f;
        S.hval[k] = pos ? z : e - 1.f;
        S.e1[k] = pos ? 1.f : e;
        S.e2[k] = pos ? 0.f : e;
      }
    }
    tc_fence_before();
    __syncthreads();  // e1 / e2 / hval published; every accumulator read is done, the operands may be overwritten
    const bool last = (l == n_hidden - 1);
    if (!last) {
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        const int k0 = hf * 32 + c8 * 8;
        uint32_t pd[4], pdd[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = c8 * 8 + 2 * q, k = k0 + 2 * q;
          float o0, o1, oo0, oo1;
          if (row == kValRow) {
            o0 = S.hval[k], o1 = S.hval[k + 1], oo0 = oo1 = 0.f;
          } else {
            const float z0 = __uint_as_float(dz[c]), z1 = __uint_as_float(dz[c + 1]);
            o0 = S.e1[k] * z0, o1 = S.e1[k + 1] * z1;
            oo0 = fmaf(S.e2[k] * z0, z0, S.e1[k] * __uint_as_float(ddz[c]));
            oo1 = fmaf(S.e2[k + 1] * z1, z1, S.e1[k + 1] * __uint_as_float(ddz[c + 1]));
          }
          pd[q] = pack_f16x2(o0, o1);
          pdd[q] = pack_f16x2(oo0, oo1);
        }
        *reinterpret_cast<uint4*>(S.Adz + a_off(row, k0)) = make_uint4(pd[0], pd[1], pd[2], pd[3]);
        *reinterpret_cast<uint4*>(S.Addz + a_off(row, k0)) = make_uint4(pdd[0], pdd[1], pdd[2], pdd[3]);
      }
    } else {
      // last hidden layer: u_jj = w_out . ddh over this warp's 32 columns
      float pj = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int k = hf * 32 + c;
        const float z0 = __uint_as_float(dz[c]);
        pj = fmaf(S.wout[k], fmaf(S.e2[k] * z0, z0, S.e1[k] * __uint_as_float(ddz[c])), pj);
      }
      S.red[hf][row] = pj;
      __syncthreads();
      if (tid < kRows) S.red[0][tid] += S.red[1][tid];
      __syncthreads();
    }
  }
}

template <bool DET>
__global__ void __launch_bounds__(kThreads, 2)
gbm_tc_kernel(const __grid_constant__ ProblemDev P, const __grid_constant__ MlpDev N, const float* __restrict__ tx,
              int M_I, uint64_t seed, int64_t point_base, const float* __restrict__ u_s, const float* __restrict__ zS,
              const int32_t* __restrict__ idx_in, float bound, float* __restrict__ y) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  GbmTcSmem& S = *reinterpret_cast<GbmTcSmem*>(smem_raw);
  const uint32_t sbase = smem_u32(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = P.nx, v = P.sdgd_v;
  const int64_t p = blockIdx.x;
  const uint64_t gp = (uint64_t)(point_base + p);
  const float* txp = tx + p * (int64_t)(1 + d);
  const float t = __ldg(txp), dT = horizon(P, t);
  const float fd = (float)d;
  const int n_hidden = N.n_linear - 2;

  if (tid == 0) {
    if (sbase & 1023u) __trap();
    mbar_init(sbase + (uint32_t)offsetof(GbmTcSmem, mbar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(sbase + (uint32_t)offsetof(GbmTcSmem, tmem_ptr), 128);
  // ---- stage the network and the point ----
  for (int i = tid; i < 128; i += kThreads) {
    S.xs[i] = (i < d) ? __ldg(txp + 1 + i) : 0.f;
    S.zs[i] = 0.f;
    for (int k = 0; k < kMaxGbmNodes; ++k) S.gw[k][i] = (k < P.gbm_nodes && i < d) ? P.gbm_w[k * (1 + d) + 1 + i] : 0.f;
  }
  if (tid < kMaxGbmNodes) S.scal[8 + tid] = (tid < P.gbm_nodes) ? P.gbm_w[tid * (1 + d)] : 0.f;
  for (int i = tid; i < 128 * kH; i += kThreads) {
    const int j = i / kH, h = i - j * kH;
    S.Wt1[j * kLDW1 + h] = N.Wt1[i];  // packed [d_pad=128][HP=64]
  }
  for (int l = 0; l < n_hidden; ++l) {
    for (int i = tid; i < kH * kH; i += kThreads) {
      const int k = i / kH, n = i - k * kH;  // Wth[l][k][n] = W_l[n][k]
      *reinterpret_cast<__half*>(S.Wimg[l] + a_off(n, k)) = __float2half_rn(N.Wth[l][i]);
    }
    if (tid < kH) S.bh[l][tid] = N.bh[l][tid];
  }
  if (tid < kH) {
    S.wt[tid] = N.wt[tid];
    S.wout[tid] = N.wout[tid];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&S.tmem_ptr);  // two CTAs per SM share the 512 columns
  {
    const int col = tid & 63, part = tid >> 6;
    float acc = 0.f;
    for (int j = part; j < d; j += 4) acc = fmaf(S.Wt1[j * kLDW1 + col], S.xs[j], acc);
    S.part[part][col] = acc;
    __syncthreads();
    if (tid < kH) S.c0[tid] = N.b1[tid] + S.part[0][tid] + S.part[1][tid] + S.part[2][tid] + S.part[3][tid];
    __syncthreads();
  }
  const uint32_t leader = (warp == 0) ? elect_one() : 0u;
  uint32_t phase = 0;

  // ---- baseline: source(t,x) and the full Hessian diagonal of u_k at (t,x) (data.py:1276-1278) ----
  const float src_b = gbm_source(S, P, t, 0.f, false);
  if (tid < kRows) S.idx[tid] = tid;
  first_layer(S, d, t, 0.f, false);
  jet_tile_tc(S, sbase, tmem, N, d, leader, phase);
  if (tid < 128) S.diag[tid] = (tid < d) ? S.red[0][tid] : 0.f;
  __syncthreads();

  float G = 0.f, fsum = 0.f;
  const float alpha_term = 0.5f * (1.f - P.alpha) * fd, abs_term = 0.25f * fd, inv_v = 1.f / (float)v;
  for (int m = 0; m < M_I; ++m) {
    float uu;
    if (DET)
      uu = __ldg(u_s + p * M_I + m);
    else
      uu = u01(philox_at(seed, gp, kStreamS, (uint32_t)m, 0u).x);
    const float s = uu * dT + t, a = sqrtf(uu * dT) * P.sigma;
    if (warp == 0) {
      const int j0 = 4 * lane;
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 < d) {
        if (DET)
          z = load_row4(zS + ((int64_t)p * M_I + m) * d, j0, d, (d & 3) == 0);
        else
          z = normal4(seed, gp, kStreamIntegral, (uint32_t)m, (uint32_t)lane);
      }
      S.zs[j0] = (j0 < d) ? z.x : 0.f;
      S.zs[j0 + 1] = (j0 + 1 < d) ? z.y : 0.f;
      S.zs[j0 + 2] = (j0 + 2 < d) ? z.z : 0.f;
      S.zs[j0 + 3] = (j0 + 3 < d) ? z.w : 0.f;
    } else if (warp == 1) {
      int jj[4];
      if (DET) {
#pragma unroll
        for (int q = 0; q < 4; ++q)  // caller-supplied coordinates are clamped into [0, d): never index shared memory out of range
          jj[q] = (4 * lane + q < v) ? min(max(__ldg(idx_in + ((int64_t)p * M_I + m) * v + 4 * lane + q), 0), d - 1) : 0;
      } else {
        const uint4 r = philox_at(seed, gp, kStreamIdx, (uint32_t)m, (uint32_t)lane);
        jj[0] = (int)__umulhi(r.x, (uint32_t)d), jj[1] = (int)__umulhi(r.y, (uint32_t)d);
        jj[2] = (int)__umulhi(r.z, (uint32_t)d), jj[3] = (int)__umulhi(r.w, (uint32_t)d);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) S.idx[4 * lane + q] = jj[q];
    }
    __syncthreads();
    const float src = gbm_source(S, P, s, a, true);
    first_layer(S, d, s, a, true);
    jet_tile_tc(S, sbase, tmem, N, v, leader, phase);
    float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f, z0 = 0.f;
    if (tid < v) {
      const float ujj = S.red[0][tid];
      const float dj = S.diag[S.idx[tid]];
      a1 = ujj, a2 = fabsf(ujj), b1 = dj, b2 = fabsf(dj);
    }
    block_sum4(S, a1, a2, b1, b2);
    const float f = alpha_term * (a1 * inv_v) + abs_term * (a2 * inv_v) + src;
    const float fb = alpha_term * (b1 * inv_v) + abs_term * (b2 * inv_v) + src_b;
    fsum += dT * f;
    const float coef = dT * (f - fb) / a;
    if (tid < d) G = fmaf(coef, S.zs[tid], G);
    __syncthreads();
  }

  float* yp = y + p * (int64_t)(1 + d);
  const float invM = 1.f / (float)M_I;
  const bool clip = bound > 0.f && isfinite(bound);
  if (tid < d) {
    float val = yp[1 + tid] + G * invM;
    if (clip) val = fminf(fmaxf(val, -bound), bound);
    yp[1 + tid] = val;
  }
  if (tid == 0) {
    float val = yp[0] + fsum * invM;
    if (clip) val = fminf(fmaxf(val, -bound), bound);
    yp[0] = val;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(*reinterpret_cast<volatile uint32_t*>(&S.tmem_ptr), 128);
  }
}

}  // namespace

int launch_gbm_tc(const ProblemDev& P, const MlpDev& N, const float* tx, int64_t n, int M_I, uint64_t seed, int64_t base,
                  const float* u_s, const float* zS, const int32_t* idx, float bound, float* y, cudaStream_t st) {
  if (n <= 0) return DPI_OK;
  if (P.nx > 127 || P.sdgd_v < 1 || P.sdgd_v > 127 || N.HP != 64 || N.n_linear < 3 || N.n_linear - 2 > kMaxHidden ||
      N.d_pad != 128) {
    set_error("the tensor-core SDGD kernel needs nx <= 127, 1 <= v <= 127 and an MLP with 1..%d hidden GEMM layers of "
              "width <= 64 (got nx=%d v=%d)", kMaxHidden, P.nx, P.sdgd_v);
    return DPI_ERR_UNSUPPORTED;
  }
  if (zS && !idx) {
    set_error("deterministic mode with SDGD needs the idx draws");
    return DPI_ERR_INVALID;
  }
  const size_t smem = sizeof(GbmTcSmem) + 1024;
  auto kd = gbm_tc_kernel<true>;
  auto kp = gbm_tc_kernel<false>;
  DPI_CUDA_CHECK(cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DPI_CUDA_CHECK(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (zS)
    kd<<<(unsigned)n, kThreads, smem, st>>>(P, N, tx, M_I, seed, base, u_s, zS, idx, bound, y);
  else
    kp<<<(unsigned)n, kThreads, smem, st>>>(P, N, tx, M_I, seed, base, u_s, zS, idx, bound, y);
  DPI_CUDA_CHECK(cudaGetLastError());
  return DPI_OK;
}

}  // namespace dpi
