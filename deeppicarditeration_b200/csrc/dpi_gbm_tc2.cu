// Warp-specialised tcgen05 integral estimator for the fully nonlinear equation `GBMEquationComplexExact` with SDGD
// Hessian sampling (BASELINE config 4).  Same mathematics as dpi_gbm_simt.cu / dpi_gbm_tc.cu (picard/data.py:471-527,
// :497-502, :1226-1325; picard/equations.py:457-466): per path row, the second derivatives of u_k along v sampled
// coordinate directions are propagated forward as 2-jets (dz, ddz), the hidden layers of that propagation being
// [128 directions x 64] x [64 x 64] GEMMs on the tensor cores.
//
// What changed against dpi_gbm_tc.cu (one path row at a time, ~20 block barriers per row, 34 k points/s):
//   * the VALUE chain of a path row (which fixes ELU'/ELU'' of every layer) no longer rides in the MMA as operand row
//     127: a PREP team (4 warps, two rows in parallel, 64 threads each) computes it in fp32 on the CUDA cores one or
//     two rows ahead, together with the row's draws and the source terms of the exact solution, and publishes a row
//     context (increments, sampled coordinates, ELU' / ELU'' of every layer as fp16x2, s, a, source) through a ring;
//   * the JET team (8 warps; thread = direction row x column half) keeps TWO path rows in flight: while the tensor
//     pipe runs a layer of one row, the team converts the accumulators of the other (packed fp16x2: dh = e1 dz,
//     ddh = e2 dz^2 + e1 ddz -- 7 instructions per column pair instead of ~25 in fp32) into the next A operands;
//   * one warp issues the MMAs; nobody executes a block-wide barrier in the steady state.
// One CTA per collocation point, rows in a fixed order: labels do not depend on the launch geometry.
#include <cstdio>

#include "dpi_common.cuh"
#include "dpi_kernels.h"
#include "dpi_tc_common.cuh"

namespace dpi {
namespace {
using namespace tc;

constexpr int kH = 64, kMaxHid = 3, kNCtx = 4;
constexpr int kJetWarps = 8, kPrepWarp0 = 8, kPrepWarps = 4, kMmaWarp = 12, kThreads = 32 * 13;
constexpr uint32_t kOpBytes = 128 * 128;  // one [128 rows x 64 k] fp16 operand image

struct alignas(16) RowCtx {
  float zs[128];
  int idx[128];
  __half e1[kMaxHid + 1][kH];  // ELU'(z_l) of the row's value chain, l = first layer .. last hidden layer
  __half e2[kMaxHid + 1][kH];  // ELU''(z_l)
  float s, a, src, pad;
};

struct Smem {
  alignas(1024) uint8_t A[2][2][kOpBytes];        // [row slot][dz | ddz]
  alignas(1024) uint8_t Wimg[kMaxHid][kH * 128];  // hidden weights W_l[n][k] fp16, SWIZZLE_128B
  // W1[k][1+j] as [j][k]: the first-layer jet is a row lookup.  Rows are 144 bytes apart (36 words): the 16-byte chunk a
  // warp's 32 threads fetch from 32 different rows then starts in 8 different bank groups instead of one
  alignas(16) __half W1h[128][kH + 8];
  float W1t[128][kH];                             // the same in fp32 (value chain)
  float Wht[kMaxHid][kH][kH];                     // [k'][k] = W_l[k][k'] fp32 (value chain)
  float bh[kMaxHid][kH];
  float c0[kH], wt[kH], wout[kH];
  uint32_t wout_h[kH / 2];
  RowCtx ctx[kNCtx];
  alignas(16) float hbuf[2][2][kH];               // prep: activation vector, [half][ping-pong]
  float xs[128], diag[128];
  float gw[kMaxGbmNodes][128], gw2[kMaxGbmNodes];  // GBM w_k[1+j]; sum_j w_kj^2
  float w0[kMaxGbmNodes];
  float red[2][2][128];                           // [slot][column half][direction]: output-layer partial dots
  float sums[2][4][4];                            // [slot][warp][a1, a2, b1, b2]
  uint64_t bar_ctx_full[kNCtx], bar_ctx_empty[kNCtx], bar_a_ready[2], bar_acc[2];
  uint32_t tmem_ptr;
};

__device__ __forceinline__ uint32_t a_off(int row, int k) {  // byte offset of element (row, k) in a [rows x 64] fp16 image
  return (uint32_t)row * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
}
__device__ __forceinline__ uint32_t h2_of(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

template <bool DET>
__global__ void __launch_bounds__(kThreads, 1)
gbm_tc2_kernel(const __grid_constant__ ProblemDev P, const __grid_constant__ MlpDev N, const float* __restrict__ tx,
               int M_I, uint64_t seed, int64_t point_base, const float* __restrict__ u_s, const float* __restrict__ zS,
               const int32_t* __restrict__ idx_in, float bound, float* __restrict__ y) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const uint32_t sbase = smem_u32(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = P.nx, v = P.sdgd_v;
  const int64_t p = blockIdx.x;
  const uint64_t gp = (uint64_t)(point_base + p);
  const float* txp = tx + p * (int64_t)(1 + d);
  const float t = __ldg(txp), dT = horizon(P, t);
  const int n_hid = N.n_linear - 2;  // hidden GEMM layers
  const int R = M_I + 1;             // row 0 = the baseline pseudo-row (t, x) with all d directions; row 1 + m = sample m
  auto bar = [&](const uint64_t* b) -> uint32_t { return smem_u32(b); };

  if (tid == 0) {
    if (sbase & 1023u) __trap();
    for (int i = 0; i < kNCtx; ++i) {
      mbar_init(bar(&S.bar_ctx_full[i]), 2);            // the two warps of a prep half
      mbar_init(bar(&S.bar_ctx_empty[i]), kJetWarps);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(&S.bar_a_ready[i]), kJetWarps);
      mbar_init(bar(&S.bar_acc[i]), 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(&S.tmem_ptr), 256);
  // ---- stage the network and the point ----
  for (int i = tid; i < 128; i += kThreads) {
    S.xs[i] = (i < d) ? __ldg(txp + 1 + i) : 0.f;
    S.diag[i] = 0.f;
    for (int k = 0; k < kMaxGbmNodes; ++k) S.gw[k][i] = (k < P.gbm_nodes && i < d) ? P.gbm_w[k * (1 + d) + 1 + i] : 0.f;
  }
  if (tid < kMaxGbmNodes) S.w0[tid] = (tid < P.gbm_nodes) ? P.gbm_w[tid * (1 + d)] : 0.f;
  for (int i = tid; i < 128 * kH; i += kThreads) {
    const float w = N.Wt1[i];  // packed [d_pad = 128][HP = 64]
    S.W1t[i / kH][i % kH] = w;
    S.W1h[i / kH][i % kH] = __float2half_rn(w);
  }
  for (int l = 0; l < n_hid; ++l) {
    for (int i = tid; i < kH * kH; i += kThreads) {
      const int k = i / kH, n = i - k * kH;  // Wth[l][k][n] = W_l[n][k]
      const float w = N.Wth[l][i];
      S.Wht[l][k][n] = w;
      *reinterpret_cast<__half*>(S.Wimg[l] + a_off(n, k)) = __float2half_rn(w);
    }
    if (tid < kH) S.bh[l][tid] = N.bh[l][tid];
  }
  if (tid < kH) {
    S.wt[tid] = N.wt[tid];
    S.wout[tid] = N.wout[tid];
  }
  if (tid < kH / 2) S.wout_h[tid] = pack_f16x2(N.wout[2 * tid], N.wout[2 * tid + 1]);
  // operand rows of padded directions must hold finite values: zero both slots once
  for (int i = tid; i < (int)(4 * kOpBytes / 16); i += kThreads) reinterpret_cast<uint4*>(&S.A[0][0][0])[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  {
    // c0 = b1 + W_x x; sum_j w_kj^2
    if (tid < kH) {
      float acc = N.b1[tid];
      for (int j = 0; j < d; ++j) acc = fmaf(S.W1t[j][tid], S.xs[j], acc);
      S.c0[tid] = acc;
    } else if (tid < kH + kMaxGbmNodes) {
      const int k = tid - kH;
      float acc = 0.f;
      for (int j = 0; j < d; ++j) acc = fmaf(S.gw[k][j], S.gw[k][j], acc);
      S.gw2[k] = acc;
    }
  }
  __syncthreads();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&S.tmem_ptr);

  if (warp == kMmaWarp) {
    // =================================== MMA issuer ===================================
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc = make_idesc_f16(128, 64);
    uint32_t ph[2] = {0u, 0u};
    for (int r0 = 0; r0 < R; r0 += 2) {
      const int nrow = min(2, R - r0);
      for (int l = 0; l < n_hid; ++l)
        for (int sl = 0; sl < nrow; ++sl) {
          mbar_wait(bar(&S.bar_a_ready[sl]), ph[sl], 1);
          ph[sl] ^= 1u;
          tc_fence_after();
          const uint64_t dA = make_desc_sw128(smem_u32(&S.A[sl][0][0])), dB = make_desc_sw128(smem_u32(&S.A[sl][1][0]));
          const uint64_t dW = make_desc_sw128(smem_u32(&S.Wimg[l][0]));
          const uint32_t acc = tmem + (uint32_t)sl * 128u;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma_bf16_p(acc, dA + (uint64_t)(2 * ks), dW + (uint64_t)(2 * ks), idesc, ks > 0, leader);
            umma_bf16_p(acc + 64u, dB + (uint64_t)(2 * ks), dW + (uint64_t)(2 * ks), idesc, ks > 0, leader);
          }
          tc_commit_p(bar(&S.bar_acc[sl]), leader);
        }
    }
    __syncwarp();
  } else if (warp >= kPrepWarp0) {
    // =================================== PREP team ===================================
    // half h (64 threads) prepares rows r = h, h + 2, ... into context slot r & 3: draws, source term, value chain
    const int ht = tid - 32 * kPrepWarp0, h = ht >> 6, k = ht & 63, hl = ht & 31, hw = (ht >> 5) & 1;
    uint32_t ph_empty = 0u, used = 0u;
    for (int r = h; r < R; r += 2) {
      const int cs = r & 3;
      RowCtx& C = S.ctx[cs];
      if (used & (1u << cs)) {
        mbar_wait(bar(&S.bar_ctx_empty[cs]), (ph_empty >> cs) & 1u, 2);
        ph_empty ^= 1u << cs;
      }
      used |= 1u << cs;
      const int m = r - 1;
      float uu = 0.f;
      if (r > 0) uu = DET ? __ldg(u_s + p * M_I + m) : u01(philox_at(seed, gp, kStreamS, (uint32_t)m, 0u).x);
      const float s = uu * dT + t, a = sqrtf(uu * dT) * P.sigma;
      // ---- draws ----
      if (r == 0) {
        C.zs[k] = 0.f, C.zs[64 + k] = 0.f;
        C.idx[k] = k, C.idx[64 + k] = 64 + k;
      } else {
        if (4 * k < 128) {  // increments: thread k owns dims [4k, 4k+4) (half a Philox block, like the fp32 kernel)
          const int j0 = 4 * k;
          float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j0 < d) {
            if (DET) z = load_row4(zS + ((int64_t)p * M_I + m) * d, j0, d, (d & 3) == 0);
            else z = normal4(seed, gp, kStreamIntegral, (uint32_t)m, (uint32_t)k);
          }
          C.zs[j0] = (j0 < d) ? z.x : 0.f;
          C.zs[j0 + 1] = (j0 + 1 < d) ? z.y : 0.f;
          C.zs[j0 + 2] = (j0 + 2 < d) ? z.z : 0.f;
          C.zs[j0 + 3] = (j0 + 3 < d) ? z.w : 0.f;
        } else {  // SDGD coordinates: thread 32 + b owns Philox block b = coordinates 4b .. 4b+3 (data.py:497-502)
          const int b = k - 32;
          int jj[4];
          if (DET) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              jj[q] = (4 * b + q < v) ? min(max(__ldg(idx_in + ((int64_t)p * M_I + m) * v + 4 * b + q), 0), d - 1) : 0;
          } else {
            const uint4 rr = philox_at(seed, gp, kStreamIdx, (uint32_t)m, (uint32_t)b);
            jj[0] = (int)__umulhi(rr.x, (uint32_t)d), jj[1] = (int)__umulhi(rr.y, (uint32_t)d);
            jj[2] = (int)__umulhi(rr.z, (uint32_t)d), jj[3] = (int)__umulhi(rr.w, (uint32_t)d);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) C.idx[4 * b + q] = jj[q];
        }
      }
      named_bar_sync(10 + h, 64);
      // ---- source term of the exact solution at (s, X = x + a Z) (equations.py:457-466): first warp of the half ----
      if (hw == 0) {
        float pk[kMaxGbmNodes];
#pragma unroll
        for (int q = 0; q < kMaxGbmNodes; ++q) pk[q] = 0.f;
        for (int j = hl; j < d; j += 32) {
          const float X = fmaf(a, C.zs[j], S.xs[j]);
#pragma unroll
          for (int q = 0; q < kMaxGbmNodes; ++q) pk[q] = fmaf(S.gw[q][j], X, pk[q]);
        }
        float sn[kMaxGbmNodes], ut = 0.f, lap = 0.f;
#pragma unroll
        for (int q = 0; q < kMaxGbmNodes; ++q) {
          sn[q] = 0.f;
          if (q < P.gbm_nodes) {
            const float ak = warp_sum(pk[q]) + S.w0[q] * s;
            sn[q] = sinf(ak);
            ut = fmaf(P.gbm_v[q] * S.w0[q], cosf(ak), ut);
            lap = fmaf(-P.gbm_v[q] * S.gw2[q], sn[q], lap);
          }
        }
        float ab = 0.f;
        for (int j = hl; j < d; j += 32) {
          float hj = 0.f;
#pragma unroll
          for (int q = 0; q < kMaxGbmNodes; ++q)
            if (q < P.gbm_nodes) hj = fmaf(-P.gbm_v[q] * S.gw[q][j] * S.gw[q][j], sn[q], hj);
          ab += fabsf(hj);
        }
        ab = warp_sum(ab);
        if (hl == 0) C.s = s, C.a = a, C.src = -ut - 0.5f * lap - 0.25f * ab;
      }
      // ---- value chain: thread k owns hidden unit k of every layer ----
      {
        // the increments (and below the activations) are read four at a time: one broadcast LDS.128 instead of four LDS.32
        float z0 = 0.f, z1 = 0.f;
        int j = 0;
        for (; j + 3 < d; j += 4) {
          const float4 zz = *reinterpret_cast<const float4*>(&C.zs[j]);
          z0 = fmaf(S.W1t[j][k], zz.x, z0);
          z1 = fmaf(S.W1t[j + 1][k], zz.y, z1);
          z0 = fmaf(S.W1t[j + 2][k], zz.z, z0);
          z1 = fmaf(S.W1t[j + 3][k], zz.w, z1);
        }
        for (; j < d; ++j) z0 = fmaf(S.W1t[j][k], C.zs[j], z0);
        float z = S.c0[k] + s * S.wt[k] + a * (z0 + z1);
        int pp = 0;
        for (int l = 0;; ++l) {
          const float e = __expf(fminf(z, 0.f));
          const bool pos = z > 0.f;
          C.e1[l][k] = __float2half_rn(pos ? 1.f : e);
          C.e2[l][k] = __float2half_rn(pos ? 0.f : e);
          if (l == n_hid) break;
          S.hbuf[h][pp][k] = pos ? z : e - 1.f;
          named_bar_sync(10 + h, 64);
          float y0 = 0.f, y1 = 0.f;
#pragma unroll 4
          for (int kk = 0; kk < kH; kk += 4) {
            const float4 hh = *reinterpret_cast<const float4*>(&S.hbuf[h][pp][kk]);
            y0 = fmaf(S.Wht[l][kk][k], hh.x, y0);
            y1 = fmaf(S.Wht[l][kk + 1][k], hh.y, y1);
            y0 = fmaf(S.Wht[l][kk + 2][k], hh.z, y0);
            y1 = fmaf(S.Wht[l][kk + 3][k], hh.w, y1);
          }
          z = S.bh[l][k] + (y0 + y1);
          pp ^= 1;
        }
      }
      __syncwarp();
      if (hl == 0) mbar_arrive(bar(&S.bar_ctx_full[cs]));  // (release: orders this warp's context stores)
    }
  } else {
    // =================================== JET team ===================================
    const int row = (warp & 3) * 32 + lane, hf = warp >> 2;  // direction row = TMEM lane; this warp's 32 columns
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(hf * 32);
    uint32_t ph_ctx = 0u, ph_acc[2] = {0u, 0u};
    float G = 0.f, fsum = 0.f;
    const float alpha_term = 0.5f * (1.f - P.alpha) * (float)d, abs_term = 0.25f * (float)d, inv_v = 1.f / (float)v;
    float src_b = 0.f;
    for (int r0 = 0; r0 < R; r0 += 2) {
      const int nrow = min(2, R - r0);
      // ---- first-layer jets of both rows: dh = e1 w, ddh = e2 w^2 with w = W1[:, 1 + j] ----
      for (int sl = 0; sl < nrow; ++sl) {
        const int r = r0 + sl, cs = r & 3;
        const RowCtx& C = S.ctx[cs];
        mbar_wait(bar(&S.bar_ctx_full[cs]), (ph_ctx >> cs) & 1u, 3);
        ph_ctx ^= 1u << cs;
        const int ndir = (r == 0) ? d : v;
        const int j = (row < ndir) ? C.idx[row] : -1;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const int k0 = hf * 32 + c8 * 8;
          uint4 wd = make_uint4(0, 0, 0, 0), wdd = wd;
          if (j >= 0) {
            const uint4 w8 = *reinterpret_cast<const uint4*>(&S.W1h[j][k0]);
            const uint4 e1 = *reinterpret_cast<const uint4*>(&C.e1[0][k0]), e2 = *reinterpret_cast<const uint4*>(&C.e2[0][k0]);
            wd = make_uint4(h2_mul(e1.x, w8.x), h2_mul(e1.y, w8.y), h2_mul(e1.z, w8.z), h2_mul(e1.w, w8.w));
            wdd = make_uint4(h2_mul(h2_mul(e2.x, w8.x), w8.x), h2_mul(h2_mul(e2.y, w8.y), w8.y),
                             h2_mul(h2_mul(e2.z, w8.z), w8.z), h2_mul(h2_mul(e2.w, w8.w), w8.w));
          }
          *reinterpret_cast<uint4*>(&S.A[sl][0][0] + a_off(row, k0)) = wd;
          *reinterpret_cast<uint4*>(&S.A[sl][1][0] + a_off(row, k0)) = wdd;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(&S.bar_a_ready[sl]));
      }
      // ---- hidden layers ----
      for (int l = 0; l < n_hid; ++l)
        for (int sl = 0; sl < nrow; ++sl) {
          const int r = r0 + sl;
          const RowCtx& C = S.ctx[r & 3];
          mbar_wait(bar(&S.bar_acc[sl]), ph_acc[sl], 4);
          ph_acc[sl] ^= 1u;
          tc_fence_after();
          const bool last = l == n_hid - 1;
          uint32_t accu = 0u;
          float pj = 0.f;
#pragma unroll
          for (int c16 = 0; c16 < 2; ++c16) {
            uint32_t dz[16], ddz[16];
            tmem_ld16(trow + (uint32_t)(sl * 128 + c16 * 16), dz);
            tmem_ld16(trow + (uint32_t)(sl * 128 + 64 + c16 * 16), ddz);
            const int k0 = hf * 32 + c16 * 16;
            const uint4 ea = *reinterpret_cast<const uint4*>(&C.e1[l + 1][k0]), eb = *reinterpret_cast<const uint4*>(&C.e1[l + 1][k0 + 8]);
            const uint4 fa = *reinterpret_cast<const uint4*>(&C.e2[l + 1][k0]), fb = *reinterpret_cast<const uint4*>(&C.e2[l + 1][k0 + 8]);
            const uint32_t e1v[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
            const uint32_t e2v[8] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
            tmem_ld_wait();
            uint32_t od[8], odd[8];
#pragma unroll
            for (int i2 = 0; i2 < 8; ++i2) {
              const uint32_t z2 = pack_f16x2(__uint_as_float(dz[2 * i2]), __uint_as_float(dz[2 * i2 + 1]));
              const uint32_t zz2 = pack_f16x2(__uint_as_float(ddz[2 * i2]), __uint_as_float(ddz[2 * i2 + 1]));
              od[i2] = h2_mul(e1v[i2], z2);
              odd[i2] = h2_fma(h2_mul(e2v[i2], z2), z2, h2_mul(e1v[i2], zz2));
            }
            if (!last) {
              *reinterpret_cast<uint4*>(&S.A[sl][0][0] + a_off(row, k0)) = make_uint4(od[0], od[1], od[2], od[3]);
              *reinterpret_cast<uint4*>(&S.A[sl][0][0] + a_off(row, k0 + 8)) = make_uint4(od[4], od[5], od[6], od[7]);
              *reinterpret_cast<uint4*>(&S.A[sl][1][0] + a_off(row, k0)) = make_uint4(odd[0], odd[1], odd[2], odd[3]);
              *reinterpret_cast<uint4*>(&S.A[sl][1][0] + a_off(row, k0 + 8)) = make_uint4(odd[4], odd[5], odd[6], odd[7]);
            } else {
              // u_jj = w_out . ddh over this warp's columns (fp16x2 partial sums, flushed to fp32 every 16 columns)
#pragma unroll
              for (int i2 = 0; i2 < 8; ++i2) accu = h2_fma(S.wout_h[(k0 >> 1) + i2], odd[i2], accu);
              const float2 f2 = h2_unpack(accu);
              pj += f2.x + f2.y;
              accu = 0u;
            }
          }
          if (!last) {
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(&S.bar_a_ready[sl]));
          } else {
            S.red[sl][hf][row] = pj;
            tc_fence_before();
          }
        }
      // ---- the rows' f, control variate, Bismut sums (jet team barriers 1, 2: 256 threads) ----
      named_bar_sync(1, kJetWarps * 32);
      if (r0 == 0) {  // baseline row: full Hessian diagonal at (t, x) (data.py:1276-1278) and its source term
        if (hf == 0) S.diag[row] = (row < d) ? S.red[0][0][row] + S.red[0][1][row] : 0.f;
        src_b = S.ctx[0].src;
        named_bar_sync(1, kJetWarps * 32);
      }
      if (hf == 0) {
        for (int sl = 0; sl < nrow; ++sl) {
          const int r = r0 + sl;
          if (r == 0) continue;
          const RowCtx& C = S.ctx[r & 3];
          float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
          if (row < v) {
            const float ujj = S.red[sl][0][row] + S.red[sl][1][row];
            const float dj = S.diag[C.idx[row]];
            a1 = ujj, a2 = fabsf(ujj), b1 = dj, b2 = fabsf(dj);
          }
          a1 = warp_sum(a1), a2 = warp_sum(a2), b1 = warp_sum(b1), b2 = warp_sum(b2);
          if (lane == 0) S.sums[sl][warp][0] = a1, S.sums[sl][warp][1] = a2, S.sums[sl][warp][2] = b1, S.sums[sl][warp][3] = b2;
        }
      }
      named_bar_sync(2, kJetWarps * 32);
      for (int sl = 0; sl < nrow; ++sl) {
        const int r = r0 + sl;
        const RowCtx& C = S.ctx[r & 3];
        if (r > 0) {
          float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) a1 += S.sums[sl][w][0], a2 += S.sums[sl][w][1], b1 += S.sums[sl][w][2], b2 += S.sums[sl][w][3];
          const float f = alpha_term * (a1 * inv_v) + abs_term * (a2 * inv_v) + C.src;
          const float fbv = alpha_term * (b1 * inv_v) + abs_term * (b2 * inv_v) + src_b;
          fsum += dT * f;
          const float coef = dT * (f - fbv) / C.a;
          if (hf == 0 && row < d) G = fmaf(coef, C.zs[row], G);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(&S.bar_ctx_empty[r & 3]));  // the context slot may be refilled
      }
    }
    // ---- y += integral part; clip (data.py:222) ----
    float* yp = y + p * (int64_t)(1 + d);
    const float invM = 1.f / (float)M_I;
    const bool clip = bound > 0.f && isfinite(bound);
    if (hf == 0 && row < d) {
      float val = yp[1 + row] + G * invM;
      if (clip) val = fminf(fmaxf(val, -bound), bound);
      yp[1 + row] = val;
    }
    if (tid == 0) {
      float val = yp[0] + fsum * invM;
      if (clip) val = fminf(fmaxf(val, -bound), bound);
      yp[0] = val;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

}  // namespace

int launch_gbm_tc2(const ProblemDev& P, const MlpDev& N, const float* tx, int64_t n, int M_I, uint64_t seed, int64_t base,
                   const float* u_s, const float* zS, const int32_t* idx, float bound, float* y, cudaStream_t st) {
  if (n <= 0) return DPI_OK;
  const size_t smem = sizeof(Smem) + 1024;
  auto kd = gbm_tc2_kernel<true>;
  auto kp = gbm_tc2_kernel<false>;
  DPI_CUDA_CHECK(cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DPI_CUDA_CHECK(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (zS)
    kd<<<(unsigned)n, kThreads, smem, st>>>(P, N, tx, M_I, seed, base, u_s, zS, idx, bound, y);
  else
    kp<<<(unsigned)n, kThreads, smem, st>>>(P, N, tx, M_I, seed, base, u_s, zS, idx, bound, y);
  DPI_CUDA_CHECK(cudaGetLastError());
  return DPI_OK;
}

bool gbm_tc2_eligible(const ProblemDev& P, const MlpDev& N) {
  return P.nx <= 128 && P.sdgd_v >= 1 && P.sdgd_v <= 128 && N.HP == 64 && N.n_linear >= 3 && N.n_linear - 2 <= kMaxHid &&
         N.d_pad == 128 && sizeof(Smem) + 1024 <= 232448;
}

}  // namespace dpi
