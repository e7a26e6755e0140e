// Third tcgen05 integral estimator for `GBMEquationComplexExact` with SDGD Hessian sampling (BASELINE config 4).
// Same mathematics as dpi_gbm_simt.cu / dpi_gbm_tc2.cu (picard/data.py:471-527, :497-502, :1226-1325;
// picard/equations.py:457-466).  ncu of dpi_gbm_tc2.cu (profiles/r02_gbm_tc2_v1_*): shared-memory wavefronts at 66 % of
// peak -- 2660 per path row, of which 570 are the fp32 value chain streaming its weights from shared memory, 480 the
// gather of first-layer weight rows by sampled coordinate (with its bank conflicts), 130 the output weights.
// What changed:
//   * directions in NATURAL order: operand row j of every path row is coordinate direction j (all d <= 128 of them; the
//     MMA tile has 128 rows whatever v is), so a jet thread keeps its row of W1 in registers for the whole kernel and
//     the first-layer jets need no gather.  The v sampled coordinates (with replacement, data.py:497-502) enter as
//     multiplicities: sum_i u_{j_i j_i} = sum_j cnt_j u_jj, the prep team histograms them;
//   * the value chain keeps its WEIGHTS IN REGISTERS: prep thread (k, q) owns hidden unit k and a quarter of the
//     reduction (32 inputs of layer 1, 16 of every hidden layer), two path rows at a time, partial sums combined by two
//     shuffles; what it reads from shared memory are broadcast 16-byte loads of the increments / activations;
//   * the source term of the exact solution moved to three warps of its own (one row each, round robin);
//   * FOUR path rows in flight on the jet team (TMEM 4 x 128 columns) instead of two;
//   * fp16 ACCUMULATORS (one half per TMEM cell, read back two columns per register): the jet team converts nothing,
//     and ELU' / ELU'' of a unit travel in one fp16 (+1 or -e^z) instead of two vectors.
// One CTA per collocation point, rows in a fixed order: labels do not depend on the launch geometry.
#include <cstdio>

#include "dpi_common.cuh"
#include "dpi_kernels.h"
#include "dpi_tc_common.cuh"

namespace dpi {
namespace {
using namespace tc;

constexpr int kH = 64, kMaxHid = 3, kNCtx = 8;
#ifndef DPI_GBM3_SLOTS
#define DPI_GBM3_SLOTS 4
#endif
constexpr int kSlots = DPI_GBM3_SLOTS;  // path rows in flight on the jet team
constexpr int kJetWarps = 8, kPrepWarp0 = 8, kPrepWarps = 8, kSrcWarp0 = 16, kSrcWarps = 3, kMmaWarp = 19, kThreads = 32 * 20;
constexpr int kPrepThreads = 32 * kPrepWarps;
// Register split (setmaxnreg, per warpgroup; a sub-partition holds one warp of each of the five warpgroups, 5 x 96
// registers per lane at launch): jet team 2 x kJetRegs, prep team 2 x 96, source-term + MMA warps kAuxRegs.
#ifndef DPI_GBM3_JET_REGS
#define DPI_GBM3_JET_REGS 120
#endif
#ifndef DPI_GBM3_AUX_REGS
#define DPI_GBM3_AUX_REGS 40
#endif
constexpr int kJetRegs = DPI_GBM3_JET_REGS, kAuxRegs = DPI_GBM3_AUX_REGS;
static_assert(2 * kJetRegs + 2 * 96 + kAuxRegs <= 480, "register split exceeds the launch allocation");
constexpr uint32_t kOpBytes = 128 * 128;  // one [128 rows x 64 k] fp16 operand image

// a prep thread reads a quarter of a vector with 16-byte loads; four quarters 128 bytes apart would share their banks,
// so every quarter is followed by 16 bytes of padding
__device__ __forceinline__ int zi(int j) { return j + 4 * (j >> 5); }  // increments: quarters of 32
__device__ __forceinline__ int hi(int k) { return k + 4 * (k >> 4); }  // activations: quarters of 16

struct alignas(16) RowCtx {
  float zs[144];               // index zi(j)
  int cnt[128];                // multiplicity of coordinate j among the row's v sampled directions
  // ELU' and ELU'' of the row's value chain in ONE fp16 per unit and layer (l = first layer .. last hidden layer):
  // +1 where z > 0, -e^z elsewhere; ELU' = |.|, ELU'' = max(-., 0)
  __half es[kMaxHid + 1][kH];
  float s, a, src, pad;
};

struct Smem {
  alignas(1024) uint8_t A[kSlots][2][kOpBytes];   // [row slot][dz | ddz]
  alignas(1024) uint8_t Wimg[kMaxHid][kH * 128];  // hidden weights W_l[n][k] fp16, SWIZZLE_128B
  RowCtx ctx[kNCtx];
  alignas(16) float hbuf[2][2][80];               // prep: activation vectors, [layer parity][row of the pair][hi(k)]
  alignas(16) float xs[128];
  alignas(16) float gw[kMaxGbmNodes][128];        // GBM w_k[1+j]
  float gw2[kMaxGbmNodes], w0[kMaxGbmNodes];      // sum_j w_kj^2; w_k[0]
  float red[kSlots][2][128];                      // [slot][column half][direction]: output-layer partial dots
  float sums[kSlots][4][4];                       // [slot][warp][a1, a2, b1, b2]
  float gwv[kMaxGbmNodes][128];                   // v_k w_kj^2
  uint64_t bar_ctx_full[kNCtx], bar_ctx_empty[kNCtx], bar_draws[kNCtx], bar_a_ready[kSlots], bar_acc[kSlots];
  uint32_t tmem_ptr;
};

__device__ __forceinline__ uint32_t a_off(int row, int k) {  // byte offset of element (row, k) in a [rows x 64] fp16 image
  return (uint32_t)row * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
}

template <bool DET, int NH>
__global__ void __launch_bounds__(kThreads, 1)
gbm_tc3_kernel(const __grid_constant__ ProblemDev P, const __grid_constant__ MlpDev N, const float* __restrict__ tx,
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
  const int R = M_I + 1;  // row 0 = the baseline pseudo-row (t, x); row 1 + m = sample m
  auto bar = [&](const uint64_t* b) -> uint32_t { return smem_u32(b); };

  if (tid == 0) {
    if (sbase & 1023u) __trap();
    for (int i = 0; i < kNCtx; ++i) {
      mbar_init(bar(&S.bar_ctx_full[i]), kPrepWarps + 1);  // the prep warps and the row's source-term warp
      mbar_init(bar(&S.bar_draws[i]), 1);
      mbar_init(bar(&S.bar_ctx_empty[i]), kJetWarps);
    }
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(bar(&S.bar_a_ready[i]), kJetWarps);
      mbar_init(bar(&S.bar_acc[i]), 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(&S.tmem_ptr), 128 * kSlots);
  // ---- stage the point and what of the network lives in shared memory ----
  for (int i = tid; i < 128; i += kThreads) {
    S.xs[i] = (i < d) ? __ldg(txp + 1 + i) : 0.f;
    for (int k = 0; k < kMaxGbmNodes; ++k) {
      const float w = (k < P.gbm_nodes && i < d) ? P.gbm_w[k * (1 + d) + 1 + i] : 0.f;
      S.gw[k][i] = w;
      S.gwv[k][i] = (k < P.gbm_nodes) ? P.gbm_v[k] * w * w : 0.f;
    }
  }
  if (tid < kMaxGbmNodes) S.w0[tid] = (tid < P.gbm_nodes) ? P.gbm_w[tid * (1 + d)] : 0.f;
  for (int l = 0; l < NH; ++l)
    for (int i = tid; i < kH * kH; i += kThreads) {
      const int k = i / kH, n = i - k * kH;  // Wth[l][k][n] = W_l[n][k]
      *reinterpret_cast<__half*>(S.Wimg[l] + a_off(n, k)) = __float2half_rn(N.Wth[l][i]);
    }
  for (int i = tid; i < kNCtx * 128; i += kThreads) S.ctx[i >> 7].cnt[i & 127] = 0;
  // operand rows of padded directions must hold finite values: zero every slot once
  for (int i = tid; i < (int)(2 * kSlots * kOpBytes / 16); i += kThreads)
    reinterpret_cast<uint4*>(&S.A[0][0][0])[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid < kMaxGbmNodes) {
    float acc = 0.f;
    for (int j = 0; j < d; ++j) acc = fmaf(S.gw[tid][j], S.gw[tid][j], acc);
    S.gw2[tid] = acc;
  }
  __syncthreads();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&S.tmem_ptr);

  if (warp >= kSrcWarp0) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kAuxRegs));
  if (warp == kMmaWarp) {
    // =================================== MMA issuer ===================================
    const uint32_t leader = elect_one();
    // fp16 accumulators (one half per TMEM cell, read back two columns per register): the jet team converts nothing
    constexpr uint32_t idesc = make_idesc_f16_acc16(128, 64);
    uint32_t ph = 0u;
    for (int r0 = 0; r0 < R; r0 += kSlots) {
      const int nrow = min(kSlots, R - r0);
      for (int l = 0; l < NH; ++l)
        for (int sl = 0; sl < nrow; ++sl) {
          mbar_wait(bar(&S.bar_a_ready[sl]), (ph >> sl) & 1u, 1);
          ph ^= 1u << sl;
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
  } else {
    // =================================== source term ===================================
    // f of the exact solution at (s, X = x + a Z) (equations.py:457-466); warp w takes rows w, w + 3, ...
    // (one warp needs about two row times for it: the dot products, sin / cos of every node, the |.| sum over d)
    for (int r = warp - kSrcWarp0; r < R; r += kSrcWarps) {
      const int cs = r & (kNCtx - 1);
      RowCtx& C = S.ctx[cs];
      mbar_wait_relaxed(bar(&S.bar_draws[cs]), (uint32_t)(r / kNCtx) & 1u);  // the slot's (r / kNCtx)-th use
      const float s = C.s, a = C.a;
      float pk[kMaxGbmNodes];
#pragma unroll
      for (int q = 0; q < kMaxGbmNodes; ++q) pk[q] = 0.f;
      for (int j = lane; j < d; j += 32) {
        const float X = fmaf(a, C.zs[zi(j)], S.xs[j]);
#pragma unroll
        for (int q = 0; q < kMaxGbmNodes; ++q) pk[q] = fmaf(S.gw[q][j], X, pk[q]);
      }
      float sn[kMaxGbmNodes], ut = 0.f, lap = 0.f;
#pragma unroll
      for (int q = 0; q < kMaxGbmNodes; ++q) {
        sn[q] = 0.f;
        if (q < P.gbm_nodes) {
          const float ak = warp_sum(pk[q]) + S.w0[q] * s;
          // two-constant reduction to [-pi, pi], then the MUFU sine / cosine (absolute error ~1e-6: this is the 16-bit path)
          const float kq = rintf(ak * 0.15915494309189535f);
          const float ar = fmaf(kq, 1.7484555e-7f, fmaf(kq, -6.2831854820251465f, ak));
          float cs_;
          __sincosf(ar, &sn[q], &cs_);
          ut = fmaf(P.gbm_v[q] * S.w0[q], cs_, ut);
          lap = fmaf(-P.gbm_v[q] * S.gw2[q], sn[q], lap);
        }
      }
      float ab = 0.f;
      for (int j = lane; j < d; j += 32) {
        float hj = 0.f;
#pragma unroll
        for (int q = 0; q < kMaxGbmNodes; ++q) hj = fmaf(-S.gwv[q][j], sn[q], hj);
        ab += fabsf(hj);
      }
      ab = warp_sum(ab);
      if (lane == 0) {
        C.src = -ut - 0.5f * lap - 0.25f * ab;
        mbar_arrive(bar(&S.bar_ctx_full[cs]));
      }
    }
  }
  } else if (warp >= kPrepWarp0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(96));
    // =================================== PREP team ===================================
    // thread (k, q): hidden unit k of every layer, quarter q of every reduction; rows r0, r0 + 1 together
    const int pt = tid - 32 * kPrepWarp0, k = pt >> 2, q = pt & 3;
    float w1[32], wh[NH][16], bh[NH];
#pragma unroll
    for (int i = 0; i < 32; ++i) w1[i] = N.Wt1[(32 * q + i) * kH + k];  // packed [d_pad = 128][HP = 64], zero beyond d
#pragma unroll
    for (int l = 0; l < NH; ++l) {
#pragma unroll
      for (int i = 0; i < 16; ++i) wh[l][i] = N.Wth[l][(16 * q + i) * kH + k];
      bh[l] = N.bh[l][k];
    }
    const float wtk = N.wt[k];
    float c0;  // b1 + W_x x
    {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) acc = fmaf(w1[i], S.xs[32 * q + i], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      c0 = N.b1[k] + acc;
    }
    uint32_t ph_empty = 0u, used = 0u;
    for (int r0 = 0; r0 < R; r0 += 2) {
      const int npair = min(2, R - r0);
      for (int rr = 0; rr < npair; ++rr) {
        const int cs = (r0 + rr) & (kNCtx - 1);
        if (used & (1u << cs)) {
          mbar_wait_relaxed(bar(&S.bar_ctx_empty[cs]), (ph_empty >> cs) & 1u);
          ph_empty ^= 1u << cs;
        }
        used |= 1u << cs;
      }
      // ---- draws: 128 threads per row of the pair ----
      {
        const int rr = pt >> 7, u = pt & 127, r = r0 + rr, m = r - 1;
        if (rr < npair) {
          RowCtx& C = S.ctx[r & (kNCtx - 1)];
          if (u < 16) {  // increments: Philox block u = dims [8u, 8u + 8)
            float4 za = make_float4(0.f, 0.f, 0.f, 0.f), zb = za;
            const int j0 = 8 * u;
            if (r > 0 && j0 < d) {
              if (DET) {
                za = load_row4(zS + ((int64_t)p * M_I + m) * d, j0, d, (d & 3) == 0);
                zb = load_row4(zS + ((int64_t)p * M_I + m) * d, j0 + 4, d, (d & 3) == 0);
              } else {
                normal8(seed, gp, kStreamIntegral, (uint32_t)m, (uint32_t)u, za, zb);
              }
              if (j0 + 1 >= d) za.y = 0.f;
              if (j0 + 2 >= d) za.z = 0.f;
              if (j0 + 3 >= d) za.w = 0.f;
              if (j0 + 4 >= d) zb.x = 0.f;
              if (j0 + 5 >= d) zb.y = 0.f;
              if (j0 + 6 >= d) zb.z = 0.f;
              if (j0 + 7 >= d) zb.w = 0.f;
            }
            *reinterpret_cast<float4*>(&C.zs[zi(j0)]) = za;
            *reinterpret_cast<float4*>(&C.zs[zi(j0) + 4]) = zb;
          } else if (u >= 32 && u < 64) {  // SDGD coordinates: Philox block b = coordinates 4b .. 4b+3 (data.py:497-502)
            const int b = u - 32;
            if (r > 0 && 4 * b < v) {
              int jj[4];
              if (DET) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  jj[c] = (4 * b + c < v) ? min(max(__ldg(idx_in + ((int64_t)p * M_I + m) * v + 4 * b + c), 0), d - 1) : -1;
              } else {
                const uint4 rnd = philox_at(seed, gp, kStreamIdx, (uint32_t)m, (uint32_t)b);
                jj[0] = (int)__umulhi(rnd.x, (uint32_t)d), jj[1] = (int)__umulhi(rnd.y, (uint32_t)d);
                jj[2] = (int)__umulhi(rnd.z, (uint32_t)d), jj[3] = (int)__umulhi(rnd.w, (uint32_t)d);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  if (4 * b + c >= v) jj[c] = -1;
              }
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (jj[c] >= 0) atomicAdd(&C.cnt[jj[c]], 1);
            }
          } else if (u == 64) {
            float uu = 0.f;
            if (r > 0) uu = DET ? __ldg(u_s + p * M_I + m) : u01(philox_at(seed, gp, kStreamS, (uint32_t)m, 0u).x);
            C.s = uu * dT + t;
            C.a = sqrtf(uu * dT) * P.sigma;
          }
        }
      }
      named_bar_sync(3, kPrepThreads);
      if (pt == 0)
        for (int rr = 0; rr < npair; ++rr) mbar_arrive(bar(&S.bar_draws[(r0 + rr) & (kNCtx - 1)]));  // to the source-term warps
      // ---- first layer ----
      const RowCtx& C0 = S.ctx[r0 & (kNCtx - 1)];
      const RowCtx& C1 = S.ctx[(r0 + npair - 1) & (kNCtx - 1)];  // a lone last row is computed twice
      float z0, z1;
      {
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 u0 = *reinterpret_cast<const float4*>(&C0.zs[36 * q + i]);
          const float4 u1 = *reinterpret_cast<const float4*>(&C1.zs[36 * q + i]);
          a0 = fmaf(w1[i], u0.x, a0), b0 = fmaf(w1[i], u1.x, b0);
          a1 = fmaf(w1[i + 1], u0.y, a1), b1 = fmaf(w1[i + 1], u1.y, b1);
          a0 = fmaf(w1[i + 2], u0.z, a0), b0 = fmaf(w1[i + 2], u1.z, b0);
          a1 = fmaf(w1[i + 3], u0.w, a1), b1 = fmaf(w1[i + 3], u1.w, b1);
        }
        a0 += a1, b0 += b1;
        a0 += __shfl_xor_sync(0xffffffffu, a0, 1), b0 += __shfl_xor_sync(0xffffffffu, b0, 1);
        a0 += __shfl_xor_sync(0xffffffffu, a0, 2), b0 += __shfl_xor_sync(0xffffffffu, b0, 2);
        z0 = c0 + C0.s * wtk + C0.a * a0;
        z1 = c0 + C1.s * wtk + C1.a * b0;
      }
#pragma unroll
      for (int l = 0; l <= NH; ++l) {
        const float e0 = __expf(fminf(z0, 0.f)), e1v = __expf(fminf(z1, 0.f));
        const bool p0 = z0 > 0.f, p1 = z1 > 0.f;
        if (q == 0) {
          S.ctx[r0 & (kNCtx - 1)].es[l][k] = __float2half_rn(p0 ? 1.f : -e0);
          if (l < NH) S.hbuf[l & 1][0][hi(k)] = p0 ? z0 : e0 - 1.f;
        } else if (q == 1) {
          S.ctx[(r0 + npair - 1) & (kNCtx - 1)].es[l][k] = __float2half_rn(p1 ? 1.f : -e1v);
          if (l < NH) S.hbuf[l & 1][1][hi(k)] = p1 ? z1 : e1v - 1.f;
        }
        if (l < NH) {
          constexpr int kLastHid = NH > 0 ? NH - 1 : 0;
          const int ll = l < NH ? l : kLastHid;  // (static after unrolling; keeps the index in range for the compiler)
          named_bar_sync(4, kPrepThreads);
          float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 u0 = *reinterpret_cast<const float4*>(&S.hbuf[l & 1][0][20 * q + i]);
            const float4 u1 = *reinterpret_cast<const float4*>(&S.hbuf[l & 1][1][20 * q + i]);
            a0 = fmaf(wh[ll][i], u0.x, a0), b0 = fmaf(wh[ll][i], u1.x, b0);
            a1 = fmaf(wh[ll][i + 1], u0.y, a1), b1 = fmaf(wh[ll][i + 1], u1.y, b1);
            a0 = fmaf(wh[ll][i + 2], u0.z, a0), b0 = fmaf(wh[ll][i + 2], u1.z, b0);
            a1 = fmaf(wh[ll][i + 3], u0.w, a1), b1 = fmaf(wh[ll][i + 3], u1.w, b1);
          }
          a0 += a1, b0 += b1;
          a0 += __shfl_xor_sync(0xffffffffu, a0, 1), b0 += __shfl_xor_sync(0xffffffffu, b0, 1);
          a0 += __shfl_xor_sync(0xffffffffu, a0, 2), b0 += __shfl_xor_sync(0xffffffffu, b0, 2);
          z0 = bh[ll] + a0;
          z1 = bh[ll] + b0;
        }
      }
      __syncwarp();
      if (lane == 0)
        for (int rr = 0; rr < npair; ++rr) mbar_arrive(bar(&S.bar_ctx_full[(r0 + rr) & (kNCtx - 1)]));  // (release)
    }
  } else {
    // =================================== JET team ===================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kJetRegs));
    const int row = (warp & 3) * 32 + lane, hf = warp >> 2;  // direction row = TMEM lane = coordinate; this warp's 32 columns
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(hf * 32);
    uint32_t w1r[16];  // W1[hf*32 .. hf*32+31][1 + row] as fp16x2: the first-layer jet of direction `row` before ELU'
#pragma unroll
    for (int i = 0; i < 16; ++i)
      w1r[i] = (row < d) ? pack_f16x2(N.Wt1[row * kH + hf * 32 + 2 * i], N.Wt1[row * kH + hf * 32 + 2 * i + 1]) : 0u;
    uint32_t w2r[16];  // their squares (the first-layer ddh = ELU'' w^2)
#pragma unroll
    for (int i = 0; i < 16; ++i) w2r[i] = h2_mul(w1r[i], w1r[i]);
    uint32_t woh[16];  // the output weights of this warp's columns
#pragma unroll
    for (int i = 0; i < 16; ++i) woh[i] = pack_f16x2(N.wout[hf * 32 + 2 * i], N.wout[hf * 32 + 2 * i + 1]);
    // shared addresses of this thread's four 16-byte operand chunks (logical chunk hf*4 + c of operand row `row`)
    const uint32_t sA = smem_u32(&S.A[0][0][0]), sCtx = smem_u32(&S.ctx[0]);
    uint32_t poff[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) poff[c] = sA + (uint32_t)row * 128u + ((((uint32_t)(hf * 4 + c)) ^ ((uint32_t)row & 7u)) << 4);
    constexpr uint32_t kEsOff = (uint32_t)offsetof(RowCtx, es);
    uint32_t ph_ctx = 0u, ph_acc = 0u;
    float G = 0.f, facc = 0.f, srcsum = 0.f, diag = 0.f;
    const float alpha_term = 0.5f * (1.f - P.alpha) * (float)d, abs_term = 0.25f * (float)d, inv_v = 1.f / (float)v;
    float src_b = 0.f;
    for (int r0 = 0; r0 < R; r0 += kSlots) {
      const int nrow = min(kSlots, R - r0);
      // ---- first-layer jets: dh = e1 w, ddh = e2 w^2 ----
      for (int sl = 0; sl < nrow; ++sl) {
        const int cs = (r0 + sl) & (kNCtx - 1);
        mbar_wait(bar(&S.bar_ctx_full[cs]), (ph_ctx >> cs) & 1u, 3);
        ph_ctx ^= 1u << cs;
        const uint32_t sEs = sCtx + (uint32_t)cs * (uint32_t)sizeof(RowCtx) + kEsOff + (uint32_t)hf * 64u;
        const uint32_t sSlot = (uint32_t)sl * 2u * kOpBytes;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const uint4 es = lds128(sEs + 16u * c8);
          const uint4 e1 = make_uint4(h2_abs(es.x), h2_abs(es.y), h2_abs(es.z), h2_abs(es.w));
          const uint4 e2 = make_uint4(h2_negrelu(es.x), h2_negrelu(es.y), h2_negrelu(es.z), h2_negrelu(es.w));
          const uint32_t* w = &w1r[4 * c8];
          const uint4 wd = make_uint4(h2_mul(e1.x, w[0]), h2_mul(e1.y, w[1]), h2_mul(e1.z, w[2]), h2_mul(e1.w, w[3]));
          const uint32_t* w2 = &w2r[4 * c8];
          const uint4 wdd = make_uint4(h2_mul(e2.x, w2[0]), h2_mul(e2.y, w2[1]), h2_mul(e2.z, w2[2]), h2_mul(e2.w, w2[3]));
          sts128(poff[c8] + sSlot, wd.x, wd.y, wd.z, wd.w);
          sts128(poff[c8] + sSlot + kOpBytes, wdd.x, wdd.y, wdd.z, wdd.w);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(&S.bar_a_ready[sl]));
      }
      // ---- hidden layers ----
#pragma unroll
      for (int l = 0; l < NH; ++l)
        for (int sl = 0; sl < nrow; ++sl) {
          const uint32_t sEs = sCtx + (uint32_t)((r0 + sl) & (kNCtx - 1)) * (uint32_t)sizeof(RowCtx) + kEsOff +
                               (uint32_t)(l + 1) * (uint32_t)(kH * 2) + (uint32_t)hf * 64u;
          const uint32_t sSlot = (uint32_t)sl * 2u * kOpBytes;
          mbar_wait(bar(&S.bar_acc[sl]), (ph_acc >> sl) & 1u, 4);
          ph_acc ^= 1u << sl;
          tc_fence_after();
          const bool last = l == NH - 1;
          uint32_t accu = 0u;
          float pj = 0.f;
          // all four accumulator loads of the row are in flight before the first conversion
          uint32_t dz[2][8], ddz[2][8];  // packed fp16x2 pairs of adjacent columns
          tmem_ld8_pack(trow + (uint32_t)(sl * 128), dz[0]);
          tmem_ld8_pack(trow + (uint32_t)(sl * 128 + 64), ddz[0]);
          tmem_ld8_pack(trow + (uint32_t)(sl * 128 + 16), dz[1]);
          tmem_ld8_pack(trow + (uint32_t)(sl * 128 + 64 + 16), ddz[1]);
          uint4 ev[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) ev[c] = lds128(sEs + 16u * c);
          tmem_ld_wait();
#pragma unroll
          for (int c16 = 0; c16 < 2; ++c16) {
            const uint32_t esv[8] = {ev[2 * c16].x, ev[2 * c16].y, ev[2 * c16].z, ev[2 * c16].w,
                                     ev[2 * c16 + 1].x, ev[2 * c16 + 1].y, ev[2 * c16 + 1].z, ev[2 * c16 + 1].w};
            uint32_t od[8], odd[8];
#pragma unroll
            for (int i2 = 0; i2 < 8; ++i2) {
              const uint32_t z2 = dz[c16][i2], zz2 = ddz[c16][i2];
              const uint32_t e1 = h2_abs(esv[i2]), e2 = h2_negrelu(esv[i2]);
              od[i2] = h2_mul(e1, z2);
              odd[i2] = h2_fma(h2_mul(e2, z2), z2, h2_mul(e1, zz2));
            }
            if (!last) {
              sts128(poff[2 * c16] + sSlot, od[0], od[1], od[2], od[3]);
              sts128(poff[2 * c16 + 1] + sSlot, od[4], od[5], od[6], od[7]);
              sts128(poff[2 * c16] + sSlot + kOpBytes, odd[0], odd[1], odd[2], odd[3]);
              sts128(poff[2 * c16 + 1] + sSlot + kOpBytes, odd[4], odd[5], odd[6], odd[7]);
            } else {
              // u_jj = w_out . ddh over this warp's columns (fp16x2 partial sums, flushed to fp32 every 16 columns)
#pragma unroll
              for (int i2 = 0; i2 < 8; ++i2) accu = h2_fma(woh[8 * c16 + i2], odd[i2], accu);
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
      // f - f_baseline = sum_j cnt_j [alpha_term (u_jj - diag_j) + abs_term (|u_jj| - |diag_j|)] / v + src - src_b needs
      // ONE sum over the directions per row; the sum of f over the rows is linear in the per-direction terms and is
      // reduced once, at the end of the kernel
      named_bar_sync(1, kJetWarps * 32);
      if (hf == 0) {
        for (int sl = 0; sl < nrow; ++sl) {
          const int r = r0 + sl;
          RowCtx& C = S.ctx[r & (kNCtx - 1)];
          const float ujj = (row < d) ? S.red[sl][0][row] + S.red[sl][1][row] : 0.f;
          if (r == 0) {  // baseline row: the Hessian diagonal at (t, x) (data.py:1276-1278) stays in a register
            diag = ujj;
            continue;
          }
          const float c = (float)C.cnt[row] * inv_v;
          C.cnt[row] = 0;  // the context comes back to the prep team with its histogram cleared
          const float fu = alpha_term * ujj + abs_term * fabsf(ujj), fb = alpha_term * diag + abs_term * fabsf(diag);
          facc = fmaf(c, fu, facc);
          const float w = warp_sum(c * (fu - fb));
          if (lane == 0) S.sums[sl][warp][0] = w;
        }
      }
      if (r0 == 0) src_b = S.ctx[0].src;
      named_bar_sync(2, kJetWarps * 32);
      for (int sl = 0; sl < nrow; ++sl) {
        const int r = r0 + sl;
        const RowCtx& C = S.ctx[r & (kNCtx - 1)];
        if (r > 0) {
          const float src = C.src;
          const float df = (S.sums[sl][0][0] + S.sums[sl][1][0]) + (S.sums[sl][2][0] + S.sums[sl][3][0]) + (src - src_b);
          srcsum += src;
          const float coef = dT * df / C.a;
          if (hf == 0 && row < d) G = fmaf(coef, C.zs[zi(row)], G);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(&S.bar_ctx_empty[r & (kNCtx - 1)]));  // the context slot may be refilled
      }
    }
    // sum of f over the rows: the per-direction terms of the hf == 0 threads plus the source terms
    facc = warp_sum(facc);
    if (hf == 0 && lane == 0) S.sums[0][warp][1] = facc;
    named_bar_sync(1, kJetWarps * 32);
    const float fsum = dT * ((S.sums[0][0][1] + S.sums[0][1][1]) + (S.sums[0][2][1] + S.sums[0][3][1]) + srcsum);
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
    tmem_dealloc(tmem, 128 * kSlots);
  }
}

template <int NH>
int launch_nh(const ProblemDev& P, const MlpDev& N, const float* tx, int64_t n, int M_I, uint64_t seed, int64_t base,
              const float* u_s, const float* zS, const int32_t* idx, float bound, float* y, cudaStream_t st) {
  const size_t smem = sizeof(Smem) + 1024;
  auto kd = gbm_tc3_kernel<true, NH>;
  auto kp = gbm_tc3_kernel<false, NH>;
  DPI_CUDA_CHECK(cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DPI_CUDA_CHECK(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (zS)
    kd<<<(unsigned)n, kThreads, smem, st>>>(P, N, tx, M_I, seed, base, u_s, zS, idx, bound, y);
  else
    kp<<<(unsigned)n, kThreads, smem, st>>>(P, N, tx, M_I, seed, base, u_s, zS, idx, bound, y);
  DPI_CUDA_CHECK(cudaGetLastError());
  return DPI_OK;
}

}  // namespace

int launch_gbm_tc3(const ProblemDev& P, const MlpDev& N, const float* tx, int64_t n, int M_I, uint64_t seed, int64_t base,
                   const float* u_s, const float* zS, const int32_t* idx, float bound, float* y, cudaStream_t st) {
  if (n <= 0) return DPI_OK;
  switch (N.n_linear - 2) {
    case 1: return launch_nh<1>(P, N, tx, n, M_I, seed, base, u_s, zS, idx, bound, y, st);
    case 2: return launch_nh<2>(P, N, tx, n, M_I, seed, base, u_s, zS, idx, bound, y, st);
    case 3: return launch_nh<3>(P, N, tx, n, M_I, seed, base, u_s, zS, idx, bound, y, st);
  }
  return DPI_ERR_UNSUPPORTED;
}

bool gbm_tc3_eligible(const ProblemDev& P, const MlpDev& N) {
  // natural order computes all d directions per path row: not worth it when only a few are sampled
  return P.nx <= 128 && P.sdgd_v >= 1 && P.sdgd_v <= 128 && 4 * P.sdgd_v >= P.nx && N.HP == 64 && N.n_linear >= 3 && N.n_linear - 2 <= kMaxHid &&
         N.d_pad == 128 && sizeof(Smem) + 1024 <= 232448;
}

}  // namespace dpi
