// bf16 tcgen05 integral estimator for the Burgers-type equation `Cha` (u_k = MLP [1+d,128,128,128,128,1], d <= 128).
//
// Same mathematics as dpi_mlp_simt.cu (picard/data.py:471-527,350-366,1226-1325; equations.py:297-302;
// solution.py:123-135) with the three hidden 128x128 GEMMs (value and tangent) and the first-layer GEMM on the
// 5th-generation tensor cores:
//
//   * one persistent CTA per SM; the four weight matrices live in shared memory for the whole kernel as bf16
//     K-major SWIZZLE_128B operand images (4 x 32 KB, fetched once with TMA bulk copies);
//   * a work unit is (collocation point, group of 8 tiles); a tile is 128 MC samples = 128 TMEM lanes;
//   * producer warpgroups (2 x 4 warps, alternating tiles) draw the Brownian increments with Philox4x32-10 +
//     Box-Muller in registers, round them to bf16 and write the [128 x d] A-operand tile; they KEEP their
//     increments in registers and, once the tile's Bismut coefficients are known, accumulate
//     G_j += coef_m Z_mj there, so increments never leave the SM;
//   * one thread issues tcgen05.mma: layer 1 (Z . W_x^T) into accumulator set 0, hidden layers alternate between the
//     two TMEM accumulator sets (value 128 cols + tangent 128 cols each = all 512 columns);
//   * 8 epilogue warps (TMEM lane quarter x column half) read the accumulators with tcgen05.ld, apply
//     bias/ELU/ELU' in fp32, and write the next layer's bf16 A operands (value rows and tangent rows) straight
//     into the swizzled shared-memory layout.  They signal the MMA thread per 32-column K chunk, so the next
//     layer's MMAs run underneath the remaining epilogue work of the current layer;
//   * the last layer is a dot product in the epilogue (fp32), followed by f, the Bismut coefficient and the
//     value-label partial sum.
//
// Partial sums go to a small scratch ([unit][producer group][row class][128]) and a finish kernel adds them
// in a fixed order, so labels are bitwise independent of the launch geometry and of how points are sharded.
#include <cstdio>
#include <vector>

#include "dpi_common.cuh"
#include "dpi_kernels.h"
#include "dpi_tc_common.cuh"

namespace dpi {
namespace {
using namespace tc;

// 8 epilogue warps + 8 producer warps + one warpgroup whose first warp issues the MMAs.  The register file is
// per SM sub-partition (16K registers, warp w lives on sub-partition w % 4): five warps per sub-partition cap the
// launch at 96 registers/thread, so the MMA warpgroup hands its registers back (setmaxnreg.dec 32) and the sixteen
// worker warps grow to 112 (4 x 112 + 32 = 480 <= 512 per lane).
constexpr int kTcThreads = 640;
constexpr int kMmaWarp = 16;
constexpr int kWorkerRegs = 112, kMmaRegs = 32;
constexpr float kLog2e = 1.4426950408889634f;

// shared memory map (dynamic, 1024-byte aligned)
constexpr uint32_t kOffW = 0;                          // 4 x 32 KB weight images
constexpr uint32_t kOffHV = 4 * kOperandBytes;         // value rows of the next layer's A operand
constexpr uint32_t kOffHT = kOffHV + kOperandBytes;    // tangent rows
constexpr uint32_t kOffZ = kOffHT + kOperandBytes;     // increments (A operand of layer 1)
constexpr uint32_t kOffBars = kOffZ + kOperandBytes;   // 16 mbarriers
constexpr uint32_t kOffTmemPtr = kOffBars + 128;
constexpr uint32_t kOffC0 = kOffBars + 160;            // [128] per-point first-layer constant
constexpr uint32_t kOffCoef = kOffC0 + 512;            // [2][128] Bismut coefficients (one buffer per producer group)
constexpr uint32_t kOffXch = kOffCoef + 1024;          // [128] float2: output-layer partial sums of column half 1
constexpr uint32_t kTcSmemBytes = kOffXch + 1024;      // 232,096 B <= 232,448 B

enum Bar : int { kBarW = 0, kBarZFull = 1, kBarZEmpty = 2, kBarAcc0 = 3, kBarAcc1 = 4, kBarReady0 = 5, kBarCoef0 = 9 };

struct TcArgs {
  const void* W;  // 4 consecutive operand images: W1 (x columns), W2, W3, W4
  const float* tx;
  int64_t n_units;
  int64_t point_base;
  uint64_t seed;
  const float* u_s;
  const float* zS;
  const float* fb;  // [n] control variate f(t,x)
  const float* c0;  // [n][128]
  float* partG;     // [units][16][128]
  float* partF;     // [units][4]
  int M_I;
  int units_per_point;
  int tiles_per_point;
  int ksteps1;  // ceil(d/16)
  int nq;       // ceil(d/4) quads of dims
  int cls;      // row classes of the producer mapping
};

// All activations are carried scaled by log2(e): with z' = log2e z the ELU pair needs one op less per element,
//   e = 2^min(z',0) = ELU'(z),   h' = log2e ELU(z) = max(z', e log2e - log2e),   tangent' = e * (log2e dz),
// and because every layer is linear in its input only the biases (x log2e), the first-layer constants and the
// output weights (x ln2) change -- the hidden weight matrices are used as they are (TcConsts holds the scaled
// vectors; see tc_pack_network).
//
// One hidden-layer epilogue for this warp's 64 columns: TMEM -> bias/ELU/ELU' -> bf16 A operands of the next layer.
// FIRST: z' = c0' + s' w_t + a' acc, tangent' = e * wsum' (constant first-layer tangent);  else z' = acc + b'[layer].
template <bool FIRST>
__device__ __forceinline__ void epi_hidden(uint8_t* smem, uint32_t sbase, const TcConsts& C, uint32_t tacc, int hf,
                                           int lane, uint32_t rowbase, uint32_t rx, int layer, float s, float a) {
  const float* c0_s = reinterpret_cast<const float*>(smem + kOffC0);
  const float* bias = C.b[layer];
#pragma unroll 1
  for (int jj = 0; jj < 2; ++jj) {
    const int col0 = hf * 64 + jj * 32;
    uint32_t v[32], tv[32];
    tmem_ld32(tacc + col0, v);
    if (!FIRST) tmem_ld32(tacc + 128 + col0, tv);
    tmem_ld_wait();
    uint8_t* const hv_row = smem + kOffHV + rowbase;
    uint8_t* const ht_row = smem + kOffHT + rowbase;
#pragma unroll
    for (int c8 = 0; c8 < 32; c8 += 8) {
      uint32_t hv[4], ht[4];
#pragma unroll
      for (int k4 = 0; k4 < 8; k4 += 4) {
        const int col = col0 + c8 + k4;
        float zb[4], tg[4];
        if (FIRST) {
          const float4 w = *reinterpret_cast<const float4*>(&C.wt[col]);
          const float4 ws = *reinterpret_cast<const float4*>(&C.wsum[col]);
          const float4 c0 = *reinterpret_cast<const float4*>(&c0_s[col]);
          zb[0] = fmaf(a, __uint_as_float(v[c8 + k4 + 0]), fmaf(s, w.x, c0.x));
          zb[1] = fmaf(a, __uint_as_float(v[c8 + k4 + 1]), fmaf(s, w.y, c0.y));
          zb[2] = fmaf(a, __uint_as_float(v[c8 + k4 + 2]), fmaf(s, w.z, c0.z));
          zb[3] = fmaf(a, __uint_as_float(v[c8 + k4 + 3]), fmaf(s, w.w, c0.w));
          tg[0] = ws.x, tg[1] = ws.y, tg[2] = ws.z, tg[3] = ws.w;
        } else {
          const float4 b = *reinterpret_cast<const float4*>(&bias[col]);
          zb[0] = __uint_as_float(v[c8 + k4 + 0]) + b.x;
          zb[1] = __uint_as_float(v[c8 + k4 + 1]) + b.y;
          zb[2] = __uint_as_float(v[c8 + k4 + 2]) + b.z;
          zb[3] = __uint_as_float(v[c8 + k4 + 3]) + b.w;
#pragma unroll
          for (int i = 0; i < 4; ++i) tg[i] = __uint_as_float(tv[c8 + k4 + i]);
        }
        float h[4], dh[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float e = exp2f(fminf(zb[i], 0.f));
          h[i] = fmaxf(zb[i], fmaf(e, kLog2e, -kLog2e));
          dh[i] = e * tg[i];
        }
        hv[k4 / 2 + 0] = pack_bf16x2(h[0], h[1]);
        hv[k4 / 2 + 1] = pack_bf16x2(h[2], h[3]);
        ht[k4 / 2 + 0] = pack_bf16x2(dh[0], dh[1]);
        ht[k4 / 2 + 1] = pack_bf16x2(dh[2], dh[3]);
      }
      // 16-byte chunk (jj*4 + c8/8) of this row's 128-byte line in K panel hf, XOR-swizzled with (row & 7)
      const uint32_t off = ((uint32_t)((jj * 4 + c8 / 8) << 4)) ^ rx;
      *reinterpret_cast<uint4*>(hv_row + off) = make_uint4(hv[0], hv[1], hv[2], hv[3]);
      *reinterpret_cast<uint4*>(ht_row + off) = make_uint4(ht[0], ht[1], ht[2], ht[3]);
    }
    // this warp's 32 K-columns of the next A operand are written: publish them to the MMA thread
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(sbase + kOffBars + 8u * (kBarReady0 + hf * 2 + jj));
  }
}

template <int ITEMS, bool DET>
__global__ void __launch_bounds__(kTcThreads, 1)
cha_tc_kernel(const __grid_constant__ ProblemDev P, const __grid_constant__ TcConsts C,
              const __grid_constant__ TcArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = P.nx;
  const int UP = A.units_per_point, TP = A.tiles_per_point, M_I = A.M_I;
  auto bar = [&](int i) -> uint32_t { return sbase + kOffBars + 8u * (uint32_t)i; };

  if (threadIdx.x == 0) {
    if (sbase & 1023u) {
      printf("dpi tc: dynamic shared memory base 0x%x is not 1024-byte aligned\n", sbase);
      __trap();
    }
    mbar_init(bar(kBarW), 1);
    mbar_init(bar(kBarZFull), 8);   // the 8 producer warps
    mbar_init(bar(kBarZEmpty), 1);  // tcgen05.commit
    mbar_init(bar(kBarAcc0), 1);
    mbar_init(bar(kBarAcc1), 1);
    for (int c = 0; c < 4; ++c) mbar_init(bar(kBarReady0 + c), 4);  // 4 epilogue warps per K chunk
    mbar_init(bar(kBarCoef0), 4);
    mbar_init(bar(kBarCoef0 + 1), 4);
    fence_barrier_init();
  }
  // the padded K columns of the increment tile must read as zero
  for (int i = threadIdx.x; i < kOperandBytes / 16; i += kTcThreads)
    reinterpret_cast<uint4*>(smem + kOffZ)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  if (warp == kMmaWarp) tmem_alloc(sbase + kOffTmemPtr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr);

  if (warp >= kMmaWarp) {
    // =================================== MMA issuer (one thread) ===================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kMmaRegs));
    if (warp == kMmaWarp && lane == 0) {
      mbar_arrive_expect_tx(bar(kBarW), 4u * kOperandBytes);
      for (int i = 0; i < 16; ++i)
        bulk_g2s(sbase + kOffW + (uint32_t)i * 8192u, reinterpret_cast<const uint8_t*>(A.W) + (size_t)i * 8192, 8192u,
                 bar(kBarW));
      mbar_wait(bar(kBarW), 0, 0);
      constexpr uint32_t idesc = make_idesc_bf16(128, 128);
      // descriptors differ only in the 14-bit start-address field: +2 per UMMA_K step inside a K panel (32 B),
      // +1024 per panel (16 KB).  Kept as base + small offsets so the loop needs few registers (this thread runs
      // with 32 registers and there is next to no L1 left for spills).
      const uint64_t dZ = make_desc_sw128(sbase + kOffZ), dV = make_desc_sw128(sbase + kOffHV),
                     dT = make_desc_sw128(sbase + kOffHT), dW = make_desc_sw128(sbase + kOffW);
      uint32_t it = 0, ph_ready = 0;
      for (int64_t u = blockIdx.x; u < A.n_units; u += gridDim.x) {
        const int g = (int)(u % UP);
        const int tiles = min(kTcGroupTiles, TP - g * kTcGroupTiles);
        for (int tl = 0; tl < tiles; ++tl, ++it) {
          // layer 1: acc set 0 (value columns) = Z . W_x^T.  Set 0 is free: the chunks of layer 4 of the previous
          // tile were published only after every epilogue warp had finished reading set 0 (layer 3).
          mbar_wait(bar(kBarZFull), it & 1u, 1);
          tc_fence_after();
#pragma unroll 1
          for (int ks = 0; ks < A.ksteps1; ++ks) {
            const uint64_t o = (uint64_t)((ks >> 2) * 1024 + (ks & 3) * 2);
            umma_bf16(tmem, dZ + o, dW + o, idesc, ks > 0);
          }
          tc_commit(bar(kBarZEmpty));
          tc_commit(bar(kBarAcc0));
          // hidden layers 2..4: weights image l, accumulator set (l & 1); K chunks in the order 0,2,1,3 (chunk 0
          // first: its completion implies every epilogue warp is past the previous tile's output layer)
#pragma unroll 1
          for (int l = 1; l <= 3; ++l) {
            const uint32_t acc = tmem + (uint32_t)(l & 1) * 256u;
            const uint64_t dWl = dW + (uint64_t)(l * (kOperandBytes >> 4));
            uint32_t accum = 0;
#pragma unroll 1
            for (int ci = 0; ci < 4; ++ci) {
              const int c = (ci & 1) * 2 + (ci >> 1);
              mbar_wait(bar(kBarReady0 + c), ph_ready, 10 + c);
              tc_fence_after();
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const int ks = 2 * c + kk;
                const uint64_t o = (uint64_t)((ks >> 2) * 1024 + (ks & 3) * 2);
                umma_bf16(acc, dV + o, dWl + o, idesc, accum);
                umma_bf16(acc + 128u, dT + o, dWl + o, idesc, accum);
                accum = 1;
              }
            }
            ph_ready ^= 1u;
            tc_commit(bar(kBarAcc0 + (l & 1)));
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < 8) {
    // =================================== epilogue warps ===================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWorkerRegs));
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t rowbase = (uint32_t)hf * kPanelBytes + (uint32_t)r * 128u, rx = (uint32_t)(r & 7) << 4;
    float* c0_s = reinterpret_cast<float*>(smem + kOffC0);
    float* coef_s = reinterpret_cast<float*>(smem + kOffCoef);
    float2* xch = reinterpret_cast<float2*>(smem + kOffXch);
    uint32_t ph0 = 0, ph1 = 0, it = 0;
    for (int64_t u = blockIdx.x; u < A.n_units; u += gridDim.x) {
      const int64_t p = u / UP;
      const int g = (int)(u % UP);
      const int tiles = min(kTcGroupTiles, TP - g * kTcGroupTiles);
      const float t = __ldg(A.tx + p * (int64_t)(1 + d));
      const float dT = P.T - t;
      const float fbp = __ldg(A.fb + p);
      const uint64_t gp = (uint64_t)(A.point_base + p);
      named_bar_sync(1, 256);  // everyone is done with the previous point's c0
      if (threadIdx.x < 128) c0_s[threadIdx.x] = kLog2e * __ldg(A.c0 + p * 128 + threadIdx.x);
      named_bar_sync(1, 256);
      float fsum = 0.f;
      for (int tl = 0; tl < tiles; ++tl, ++it) {
        const int m = (g * kTcGroupTiles + tl) * 128 + r;
        const bool valid = m < M_I;
        float uu = 0.5f;
        if (valid) uu = DET ? __ldg(A.u_s + p * M_I + m) : u01(philox_at(A.seed, gp, kStreamS, (uint32_t)m, 0u).x);
        const float a = sqrtf(uu * dT) * P.sigma;
        const float s2 = (uu * dT + t) * kLog2e, a2 = a * kLog2e;

        mbar_wait(bar(kBarAcc0), ph0, 20);
        ph0 ^= 1u;
        tc_fence_after();
        epi_hidden<true>(smem, sbase, C, tlane, hf, lane, rowbase, rx, 0, s2, a2);
#pragma unroll 1
        for (int l = 0; l < 2; ++l) {  // hidden GEMM layers 2 and 3: accumulator sets 1 and 0
          if (l == 0) {
            mbar_wait(bar(kBarAcc1), ph1, 21);
            ph1 ^= 1u;
          } else {
            mbar_wait(bar(kBarAcc0), ph0, 22);
            ph0 ^= 1u;
          }
          tc_fence_after();
          epi_hidden<false>(smem, sbase, C, tlane + (l == 0 ? 256u : 0u), hf, lane, rowbase, rx, l, 0.f, 0.f);
        }

        // ---- output layer: u = w_out . h4 + b_out, D = w_out . h4' over this warp's 64 columns ----
        mbar_wait(bar(kBarAcc1), ph1, 23);
        ph1 ^= 1u;
        tc_fence_after();
        float pu = 0.f, pD = 0.f;
#pragma unroll 1
        for (int jj = 0; jj < 2; ++jj) {
          const int col0 = hf * 64 + jj * 32;
          uint32_t v[32], tv[32];
          tmem_ld32(tlane + 256u + col0, v);
          tmem_ld32(tlane + 256u + 128u + col0, tv);
          tmem_ld_wait();
#pragma unroll
          for (int c4 = 0; c4 < 32; c4 += 4) {
            const float4 b = *reinterpret_cast<const float4*>(&C.b[2][col0 + c4]);
            const float4 wo = *reinterpret_cast<const float4*>(&C.wout[col0 + c4]);
            const float bb[4] = {b.x, b.y, b.z, b.w}, ww[4] = {wo.x, wo.y, wo.z, wo.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float zb = __uint_as_float(v[c4 + i]) + bb[i];
              const float e = exp2f(fminf(zb, 0.f));
              pu = fmaf(ww[i], fmaxf(zb, fmaf(e, kLog2e, -kLog2e)), pu);
              pD = fmaf(ww[i], e * __uint_as_float(tv[c4 + i]), pD);
            }
          }
        }
        tc_fence_before();
        if (hf == 1) {
          xch[r] = make_float2(pu, pD);
          __threadfence_block();
          named_bar_arrive(2, 256);
        } else {
          named_bar_sync(2, 256);
          const float2 o = xch[r];
          const float uval = pu + o.x + C.bout, D = pD + o.y;
          const float f = P.sigma * (P.cha_k * uval - P.cha_c) * (P.sigma * D);
          if (valid) fsum += dT * f;
          coef_s[(it & 1u) * 128u + r] = valid ? dT * (f - fbp) / a : 0.f;
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(kBarCoef0 + (int)(it & 1u)));
        }
      }
      if (hf == 0) {
        const float ws = warp_sum(fsum);
        if (lane == 0) A.partF[u * 4 + q] = ws;
      }
    }
  } else {
    // =================================== increment producers ===================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWorkerRegs));
    // All 8 warps draw every tile: thread (q, c) owns the 4 dims [4q, 4q+4) of the rows r = c + cls * i.  A tile's
    // increments are generated into registers first (no dependence on the shared Z buffer), stored once the
    // layer-1 MMA of the previous tile has released the buffer, and kept until the tile's Bismut coefficients
    // arrive one tile later.
    const int tl_id = threadIdx.x - 256;
    const int nq = A.nq, cls = A.cls;
    const int q = tl_id % nq, c = tl_id / nq;
    const bool active = c < cls;
    const bool vec_ok = (d & 3) == 0;
    const float* coef_s = reinterpret_cast<const float*>(smem + kOffCoef);
    uint8_t* const zdst = smem + kOffZ + (uint32_t)(q >> 4) * kPanelBytes + (uint32_t)(q & 1) * 8u;
    const uint32_t zchunk = (uint32_t)((q & 15) >> 1);
    uint32_t zold[ITEMS][2], znew[ITEMS][2];
    float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t it = 0;
    bool prev_valid = false, prev_last = false;
    int64_t prev_u = 0;
    auto finish_prev = [&]() {
      // Bismut accumulation of tile it-1: G_j += coef_m Z_mj (bf16 Z, exactly what the MLP saw)
      const uint32_t pit = it - 1u;
      mbar_wait(bar(kBarCoef0 + (int)(pit & 1u)), (pit >> 1) & 1u, 31);
      if (active) {
        const float* cf_s = coef_s + (pit & 1u) * 128u;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
          const int r = c + cls * i;
          if (r < 128) {
            const float cf = cf_s[r];
            G.x = fmaf(cf, bf16_lo(zold[i][0]), G.x);
            G.y = fmaf(cf, bf16_hi(zold[i][0]), G.y);
            G.z = fmaf(cf, bf16_lo(zold[i][1]), G.z);
            G.w = fmaf(cf, bf16_hi(zold[i][1]), G.w);
          }
        }
        if (prev_last) {
          *reinterpret_cast<float4*>(A.partG + (prev_u * 16 + c) * 128 + 4 * q) = G;
          G = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    for (int64_t u = blockIdx.x; u < A.n_units; u += gridDim.x) {
      const int64_t p = u / UP;
      const int g = (int)(u % UP);
      const int tiles = min(kTcGroupTiles, TP - g * kTcGroupTiles);
      const uint64_t gp = (uint64_t)(A.point_base + p);
      for (int tl = 0; tl < tiles; ++tl) {
        const int m0 = (g * kTcGroupTiles + tl) * 128;
        if (active) {
#pragma unroll
          for (int i = 0; i < ITEMS; ++i) {
            const int r = c + cls * i;
            const int m = m0 + r;
            float4 zz = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < 128 && m < M_I) {
              if (DET) {
                zz = load_row4(A.zS + ((int64_t)p * M_I + m) * d, 4 * q, d, vec_ok);
              } else {
                zz = normal4(A.seed, gp, kStreamIntegral, (uint32_t)m, (uint32_t)q);
                if (4 * q + 1 >= d) zz.y = 0.f;
                if (4 * q + 2 >= d) zz.z = 0.f;
                if (4 * q + 3 >= d) zz.w = 0.f;
              }
            }
            znew[i][0] = pack_bf16x2(zz.x, zz.y);
            znew[i][1] = pack_bf16x2(zz.z, zz.w);
          }
        }
        if (it > 0) mbar_wait(bar(kBarZEmpty), (it - 1u) & 1u, 30);  // layer-1 MMA of the previous tile has read Z
        if (active) {
#pragma unroll
          for (int i = 0; i < ITEMS; ++i) {
            const int r = c + cls * i;
            if (r < 128)
              *reinterpret_cast<uint2*>(zdst + (uint32_t)r * 128u + ((zchunk ^ (uint32_t)(r & 7)) << 4)) =
                  make_uint2(znew[i][0], znew[i][1]);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kBarZFull));
        if (prev_valid) finish_prev();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) zold[i][0] = znew[i][0], zold[i][1] = znew[i][1];
        prev_u = u;
        prev_last = (tl == tiles - 1);
        prev_valid = true;
        ++it;
      }
    }
    if (prev_valid) finish_prev();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// y[p] += (sum of the unit partials in a fixed order) / M_I, then clip (data.py:222)
__global__ void __launch_bounds__(128)
cha_tc_finish_kernel(const float* __restrict__ partG, const float* __restrict__ partF, int units_per_point, int cls, int d,
                     int M_I, float bound, float* __restrict__ y) {
  const int64_t p = blockIdx.x;
  const int j = threadIdx.x;
  const float invM = 1.f / (float)M_I;
  const bool clip = bound > 0.f && isfinite(bound);
  float* yp = y + p * (int64_t)(1 + d);
  if (j < d) {
    float s = 0.f;
    for (int g = 0; g < units_per_point; ++g)
      for (int c = 0; c < cls; ++c) s += partG[((p * units_per_point + g) * 16 + c) * 128 + j];
    float v = yp[1 + j] + s * invM;
    if (clip) v = fminf(fmaxf(v, -bound), bound);
    yp[1 + j] = v;
  }
  if (j == 0) {
    float s = 0.f;
    for (int g = 0; g < units_per_point; ++g)
      for (int qq = 0; qq < 4; ++qq) s += partF[(p * units_per_point + g) * 4 + qq];
    float v = yp[0] + s * invM;
    if (clip) v = fminf(fmaxf(v, -bound), bound);
    yp[0] = v;
  }
}

// dst = bf16 SWIZZLE_128B K-major image of W[n][col_off + k] (n < n_valid, k < k_valid; zero elsewhere)
__global__ void pack_sw128_kernel(const float* __restrict__ W, int ld, int col_off, int n_valid, int k_valid,
                                  uint8_t* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 128) return;
  const int n = i >> 7, k = i & 127;
  const float v = (n < n_valid && k < k_valid) ? W[(size_t)n * ld + col_off + k] : 0.f;
  *reinterpret_cast<__nv_bfloat16*>(dst + sw128_offset((uint32_t)n, (uint32_t)k)) = __float2bfloat16_rn(v);
}

// ---- UMMA self-test: D[128x128] = A[128x128] . B[128x128]^T through the same descriptors / layouts -------------
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const uint8_t* __restrict__ Aimg,
                                                               const uint8_t* __restrict__ Bimg, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t barW = sbase + 2 * kOperandBytes, barD = barW + 8, tptr = barW + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    if (sbase & 1023u) __trap();
    mbar_init(barW, 1);
    mbar_init(barD, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tptr, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 2 * kOperandBytes + 16);
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(barW, 2u * kOperandBytes);
    bulk_g2s(sbase, Aimg, kOperandBytes, barW);
    bulk_g2s(sbase + kOperandBytes, Bimg, kOperandBytes, barW);
    mbar_wait(barW, 0, 100);
    tc_fence_after();
    for (int ks = 0; ks < 8; ++ks)
      umma_bf16(tmem, desc_kstep(sbase, ks), desc_kstep(sbase + kOperandBytes, ks), make_idesc_bf16(128, 128), ks > 0);
    tc_commit(barD);
  }
  __syncwarp();
  mbar_wait(barD, 0, 101);
  tc_fence_after();
  const int r = warp * 32 + lane;
  for (int col0 = 0; col0 < 128; col0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + col0, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[r * 128 + col0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

template <int ITEMS, bool DET>
int launch_tc_t(const ProblemDev& P, const TcNetDev& N, const TcArgs& A, int grid, cudaStream_t st) {
  auto k = cha_tc_kernel<ITEMS, DET>;
  DPI_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
  k<<<grid, kTcThreads, kTcSmemBytes, st>>>(P, N.consts, A);
  DPI_CUDA_CHECK(cudaGetLastError());
  return DPI_OK;
}

}  // namespace

int tc_pack_network(const dpi_network& net, const ProblemDev& P, const MlpDev& M, std::vector<void*>& owner,
                    TcNetDev* out, bool* ready, int64_t* launches, cudaStream_t st) {
  *ready = false;
  const int L = net.n_main;
  if (L != 5 || P.nx > 128) return DPI_OK;  // this kernel: Linear(1+d,128), 3 x Linear(128,128), Linear(128,1)
  for (int l = 0; l < L - 1; ++l)
    if (net.main[l].n_out != 128) return DPI_OK;
  void* img = nullptr;
  DPI_CUDA_CHECK(cudaMalloc(&img, 4 * (size_t)tc::kOperandBytes));
  owner.push_back(img);
  uint8_t* base = reinterpret_cast<uint8_t*>(img);
  pack_sw128_kernel<<<64, 256, 0, st>>>(net.main[0].W, 1 + P.nx, 1, 128, P.nx, base);
  for (int l = 1; l <= 3; ++l)
    pack_sw128_kernel<<<64, 256, 0, st>>>(net.main[l].W, 128, 0, 128, 128, base + (size_t)l * tc::kOperandBytes);
  DPI_CUDA_CHECK(cudaGetLastError());
  *launches += 4;
  // per-column fp32 constants travel as kernel parameters (constant bank): copy the packed vectors back once
  DPI_CUDA_CHECK(cudaStreamSynchronize(st));
  TcNetDev T{};
  DPI_CUDA_CHECK(cudaMemcpy(T.consts.wt, M.wt, 128 * sizeof(float), cudaMemcpyDeviceToHost));
  DPI_CUDA_CHECK(cudaMemcpy(T.consts.wsum, M.wsum, 128 * sizeof(float), cudaMemcpyDeviceToHost));
  for (int l = 0; l < 3; ++l)
    DPI_CUDA_CHECK(cudaMemcpy(T.consts.b[l], M.bh[l], 128 * sizeof(float), cudaMemcpyDeviceToHost));
  DPI_CUDA_CHECK(cudaMemcpy(T.consts.wout, M.wout, 128 * sizeof(float), cudaMemcpyDeviceToHost));
  DPI_CUDA_CHECK(cudaMemcpy(&T.consts.bout, M.bout, sizeof(float), cudaMemcpyDeviceToHost));
  // activations are carried scaled by log2(e) (see epi_hidden): scale the vectors that add to / read from them
  for (int i = 0; i < 128; ++i) {
    T.consts.wt[i] *= kLog2e;
    T.consts.wsum[i] *= kLog2e;
    for (int l = 0; l < 3; ++l) T.consts.b[l][i] *= kLog2e;
    T.consts.wout[i] *= 0.6931471805599453f;
  }
  T.W = img;
  T.nx = P.nx;
  T.ksteps1 = (P.nx + 15) / 16;
  T.nq = (P.nx + 3) / 4;
  T.cls = T.nq <= 16 ? 16 : (T.nq <= 25 ? 10 : 8);  // row classes of the 256 producer threads: cls * nq <= 256
  int dev = 0, sms = 0;
  DPI_CUDA_CHECK(cudaGetDevice(&dev));
  DPI_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  T.num_sms = sms;
  *out = T;
  *ready = true;
  return DPI_OK;
}

int64_t tc_units_per_point(int M_I) {
  const int tiles = (M_I + 127) / 128;
  return (tiles + kTcGroupTiles - 1) / kTcGroupTiles;
}

int launch_cha_tc_path(const ProblemDev& P, const TcNetDev& N, const float* tx, int64_t n, int M_I, uint64_t seed,
                       int64_t base, const float* u_s, const float* zS, const float* fb, const float* c0, float bound,
                       float* partG, float* partF, float* y, cudaStream_t st) {
  if (n <= 0) return DPI_OK;
  TcArgs A{};
  A.W = N.W;
  A.tx = tx;
  A.tiles_per_point = (M_I + 127) / 128;
  A.units_per_point = (int)tc_units_per_point(M_I);
  A.n_units = n * A.units_per_point;
  A.point_base = base;
  A.seed = seed;
  A.u_s = u_s;
  A.zS = zS;
  A.fb = fb;
  A.c0 = c0;
  A.partG = partG;
  A.partF = partF;
  A.M_I = M_I;
  A.ksteps1 = N.ksteps1;
  A.nq = N.nq;
  A.cls = N.cls;
  const int grid = (int)(A.n_units < N.num_sms ? A.n_units : N.num_sms);
  int rc;
  const bool det = zS != nullptr;
  if (N.cls == 16)
    rc = det ? launch_tc_t<8, true>(P, N, A, grid, st) : launch_tc_t<8, false>(P, N, A, grid, st);
  else if (N.cls == 10)
    rc = det ? launch_tc_t<13, true>(P, N, A, grid, st) : launch_tc_t<13, false>(P, N, A, grid, st);
  else
    rc = det ? launch_tc_t<16, true>(P, N, A, grid, st) : launch_tc_t<16, false>(P, N, A, grid, st);
  if (rc) return rc;
  cha_tc_finish_kernel<<<(unsigned)n, 128, 0, st>>>(partG, partF, A.units_per_point, N.cls, P.nx, M_I, bound, y);
  DPI_CUDA_CHECK(cudaGetLastError());
  return DPI_OK;
}

int tc_umma_selftest(const float* A, const float* B, float* D, cudaStream_t st) {
  uint8_t* img = nullptr;
  DPI_CUDA_CHECK(cudaMalloc((void**)&img, 2 * (size_t)tc::kOperandBytes));
  pack_sw128_kernel<<<64, 256, 0, st>>>(A, 128, 0, 128, 128, img);
  pack_sw128_kernel<<<64, 256, 0, st>>>(B, 128, 0, 128, 128, img + tc::kOperandBytes);
  const int smem = 2 * tc::kOperandBytes + 64;
  cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) {
    umma_selftest_kernel<<<1, 128, smem, st>>>(img, img + tc::kOperandBytes, D);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(img);
  if (e != cudaSuccess) {
    set_error("umma selftest failed: %s", cudaGetErrorString(e));
    return DPI_ERR_CUDA;
  }
  return DPI_OK;
}

}  // namespace dpi
