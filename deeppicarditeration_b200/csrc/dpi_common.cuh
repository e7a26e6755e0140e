// Shared device/host definitions for the DPI label-generator kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dpi_b200.h"

namespace dpi {

constexpr int kMaxGbmNodes = 4;
constexpr int kMaxOuK = 8;
constexpr float kTwoPi = 6.283185307179586f;

// Philox stream ids (word 3 of the counter): one independent stream per reference draw site.
enum Stream : uint32_t {
  kStreamT = 0,         // data.py:166     rand[n,1]
  kStreamX0 = 1,        // utils.py:785    randn[n,nx]   (OU only)
  kStreamX = 2,         // equations.py:225 randn[n,nx]
  kStreamTerminal = 3,  // data.py:914     randn[n*M_T,nx]
  kStreamS = 4,         // data.py:359     rand[n*M_I,1]
  kStreamIntegral = 5,  // data.py:360     randn[n*M_I,nx]
  kStreamIdx = 6        // data.py:501     randint[n*M_I,v]
};

// Device-side view of the equation (small arrays live in handle-owned device memory).
struct ProblemDev {
  int eq;
  int nx;
  float T, alpha, sigma;  // sigma = sqrt(alpha)
  // Cha: k = k_cfg/sqrt(nx); c = (2+k^2 alpha d)/(2 k alpha d)   (equations.py:283-302)
  float cha_k, cha_c;
  // GBM: w [nodes][1+nx], v [nodes], w2v[nodes] = v_k * sum_j w_kj^2 (for the Laplacian of u*)
  int gbm_nodes;
  const float* gbm_w;
  float gbm_v[kMaxGbmNodes];
  // OU: mean [K][nx], ivar [K][nx] = 1/var, lognorm[k] = log pi_k - 0.5 (d log 2pi + log det Sigma_k)
  int ou_K;
  const float* ou_mean;
  const float* ou_ivar;
  float ou_lognorm[kMaxOuK];
  float ou_theta, ou_mu, ou_x0_std;
  int sdgd_v;
  // DATA.ESTIMATE_DELTA_T (data.py:529-575, :928-952): > 0 replaces the horizon T - t of both estimators by
  // min(t + td_delta, T) - t; set per call from dpi_label_config
  float td_delta;
};

// length of the time interval the estimators of a point at time t integrate over
__host__ __device__ __forceinline__ float horizon(const ProblemDev& P, float t) {
  return P.td_delta > 0.f ? fminf(t + P.td_delta, P.T) - t : P.T - t;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: every (point, stream, sample, dim-block) has its own
// counter, so labels are independent of launch geometry, shard size and GPU count.
// ---------------------------------------------------------------------------------------------
// One round: two 32x32 -> 64-bit products (one IMAD.WIDE each) and two three-input XORs.
__device__ __forceinline__ uint4 philox_round(uint4 c, uint32_t kx, uint32_t ky) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
  return make_uint4((uint32_t)(p1 >> 32) ^ c.y ^ kx, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c.w ^ ky, (uint32_t)p0);
}
// Rounds of the bijection.  10 is the Random123 / cuRAND default; DPI_PHILOX_ROUNDS exists for sensitivity builds
// (tools/variants.py), nothing shipped changes it.
#ifndef DPI_PHILOX_ROUNDS
#define DPI_PHILOX_ROUNDS 10
#endif
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < DPI_PHILOX_ROUNDS; ++r) {
    c = philox_round(c, k.x, k.y);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// The ten round keys of a seed, computed once on the host.  Passed as a kernel parameter they sit in the constant
// bank and feed the XORs as instruction operands: the generator-bound kernels save the 20 key additions per block.
struct PhiloxKeys {
  uint32_t kx[10], ky[10];
};
inline PhiloxKeys make_philox_keys(uint64_t seed) {
  PhiloxKeys K;
  uint32_t kx = (uint32_t)seed, ky = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    K.kx[r] = kx, K.ky[r] = ky;
    kx += 0x9E3779B9u, ky += 0xBB67AE85u;
  }
  return K;
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys& K) {
#pragma unroll
  for (int r = 0; r < DPI_PHILOX_ROUNDS; ++r) c = philox_round(c, K.kx[r], K.ky[r]);
  return c;
}

__device__ __forceinline__ uint4 philox_counter(uint64_t point, uint32_t stream, uint32_t sample, uint32_t block) {
  // counter = {dim block, sample, point[31:0], point[47:32] | stream<<16}
  return make_uint4(block, sample, (uint32_t)point, ((uint32_t)(point >> 32) & 0xFFFFu) | (stream << 16));
}
__device__ __forceinline__ uint4 philox_at(uint64_t seed, uint64_t point, uint32_t stream, uint32_t sample,
                                           uint32_t block) {
  return philox4x32_10(philox_counter(point, stream, sample, block), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
__device__ __forceinline__ uint4 philox_at(const PhiloxKeys& K, uint64_t point, uint32_t stream, uint32_t sample,
                                           uint32_t block) {
  return philox4x32_10(philox_counter(point, stream, sample, block), K);
}

// uniform in the OPEN interval (0,1): (x>>8 + 0.5) * 2^-24
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * 5.9604644775390625e-8f; }

// Box-Muller on the MUFU pipe, 16 + 16 random bits -> 2 normals (13 instructions): radius from
// u = (k + 1/2) 2^-16 in (0, 1), angle 2 pi (k' + 1/2) 2^-16.  One Philox4x32 block therefore yields EIGHT normals.
// The pair (z0, z1) takes 2^32 distinct values; the radius is a midpoint rule with 65536 equiprobable cells
// (|z| <= 4.86: the missing tail has probability 1.2e-6), the angle an equispaced rule with 65536 nodes, which is
// spectrally accurate for the smooth periodic integrands it feeds.  The midpoint rule alone leaves E[r^2] at
// 1.9999894 instead of 2 (a SYSTEMATIC 5.3e-6 deficit in the variance of every increment, invisible to
// deterministic-mode parity), so the radius constant is calibrated: kBmRadius = -2 ln2 * 2 / mean_k(-2 ln u_k)
// makes E[z^2] = 1 exactly over the 2^32 pairs (E[z^4] = 2.99982); tests/test_gpu_rng.py bounds the effect on a
// label against Gauss-Hermite quadrature.
constexpr float kBmRadius = -1.3863016922764049f;
#ifdef DPI_BM_LOP3
constexpr bool kBmLop3Default = true;
#else
constexpr bool kBmLop3Default = false;
#endif
template <bool LOP3 = kBmLop3Default>
__device__ __forceinline__ void box_muller16(uint32_t w, float& z0, float& z1) {
  // 16-bit integer -> float through the mantissa (1 + k 2^-16 in [1, 2)): shift, mask|or, add -- no int-to-float
  // conversion instruction, which would share the quarter-rate XU pipe with the four MUFU ops below
  float xr, xa;
  if constexpr (LOP3) {
    // (x & mask) | one in ONE LOP3: both constants have to sit in registers for that (an instruction takes a single
    // immediate), so they come through an opaque asm the compiler cannot fold back into immediates.  Same bits as
    // the plain form; worth 2 % where the generator bounds the kernel and two registers are to spare.
    uint32_t kmask, kone;
    asm("mov.b32 %0, 0x007FFF80;" : "=r"(kmask));
    asm("mov.b32 %0, 0x3F800000;" : "=r"(kone));
    uint32_t br, ba;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(br) : "r"(w << 7), "r"(kmask), "r"(kone));
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(ba) : "r"(w >> 9), "r"(kmask), "r"(kone));
    xr = __uint_as_float(br), xa = __uint_as_float(ba);
  } else {
    xr = __uint_as_float(((w << 7) & 0x007FFF80u) | 0x3F800000u);
    xa = __uint_as_float(((w >> 9) & 0x007FFF80u) | 0x3F800000u);
  }
  const float u = xr - 0.99999237060546875f;                 // (k + 1/2) 2^-16
#ifdef DPI_DBG_BM_NOXU  // sensitivity builds only: same instruction count without the four MUFU ops (wrong numbers)
  const float r = fmaf(kBmRadius, fmaf(u, 0.5f, 0.25f), 2.f) * 0.7f;
  const float ang = fmaf(xa, 6.283185307179586f, -6.283137370279964f) * 0.15915494f;
  float s = fmaf(ang, 0.5f, -0.25f), c = fmaf(ang, -0.5f, 0.25f);
#else
  const float r = sqrtf(kBmRadius * __log2f(u));  // sqrt(-2 ln u) = sqrt(-2 ln2 log2 u), calibrated (see above)
  float s, c;
  __sincosf(fmaf(xa, 6.283185307179586f, -6.283137370279964f), &s, &c);  // 2 pi (k' + 1/2) 2^-16
#endif
  z0 = r * c;
  z1 = r * s;
}

// normals of dims [8 b, 8 b + 8) of (point, stream, sample): za = first four, zb = last four
__device__ __forceinline__ void normal8(uint64_t seed, uint64_t point, uint32_t stream, uint32_t sample, uint32_t block8,
                                        float4& za, float4& zb) {
  const uint4 r = philox_at(seed, point, stream, sample, block8);
  box_muller16(r.x, za.x, za.y);
  box_muller16(r.y, za.z, za.w);
  box_muller16(r.z, zb.x, zb.y);
  box_muller16(r.w, zb.z, zb.w);
}

template <bool LOP3 = kBmLop3Default>
__device__ __forceinline__ void normal8(const PhiloxKeys& K, uint64_t point, uint32_t stream, uint32_t sample,
                                        uint32_t block8, float4& za, float4& zb) {
  const uint4 r = philox_at(K, point, stream, sample, block8);
  box_muller16<LOP3>(r.x, za.x, za.y);
  box_muller16<LOP3>(r.y, za.z, za.w);
  box_muller16<LOP3>(r.z, zb.x, zb.y);
  box_muller16<LOP3>(r.w, zb.z, zb.w);
}

// Two blocks at once -- same (point, stream, dim block), samples m0 and m1 -- with the Philox rounds of the two
// counters written in lock-step.  One block is a chain of ten dependent rounds (2 wide multiplies + 2 XORs each, ~10
// cycles per round); a producer role with two warps per scheduler cannot hide that chain, and the compiler does not
// interleave separate calls on its own (it serialises them to save registers).  Bit-identical to two normal8 calls.
template <bool LOP3 = kBmLop3Default>
__device__ __forceinline__ void normal8x2(const PhiloxKeys& K, uint64_t point, uint32_t stream, uint32_t m0, uint32_t m1,
                                          uint32_t block8, float4& za0, float4& zb0, float4& za1, float4& zb1) {
  uint4 a = philox_counter(point, stream, m0, block8), b = philox_counter(point, stream, m1, block8);
#pragma unroll
  for (int r = 0; r < DPI_PHILOX_ROUNDS; ++r) {
    a = philox_round(a, K.kx[r], K.ky[r]);
    b = philox_round(b, K.kx[r], K.ky[r]);
  }
  box_muller16<LOP3>(a.x, za0.x, za0.y);
  box_muller16<LOP3>(b.x, za1.x, za1.y);
  box_muller16<LOP3>(a.y, za0.z, za0.w);
  box_muller16<LOP3>(b.y, za1.z, za1.w);
  box_muller16<LOP3>(a.z, zb0.x, zb0.y);
  box_muller16<LOP3>(b.z, zb1.x, zb1.y);
  box_muller16<LOP3>(a.w, zb0.z, zb0.w);
  box_muller16<LOP3>(b.w, zb1.z, zb1.w);
}

// normals of dims [4 q, 4 q + 4): half of a normal8 block.  For the kernels that are not bound by the generator and
// keep a lane <-> 4 dims layout (they draw the block twice); the terminal and tensor-core kernels consume all eight.
__device__ __forceinline__ float4 normal4(uint64_t seed, uint64_t point, uint32_t stream, uint32_t sample,
                                          uint32_t quad) {
  const uint4 r = philox_at(seed, point, stream, sample, quad >> 1);
  float4 z;
  box_muller16((quad & 1u) ? r.z : r.x, z.x, z.y);
  box_muller16((quad & 1u) ? r.w : r.y, z.z, z.w);
  return z;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// ELU and its derivative (alpha = 1): h = z>0 ? z : e^z-1 ; h' = z>0 ? 1 : e^z
__device__ __forceinline__ void elu_d1(float z, float& h, float& d1) {
  const float e = __expf(fminf(z, 0.f));
  const bool pos = z > 0.f;
  h = pos ? z : e - 1.f;
  d1 = pos ? 1.f : e;
}

// Load 4 consecutive dims [j0, j0+4) of a row of a [rows, d] fp32 matrix; zero beyond d.
__device__ __forceinline__ float4 load_row4(const float* __restrict__ row, int j0, int d, bool vec_ok) {
  float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (vec_ok && j0 + 3 < d) {
    z = __ldg(reinterpret_cast<const float4*>(row + j0));
  } else {
    if (j0 + 0 < d) z.x = __ldg(row + j0 + 0);
    if (j0 + 1 < d) z.y = __ldg(row + j0 + 1);
    if (j0 + 2 < d) z.z = __ldg(row + j0 + 2);
    if (j0 + 3 < d) z.w = __ldg(row + j0 + 3);
  }
  return z;
}

}  // namespace dpi

// ---------------------------------------------------------------------------------------------
// host-side error plumbing (definitions in dpi_api.cu)
// ---------------------------------------------------------------------------------------------
namespace dpi {
void set_error(const char* fmt, ...);
}
#define DPI_CUDA_CHECK(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      dpi::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DPI_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)
