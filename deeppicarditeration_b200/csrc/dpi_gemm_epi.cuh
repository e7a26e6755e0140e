// Fused GEMM epilogues (GemmEpi in dpi_kernels.h), shared by the fp32 register-tiled GEMMs (dpi_hjb.cu) and the
// TF32 tcgen05 GEMM (dpi_gemm_tc.cu).
#pragma once
#include "dpi_common.cuh"
#include "dpi_kernels.h"

namespace dpi {

// epilogue of 4 consecutive columns [col, col+4) of one output row
__device__ __forceinline__ void gemm_epilogue4(const GemmArgs& g, int64_t row, int col, float (&v)[4], bool c_vec,
                                               bool d_vec, bool e_vec) {
  const int N = g.N;
  float dv[4] = {0.f, 0.f, 0.f, 0.f};
  const bool full = col + 3 < N;
  if (g.epi == kEpiBias || g.epi == kEpiBiasElu || g.epi == kEpiBiasEluD) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (col + j < N) v[j] += __ldg(g.bias + col + j);
  }
  if (g.epi == kEpiBiasElu || g.epi == kEpiBiasEluD) {
#pragma unroll
    for (int j = 0; j < 4; ++j) elu_d1(v[j], v[j], dv[j]);
  }
  if (g.epi == kEpiMulE) {
    const float* ep = g.E + row * (int64_t)g.lde + col;
    if (full && e_vec) {
      const float4 e = __ldg(reinterpret_cast<const float4*>(ep));
      v[0] *= e.x, v[1] *= e.y, v[2] *= e.z, v[3] *= e.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col + j < N) v[j] *= __ldg(ep + j);
    }
  }
  if (g.epi == kEpiMulDElu) {  // the forward pass kept h = ELU(z) only: e^z = h + 1 where h <= 0
    const float* ep = g.E + row * (int64_t)g.lde + col;
    float h[4] = {1.f, 1.f, 1.f, 1.f};
    if (full && e_vec) {
      const float4 e = __ldg(reinterpret_cast<const float4*>(ep));
      h[0] = e.x, h[1] = e.y, h[2] = e.z, h[3] = e.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col + j < N) h[j] = __ldg(ep + j);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] *= (h[j] > 0.f) ? 1.f : h[j] + 1.f;
  }
  if (g.epi == kEpiKeepMulE) {  // C = acc (pre-activation tangent), D = acc * E (tangent through the ELU)
    const float* ep = g.E + row * (int64_t)g.lde + col;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (col + j < N) dv[j] = v[j] * __ldg(ep + j);
  }
  float* cp = g.C + row * (int64_t)g.ldc + col;
  if (full && c_vec) {
    *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (col + j < N) cp[j] = v[j];
  }
  if (g.epi == kEpiBiasEluD || g.epi == kEpiKeepMulE) {
    float* dp = g.D + row * (int64_t)g.ldd + col;
    if (full && d_vec) {
      *reinterpret_cast<float4*>(dp) = make_float4(dv[0], dv[1], dv[2], dv[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col + j < N) dp[j] = dv[j];
    }
  }
}

}  // namespace dpi
