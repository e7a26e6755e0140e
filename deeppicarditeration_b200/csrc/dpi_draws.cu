// Materialise the Philox draws of a label-generation call in the reference's own layout (SURVEY.md 3.2:
// zT [n*M_T, nx], u_s [n*M_I], zS [n*M_I, nx], idx [n*M_I, v]; row = point * M + sample), and a finiteness check of a
// finished label block.
//
// The label kernels never store their increments; this kernel re-derives them from the same counters
// (dpi_common.cuh: one Philox4x32-10 block per (point, stream, sample, octet of dims) -> eight normals; one block per
// four SDGD coordinates), so that a caller can (a) feed the very draws of a Philox-mode call to the reference /
// the CPU oracle, (b) replay them through the deterministic mode, (c) test the generator's distribution.
// HBM-bound streaming writes: one 16-byte store per four values, coalesced along the row.
#include "dpi_common.cuh"
#include "dpi_kernels.h"

namespace dpi {
namespace {

// one thread per (row, octet of dims): out[row][8 o .. 8 o + 8) = normals of Philox block o of (point, stream, sample)
__global__ void __launch_bounds__(256)
export_normals_kernel(const PhiloxKeys keys, int64_t point_base, uint32_t stream, int64_t rows, int M, int d,
                      float* __restrict__ out) {
  const int no = (d + 7) >> 3;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * no) return;
  const int64_t row = i / no;
  const int o = (int)(i - row * no);
  const int64_t p = row / M;
  const uint32_t m = (uint32_t)(row - p * M);
  float4 za, zb;
  normal8(keys, (uint64_t)(point_base + p), stream, m, (uint32_t)o, za, zb);
  float* dst = out + row * (int64_t)d + 8 * o;
  const float v[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
  if ((d & 3) == 0 && 8 * o + 7 < d) {
    *reinterpret_cast<float4*>(dst) = za;
    *reinterpret_cast<float4*>(dst + 4) = zb;
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (8 * o + q < d) dst[q] = v[q];
  }
}

__global__ void __launch_bounds__(256)
export_times_kernel(const PhiloxKeys keys, int64_t point_base, int64_t rows, int M, float* __restrict__ u_s) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const int64_t p = row / M;
  u_s[row] = u01(philox_at(keys, (uint64_t)(point_base + p), kStreamS, (uint32_t)(row - p * M), 0u).x);
}

// SDGD coordinates (data.py:497-502): v draws with replacement from {0..d-1}; Philox block b holds coordinates 4b..4b+3
__global__ void __launch_bounds__(256)
export_idx_kernel(const PhiloxKeys keys, int64_t point_base, int64_t rows, int M, int d, int v, int32_t* __restrict__ idx) {
  const int nb = (v + 3) >> 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * nb) return;
  const int64_t row = i / nb;
  const int b = (int)(i - row * nb);
  const int64_t p = row / M;
  const uint4 r = philox_at(keys, (uint64_t)(point_base + p), kStreamIdx, (uint32_t)(row - p * M), (uint32_t)b);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (4 * b + q < v) idx[row * (int64_t)v + 4 * b + q] = (int32_t)__umulhi(w[q], (uint32_t)d);
}

// flag |= 1 if any of the count values is not finite (a 16-bit operand overflow upstream shows up here as inf / NaN)
__global__ void __launch_bounds__(256) check_finite_kernel(const float* __restrict__ y, int64_t count, int* __restrict__ flag) {
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    bad |= !isfinite(y[i]);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

unsigned blocks_for(int64_t items) { return (unsigned)((items + 255) / 256); }

}  // namespace

int launch_export_draws(const ProblemDev& P, int64_t n, int M_T, int M_I, uint64_t seed, int64_t base, float* zT,
                        float* u_s, float* zS, int32_t* idx, cudaStream_t st, int64_t* launches) {
  if (n <= 0) return DPI_OK;
  const PhiloxKeys keys = make_philox_keys(seed);
  const int d = P.nx, no = (d + 7) >> 3;
  if (zT) {
    export_normals_kernel<<<blocks_for(n * M_T * no), 256, 0, st>>>(keys, base, kStreamTerminal, n * M_T, M_T, d, zT);
    ++*launches;
  }
  if (zS) {
    export_normals_kernel<<<blocks_for(n * M_I * no), 256, 0, st>>>(keys, base, kStreamIntegral, n * M_I, M_I, d, zS);
    ++*launches;
  }
  if (u_s) {
    export_times_kernel<<<blocks_for(n * M_I), 256, 0, st>>>(keys, base, n * M_I, M_I, u_s);
    ++*launches;
  }
  if (idx) {
    if (P.sdgd_v < 1) {
      set_error("dpi_export_draws: idx requested but the problem has no SDGD approximation (sdgd_v = 0)");
      return DPI_ERR_INVALID;
    }
    export_idx_kernel<<<blocks_for(n * M_I * ((P.sdgd_v + 3) >> 2)), 256, 0, st>>>(keys, base, n * M_I, M_I, d, P.sdgd_v, idx);
    ++*launches;
  }
  DPI_CUDA_CHECK(cudaGetLastError());
  return DPI_OK;
}

int launch_check_finite(const float* y, int64_t count, int* flag, cudaStream_t st) {
  if (count <= 0) return DPI_OK;
  const int64_t want = (count + 255) / 256;
  check_finite_kernel<<<(unsigned)(want < 1184 ? want : 1184), 256, 0, st>>>(y, count, flag);
  DPI_CUDA_CHECK(cudaGetLastError());
  return DPI_OK;
}

}  // namespace dpi
