// Collocation-point sampling and the terminal estimator E[(g(X_T)-g(x)) (1, Z/(sqrt(T-t) sigma))] + (g(x), 0).
//
// Restates (not ports) picard/data.py:161-167 (t), picard/equations.py:118-119,217-230 (x) and
// picard/data.py:899-926 (terminal estimator with Bismut weights).  The reference materialises five
// [n*M_T, nx] fp64 tensors in HBM; here a half-warp owns one MC sample at a time, each lane generates its 8 Brownian
// increments per 128-dim slot (one Philox block) in registers (or streams them with one coalesced 16-byte load in
// deterministic mode), the payoff statistics are warp-shuffle reduced, and the Bismut-weighted sums stay in
// registers until one label row per point is written.  HBM traffic: (1+nx) floats in, (1+nx) floats out per point.
// single-LOP3 mantissa assembly in Box-Muller: 2 % faster here; in the tensor-core kernel the two extra registers of
// the constants push its producers into spills, so it is a per-translation-unit switch
#define DPI_BM_LOP3
#include "dpi_common.cuh"
#include "dpi_kernels.h"

namespace dpi {

namespace {

constexpr int kTermThreads = 256;
constexpr int kTermWarps = kTermThreads / 32;

// ---- per-equation payoff g from per-row sufficient statistics ------------------------------------
// Cha:  1 statistic  S   = sum_j Z_j                       g = sigmoid(T + k (sum x + a S))
// GBM:  nodes stats  P_k = sum_j w_kj Z_j                  g = sum_k v_k sin(w_k0 T + w_k.x + a P_k)
// OU :  K stats      Q_k = sum_j (x_j - m_kj + a Z_j)^2/var_kj   g = -logsumexp_k(lognorm_k - Q_k/2)
template <int EQ>
struct Payoff;

template <>
struct Payoff<DPI_EQ_CHA> {
  static constexpr int NS = 1;
  static constexpr int NCH = 1;  // per-dim constant channels in shared memory: 1 for j < d, 0 for the padded dims
  __device__ static void accumulate(float (&st)[NS], const float4 z, const float* cst, int ldc, int j0, float a, int) {
    const float4 w = *reinterpret_cast<const float4*>(cst + j0);
    st[0] += (w.x * z.x + w.y * z.y) + (w.z * z.z + w.w * z.w);
  }
  __device__ static float eval(const ProblemDev& P, const float (&st)[NS], const float* base, float a) {
    return sigmoidf_(P.T + P.cha_k * (base[0] + a * st[0]));
  }
};

template <>
struct Payoff<DPI_EQ_GBM> {
  static constexpr int NS = kMaxGbmNodes;
  static constexpr int NCH = kMaxGbmNodes;  // w_k[j]
  __device__ static void accumulate(float (&st)[NS], const float4 z, const float* cst, int ldc, int j0, float a, int) {
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(cst + k * ldc + j0);
      st[k] += (w.x * z.x + w.y * z.y) + (w.z * z.z + w.w * z.w);
    }
  }
  __device__ static float eval(const ProblemDev& P, const float (&st)[NS], const float* base, float a) {
    float g = 0.f;
#pragma unroll
    for (int k = 0; k < NS; ++k)
      if (k < P.gbm_nodes) g += P.gbm_v[k] * sinf(base[k] + a * st[k]);  // base[k] = w_k0 T + w_k . x
    return g;
  }
};

template <>
struct Payoff<DPI_EQ_OU> {
  static constexpr int NS = kMaxOuK;
  static constexpr int NCH = 2 * kMaxOuK;  // (x - m_k)[j], ivar_k[j]
  // K = the mixture's component count (warp-uniform): the statistics of the unused components up to kMaxOuK are
  // skipped by uniform branches (K = 5 in the shipped instance: 3/8 of this loop and of the shuffle reductions)
  __device__ static void accumulate(float (&st)[NS], const float4 z, const float* cst, int ldc, int j0, float a, int K) {
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      if (k < K) {
        const float4 dm = *reinterpret_cast<const float4*>(cst + (2 * k) * ldc + j0);
        const float4 iv = *reinterpret_cast<const float4*>(cst + (2 * k + 1) * ldc + j0);
        const float e0 = dm.x + a * z.x, e1 = dm.y + a * z.y, e2 = dm.z + a * z.z, e3 = dm.w + a * z.w;
        st[k] += (e0 * e0 * iv.x + e1 * e1 * iv.y) + (e2 * e2 * iv.z + e3 * e3 * iv.w);
      }
    }
  }
  __device__ static float eval(const ProblemDev& P, const float (&st)[NS], const float* base, float a) {
    float l[NS], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      l[k] = (k < P.ou_K) ? P.ou_lognorm[k] - 0.5f * st[k] : -INFINITY;
      mx = fmaxf(mx, l[k]);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NS; ++k)
      if (k < P.ou_K) s += __expf(l[k] - mx);
    return -(mx + __logf(s));
  }
};

// ---------------------------------------------------------------------------------------------
#ifndef DPI_TERM_WIDE_SLOTS
#define DPI_TERM_WIDE_SLOTS 8  // NSLOT from which the kernel is held to 128 registers (2 CTAs per SM): 12 % faster at d = 1000 despite 96 B of spills
#endif
#ifndef DPI_TERM_MINBLOCKS
#define DPI_TERM_MINBLOCKS 1
#endif
template <int EQ, int NSLOT, bool DET>
__global__ void __launch_bounds__(kTermThreads, (NSLOT >= DPI_TERM_WIDE_SLOTS) ? 2 : DPI_TERM_MINBLOCKS)
terminal_kernel(const ProblemDev P, const float* __restrict__ tx, int M_T, const PhiloxKeys keys, int64_t point_base,
                const float* __restrict__ zT, const float* __restrict__ u_td, float* __restrict__ y) {
  using PO = Payoff<EQ>;
  constexpr int NS = PO::NS;
  constexpr int LDC = NSLOT * 128;  // padded dim count
  extern __shared__ __align__(16) float smem[];
  float* cst = smem;                          // [NCH][LDC]
  float* red = smem + PO::NCH * LDC;          // [2 kTermWarps][LDC + 8]: one row per half-warp
  __shared__ float base[NS > 1 ? NS : 1];     // x-dependent scalars of g
  __shared__ float red_v[2 * kTermWarps];

  const int d = P.nx;
  const int64_t p = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* txp = tx + p * (int64_t)(1 + d);
  const float t = __ldg(txp);
  const float a = sqrtf(horizon(P, t)) * P.sigma;
  const bool vec_ok = (d & 3) == 0;
  // TD estimator (data.py:928-952): a point whose step ends before T takes u_k(t + delta, X) in place of g(X)
  const bool td_interior = u_td != nullptr && P.td_delta > 0.f && t + P.td_delta < P.T;

  // ---- per-point constants -----------------------------------------------------------------
  for (int i = threadIdx.x; i < PO::NCH * LDC; i += kTermThreads) {
    const int ch = i / LDC, j = i - ch * LDC;
    float v = 0.f;
    if (j < d) {
      if (EQ == DPI_EQ_CHA) {
        v = 1.f;
      } else if (EQ == DPI_EQ_GBM) {
        if (ch < P.gbm_nodes) v = P.gbm_w[ch * (1 + d) + 1 + j];
      } else if (EQ == DPI_EQ_OU) {
        const int k = ch >> 1;
        if (k < P.ou_K) v = (ch & 1) ? P.ou_ivar[k * d + j] : (__ldg(txp + 1 + j) - P.ou_mean[k * d + j]);
      }
    }
    // Stored PERMUTED inside each 128-dim slot: the MC loop reads a lane's 8 dims as two float4 (dims 8 hl .. 8 hl + 3
    // and 8 hl + 4 .. 8 hl + 7); in natural order those are 32 bytes apart from lane to lane and the 16-byte loads of a
    // quarter-warp hit every bank twice.  With [half][lane][4] every load is one contiguous 128 bytes per quarter-warp
    // (the OU payoff issues 2 K of these loads per float4 of increments: its estimator was shared-memory bound).
    const int jq = j >> 7, jr = j & 127;
    cst[ch * LDC + (jq << 7) + (((jr >> 2) & 1) << 6) + ((jr >> 3) << 2) + (jr & 3)] = v;
  }
  if (warp == 0) {
    if (EQ == DPI_EQ_CHA) {
      float s = 0.f;
      for (int j = lane; j < d; j += 32) s += __ldg(txp + 1 + j);
      s = warp_sum(s);
      if (lane == 0) base[0] = s;
    } else if (EQ == DPI_EQ_GBM) {
      for (int k = 0; k < P.gbm_nodes; ++k) {
        float s = 0.f;
        for (int j = lane; j < d; j += 32) s += P.gbm_w[k * (1 + d) + 1 + j] * __ldg(txp + 1 + j);
        s = warp_sum(s);
        if (lane == 0) base[k] = s + P.gbm_w[k * (1 + d)] * P.T;
      }
    }
  }
  __syncthreads();

  const int klive = (EQ == DPI_EQ_OU) ? P.ou_K : NS;  // statistics in use (warp-uniform)
  // g(x): the same payoff with zero increments
  float g0;
  {
    float st[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) st[k] = 0.f;
    if (EQ == DPI_EQ_OU) {
#pragma unroll
      for (int q = 0; q < NSLOT; ++q) PO::accumulate(st, make_float4(0.f, 0.f, 0.f, 0.f), cst, LDC, 4 * lane + 128 * q, 0.f, klive);
#pragma unroll
      for (int k = 0; k < NS; ++k) st[k] = warp_sum(st[k]);
    }
    g0 = PO::eval(P, st, base, 0.f);
  }

  // ---- MC loop: half-warp <-> sample, lane <-> 8 dims per 128-dim slot (one Philox block = 8 normals) ----------
  const int hl = lane & 15, half = lane >> 4;
  float4 acc[NSLOT][2];
#pragma unroll
  for (int q = 0; q < NSLOT; ++q) acc[q][0] = acc[q][1] = make_float4(0.f, 0.f, 0.f, 0.f);
  float acc_v = 0.f;
  const uint64_t gp = (uint64_t)(point_base + p);

  for (int mb = 2 * warp; mb < M_T; mb += 2 * kTermWarps) {
    const int m = mb + half;
    const bool live = m < M_T;
    float4 z[NSLOT][2];
    float st[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) st[k] = 0.f;
#pragma unroll
    for (int q = 0; q < NSLOT; ++q) {
      const int j0 = 8 * hl + 128 * q;
      z[q][0] = z[q][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 < d && live) {
        if (DET) {
          z[q][0] = load_row4(zT + ((int64_t)p * M_T + m) * d, j0, d, vec_ok);
          z[q][1] = load_row4(zT + ((int64_t)p * M_T + m) * d, j0 + 4, d, vec_ok);
        } else {
          normal8(keys, gp, kStreamTerminal, (uint32_t)m, (uint32_t)(j0 >> 3), z[q][0], z[q][1]);
          // dims >= d of the last octet stay as drawn: every payoff statistic weights them with a zero constant and
          // their Bismut sums are never written out
        }
        PO::accumulate(st, z[q][0], cst, LDC, 128 * q + 4 * hl, a, klive);       // permuted layout: see the fill loop
        PO::accumulate(st, z[q][1], cst, LDC, 128 * q + 64 + 4 * hl, a, klive);
      } else if (EQ == DPI_EQ_OU) {
        // padded dims contribute 0 (ivar = 0); dims of a dead sample are discarded below
        PO::accumulate(st, z[q][0], cst, LDC, 128 * q + 4 * hl, a, klive);       // permuted layout: see the fill loop
        PO::accumulate(st, z[q][1], cst, LDC, 128 * q + 64 + 4 * hl, a, klive);
      }
    }
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      if (k < klive) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) st[k] += __shfl_xor_sync(0xffffffffu, st[k], o);  // over the 16 lanes of a sample
      }
    }
    float dg = 0.f;
    if (live) dg = (td_interior ? __ldg(u_td + p * M_T + m) : PO::eval(P, st, base, a)) - g0;
    if (hl == 0) acc_v += dg;
#pragma unroll
    for (int q = 0; q < NSLOT; ++q)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        acc[q][h].x = fmaf(dg, z[q][h].x, acc[q][h].x);
        acc[q][h].y = fmaf(dg, z[q][h].y, acc[q][h].y);
        acc[q][h].z = fmaf(dg, z[q][h].z, acc[q][h].z);
        acc[q][h].w = fmaf(dg, z[q][h].w, acc[q][h].w);
      }
  }

  // ---- cross-warp reduction, one label row per point -------------------------------------------
  constexpr int LDR = LDC + 8;
  const int rrow = 2 * warp + half;
#pragma unroll
  for (int q = 0; q < NSLOT; ++q) {
    *reinterpret_cast<float4*>(red + rrow * LDR + 8 * hl + 128 * q) = acc[q][0];
    *reinterpret_cast<float4*>(red + rrow * LDR + 8 * hl + 128 * q + 4) = acc[q][1];
  }
  if (hl == 0) red_v[rrow] = acc_v;
  __syncthreads();
  float* yp = y + p * (int64_t)(1 + d);
  const float inv = 1.f / ((float)M_T * a);
  for (int j = threadIdx.x; j < d; j += kTermThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 2 * kTermWarps; ++w) s += red[w * LDR + j];
    yp[1 + j] = s * inv;
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 2 * kTermWarps; ++w) s += red_v[w];
    yp[0] = s / (float)M_T + g0;
  }
}

// t = (T-2eps)(1-U)+eps, or T (1 - U_1 ... U_k) when n_factors = k >= 1 (the default sampler of the reference,
// data.py:149-159: k = N - i + 1 uniforms per point, Philox block b of stream T holds factors 4b..4b+3);
// x = x0 + sqrt(t) sigma Z  (one warp per point)
__global__ void __launch_bounds__(256)
sample_tx_kernel(const ProblemDev P, int64_t n, int64_t point_base, uint64_t seed, float eps, int n_factors,
                 float* __restrict__ tx) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= n) return;
  const int d = P.nx;
  const uint64_t gp = (uint64_t)(point_base + p);
  float t;
  if (n_factors <= 0) {
    const uint4 r = philox_at(seed, gp, kStreamT, 0u, 0u);
    t = (P.T - 2.f * eps) * (1.f - u01(r.x)) + eps;
  } else {
    float prod = 1.f;  // every lane computes the same product (warp-uniform, a few Philox blocks)
    for (int b = 0; 4 * b < n_factors; ++b) {
      const uint4 r = philox_at(seed, gp, kStreamT, 0u, (uint32_t)b);
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (4 * b + q < n_factors) prod *= u01(w[q]);
    }
    // tx carries t in fp32: keep T - t representable (>= 2^-21 T).  With many factors the product underflows the fp32
    // spacing below T in ~1 % of the points; the fp64 reference keeps T - t ~ 1e-16 there, fp32 would get 0 and divide by it.
    t = P.T * (1.f - fmaxf(prod, 4.76837158203125e-7f));
  }
  float* row = tx + p * (int64_t)(1 + d);
  if (lane == 0) row[0] = t;
  const float a = sqrtf(t) * P.sigma;
  for (int j0 = 4 * lane; j0 < d; j0 += 128) {
    const float4 z = normal4(seed, gp, kStreamX, 0u, (uint32_t)(j0 >> 2));
    float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (P.ou_x0_std > 0.f && P.eq == DPI_EQ_OU) {
      x0 = normal4(seed, gp, kStreamX0, 0u, (uint32_t)(j0 >> 2));
      x0.x *= P.ou_x0_std; x0.y *= P.ou_x0_std; x0.z *= P.ou_x0_std; x0.w *= P.ou_x0_std;
    }
    const float v[4] = {x0.x + a * z.x, x0.y + a * z.y, x0.z + a * z.z, x0.w + a * z.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (j0 + q < d) row[1 + j0 + q] = v[q];
  }
}

// TD estimator: the path ends (t_next, x + sqrt(t_next - t) sigma Z) of `np` points as network input rows
// [np * M_T][1 + d], with the terminal stream's increments (one thread per row and octet of dims)
__global__ void __launch_bounds__(256)
td_rows_kernel(const ProblemDev P, const float* __restrict__ tx, int64_t np, int M_T, const PhiloxKeys keys,
               int64_t point_base, const float* __restrict__ zT, float* __restrict__ rows) {
  const int d = P.nx, no = (d + 7) / 8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np * M_T * no) return;
  const int o = (int)(i % no);
  const int64_t r = i / no, p = r / M_T;
  const int m = (int)(r - p * M_T);
  const float* txp = tx + p * (int64_t)(1 + d);
  const float t = __ldg(txp), hz = horizon(P, t), a = sqrtf(hz) * P.sigma;
  float4 za, zb;
  if (zT) {
    za = load_row4(zT + r * d, 8 * o, d, (d & 3) == 0);
    zb = load_row4(zT + r * d, 8 * o + 4, d, (d & 3) == 0);
  } else {
    normal8(keys, (uint64_t)(point_base + p), kStreamTerminal, (uint32_t)m, (uint32_t)o, za, zb);
  }
  float* row = rows + r * (int64_t)(1 + d);
  if (o == 0) row[0] = t + hz;
  const float z[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int j = 8 * o + q;
    if (j < d) row[1 + j] = fmaf(a, z[q], __ldg(txp + 1 + j));
  }
}

template <int EQ, int NSLOT>
int launch_terminal_t(const ProblemDev& P, const float* tx, int64_t n, int M_T, uint64_t seed, int64_t base,
                      const float* zT, const float* u_td, float* y, cudaStream_t st) {
  using PO = Payoff<EQ>;
  constexpr int LDC = NSLOT * 128;
  const size_t smem = sizeof(float) * (PO::NCH * LDC + 2 * kTermWarps * (LDC + 8));
  auto kd = terminal_kernel<EQ, NSLOT, true>;
  auto kp = terminal_kernel<EQ, NSLOT, false>;
  if (smem > 48 * 1024) {
    DPI_CUDA_CHECK(cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DPI_CUDA_CHECK(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (zT)
    kd<<<(unsigned)n, kTermThreads, smem, st>>>(P, tx, M_T, make_philox_keys(seed), base, zT, u_td, y);
  else
    kp<<<(unsigned)n, kTermThreads, smem, st>>>(P, tx, M_T, make_philox_keys(seed), base, nullptr, u_td, y);
  DPI_CUDA_CHECK(cudaGetLastError());
  return DPI_OK;
}

template <int EQ>
int launch_terminal_eq(const ProblemDev& P, const float* tx, int64_t n, int M_T, uint64_t seed, int64_t base,
                       const float* zT, const float* u_td, float* y, cudaStream_t st) {
  const int d = P.nx;
  if (d <= 128) return launch_terminal_t<EQ, 1>(P, tx, n, M_T, seed, base, zT, u_td, y, st);
  if (d <= 256) return launch_terminal_t<EQ, 2>(P, tx, n, M_T, seed, base, zT, u_td, y, st);
  if (d <= 512) return launch_terminal_t<EQ, 4>(P, tx, n, M_T, seed, base, zT, u_td, y, st);
  if (d <= 1024) return launch_terminal_t<EQ, 8>(P, tx, n, M_T, seed, base, zT, u_td, y, st);
  set_error("nx=%d > 1024 is not supported by the terminal kernel", d);
  return DPI_ERR_UNSUPPORTED;
}

}  // namespace

int launch_terminal(const ProblemDev& P, const float* tx, int64_t n, int M_T, uint64_t seed, int64_t base,
                    const float* zT, float* y, cudaStream_t st, const float* u_td) {
  if (n <= 0) return DPI_OK;
  switch (P.eq) {
    case DPI_EQ_CHA: return launch_terminal_eq<DPI_EQ_CHA>(P, tx, n, M_T, seed, base, zT, u_td, y, st);
    case DPI_EQ_GBM: return launch_terminal_eq<DPI_EQ_GBM>(P, tx, n, M_T, seed, base, zT, u_td, y, st);
    case DPI_EQ_OU: return launch_terminal_eq<DPI_EQ_OU>(P, tx, n, M_T, seed, base, zT, u_td, y, st);
  }
  set_error("unknown equation kind %d", P.eq);
  return DPI_ERR_INVALID;
}

int launch_sample_tx(const ProblemDev& P, int64_t n, int64_t base, uint64_t seed, float eps, int n_factors, float* tx,
                     cudaStream_t st) {
  if (n <= 0) return DPI_OK;
  const int wpb = 8;
  sample_tx_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, st>>>(P, n, base, seed, eps, n_factors, tx);
  DPI_CUDA_CHECK(cudaGetLastError());
  return DPI_OK;
}

int launch_td_rows(const ProblemDev& P, const float* tx, int64_t np, int M_T, uint64_t seed, int64_t base, const float* zT,
                   float* rows, cudaStream_t st) {
  if (np <= 0) return DPI_OK;
  const int64_t total = np * M_T * ((P.nx + 7) / 8);
  td_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(P, tx, np, M_T, make_philox_keys(seed), base, zT, rows);
  DPI_CUDA_CHECK(cudaGetLastError());
  return DPI_OK;
}

}  // namespace dpi
