// fp32 integral estimator for the fully nonlinear equation `GBMEquationComplexExact` with SDGD Hessian sampling.
//
// Restates picard/data.py:471-527 (estimator), :497-502 (SDGD index draw), :1226-1325 (get_f: value, gradient and the
// v-fold double-backward loop for the sampled Hessian-diagonal entries, :1273-1303 incl. the baseline gather) and
// picard/equations.py:457-466 (ffi) / :421-455 (exact-solution source terms).
//
// Instead of v reverse-over-reverse sweeps per path row, the second derivative along e_j is propagated forward as a
// 2-jet:   dz1 = W1[:,1+j], ddz1 = 0;   dh = ELU'(z) dz,  ddh = ELU''(z) dz^2 + ELU'(z) ddz;
// dz+ = W dh, ddz+ = W ddh;  u_jj = w_out . ddh_last.   One CTA owns one collocation point.  A tile is ONE path row
// with its v <= 128 sampled directions: the (dz, ddz) pairs of all directions form a [2*128 x H] GEMM operand in
// shared memory that goes through the hidden layers as a register-tiled GEMM chain, with the per-column ELU'/ELU''
// factors of that path row applied in the epilogue.  The baseline needs the FULL Hessian diagonal at (t, x): one
// extra tile with directions j = 0..d-1, whose result is then gathered at each path's own indices.
#include "dpi_common.cuh"
#include "dpi_kernels.h"

namespace dpi {
namespace {

constexpr int kH = 64;           // padded hidden width
constexpr int kThreads = 256;
constexpr int kDirs = 128;       // direction rows per tile (v and d must be <= 128)
constexpr int kRows = 2 * kDirs; // GEMM rows: (dz, ddz) interleaved
constexpr int kLDA = kRows + 4;
constexpr int kLDW1 = kH + 1;    // padded row of the transposed first-layer weight (bank spread for random rows)
constexpr int kMaxHidden = 4;    // hidden GEMM layers (L - 2)

struct GbmSmem {
  float act[kH * kLDA];              // [k][row]
  float Wt1[128 * kLDW1];            // [j][h] = W1[h][1+j]
  float Wh[kMaxHidden][kH * kH];     // [k][h] = W_l[h][k]
  float bh[kMaxHidden][kH];
  float e1[kMaxHidden + 1][kH];      // ELU'(z_l) of the current path row, per activation layer
  float e2[kMaxHidden + 1][kH];      // ELU''(z_l)
  float hcur[kH];                    // current activation vector of the value chain
  float c0[kH], wt[kH], wout[kH];
  float red[8][kDirs];               // cross-warp partial sums
  float xs[128], zs[128], diag[128];
  float gw[kMaxGbmNodes][128];       // GBM w_k[1+j]
  float part[4][kH];
  int idx[kDirs];
  int idx2[kDirs];                   // pair directions e_i + e_j (full Hessian): second coordinate, -1 = none
  float scal[16];
};

// value chain of one path row: z1 = c0 + s w_t + a (W_x Z), then hidden layers; fills e1/e2 per layer, returns u
__device__ float forward_row(GbmSmem& S, const MlpDev& N, int d, float s, float a, float bout, bool use_z) {
  const int tid = threadIdx.x, col = tid & 63, part = tid >> 6;
  const int n_hidden = N.n_linear - 2;
  {
    float acc = 0.f;
    if (use_z)
      for (int j = part; j < d; j += 4) acc = fmaf(S.Wt1[j * kLDW1 + col], S.zs[j], acc);
    S.part[part][col] = acc;
  }
  __syncthreads();
  if (tid < kH) {
    const float z = S.c0[tid] + s * S.wt[tid] + a * (S.part[0][tid] + S.part[1][tid] + S.part[2][tid] + S.part[3][tid]);
    const float e = __expf(fminf(z, 0.f));
    const bool pos = z > 0.f;
    S.hcur[tid] = pos ? z : e - 1.f;
    S.e1[0][tid] = pos ? 1.f : e;
    S.e2[0][tid] = pos ? 0.f : e;
  }
  __syncthreads();
  for (int l = 0; l < n_hidden; ++l) {
    float acc = 0.f;
    for (int k = part * 16; k < part * 16 + 16; ++k) acc = fmaf(S.Wh[l][k * kH + col], S.hcur[k], acc);
    S.part[part][col] = acc;
    __syncthreads();
    if (tid < kH) {
      const float z = S.bh[l][tid] + S.part[0][tid] + S.part[1][tid] + S.part[2][tid] + S.part[3][tid];
      const float e = __expf(fminf(z, 0.f));
      const bool pos = z > 0.f;
      S.hcur[tid] = pos ? z : e - 1.f;
      S.e1[l + 1][tid] = pos ? 1.f : e;
      S.e2[l + 1][tid] = pos ? 0.f : e;
    }
    __syncthreads();
  }
  float u = 0.f;
  if (tid < 32) {
    u = S.wout[tid] * S.hcur[tid] + S.wout[tid + 32] * S.hcur[tid + 32];
    u = warp_sum(u) + bout;
  }
  return u;  // valid in warp 0
}

// second derivatives u_jj for the directions idx[0..ndir): result in S.red[0][i]
// (pairs: direction i is e_idx[i] + e_idx2[i] where idx2[i] >= 0 -- polarisation for the mixed derivatives)
__device__ void jet_tile(GbmSmem& S, const MlpDev& N, int ndir, bool pairs = false) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = lane, txg = warp;  // 32 row groups (8 GEMM rows = 4 directions), 8 column groups (8 columns)
  const int n_hidden = N.n_linear - 2;
  // first-layer jet: dh = ELU'(z1) w, ddh = ELU''(z1) w^2 with w = W1[:,1+j]
  {
    const int i = tid & 127, k0 = tid >> 7;
    const int j = (i < ndir) ? S.idx[i] : -1;
    for (int k = k0; k < kH; k += 2) {
      float w = (j >= 0) ? S.Wt1[j * kLDW1 + k] : 0.f;
      if (pairs && j >= 0 && S.idx2[i] >= 0) w += S.Wt1[S.idx2[i] * kLDW1 + k];
      *reinterpret_cast<float2*>(&S.act[k * kLDA + 2 * i]) = make_float2(S.e1[0][k] * w, S.e2[0][k] * w * w);
    }
  }
  __syncthreads();
  for (int l = 0; l < n_hidden; ++l) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
    const float* W = S.Wh[l];
#pragma unroll 4
    for (int k = 0; k < kH; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&S.act[k * kLDA + ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&S.act[k * kLDA + ty * 8 + 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&W[k * kH + txg * 8]);
      const float4 w1 = *reinterpret_cast<const float4*>(&W[k * kH + txg * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[i][c] = fmaf(av[i], wv[c], acc[i][c]);
    }
    __syncthreads();
    if (l < n_hidden - 1) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int col = txg * 8 + c;
        const float e1 = S.e1[l + 1][col], e2 = S.e2[l + 1][col];
        float o[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float dz = acc[2 * i][c], ddz = acc[2 * i + 1][c];
          o[2 * i] = e1 * dz;
          o[2 * i + 1] = fmaf(e2 * dz, dz, e1 * ddz);
        }
        float4* dst = reinterpret_cast<float4*>(&S.act[col * kLDA + ty * 8]);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
      __syncthreads();
    } else {
      // last hidden layer: u_jj = w_out . ddh
      float pj[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int col = txg * 8 + c;
        const float e1 = S.e1[l + 1][col], e2 = S.e2[l + 1][col], wo = S.wout[col];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float dz = acc[2 * i][c], ddz = acc[2 * i + 1][c];
          pj[i] = fmaf(wo, fmaf(e2 * dz, dz, e1 * ddz), pj[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) S.red[txg][ty * 4 + i] = pj[i];
      __syncthreads();
      if (tid < kDirs) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += S.red[w][tid];
        S.red[0][tid] = s;  // only thread `tid` reads red[1..7][tid] and writes red[0][tid]
      }
      __syncthreads();
    }
  }
}

// block-wide sums of up to 3 per-thread values (result broadcast through shared memory)
__device__ void block_sum3(GbmSmem& S, float& a, float& b, float& c) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a = warp_sum(a), b = warp_sum(b), c = warp_sum(c);
  __syncthreads();
  if (lane == 0) S.part[0][warp] = a, S.part[1][warp] = b, S.part[2][warp] = c;
  __syncthreads();
  a = b = c = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) a += S.part[0][w], b += S.part[1][w], c += S.part[2][w];
}

// u_k-independent part of ffi at (s, X = x + a Z): -u*_t - 1/2 Lap u* - 1/4 sum_j |d2_jj u*|  (equations.py:457-466)
__device__ float gbm_source(GbmSmem& S, const ProblemDev& P, float s, float a, bool use_z) {
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
    block_sum3(S, p0, p1, p2);
    float z = 0.f;
    block_sum3(S, p3, z, dummy);
    arg[0] = p0, arg[1] = p1, arg[2] = p2, arg[3] = p3;
  }
  float ut = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxGbmNodes; ++k) {
    sn[k] = 0.f;
    if (k < P.gbm_nodes) {
      const float ak = arg[k] + S.scal[8 + k] * s;  // scal[8+k] = w_k0
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

template <bool ZERO_NET>
__global__ void __launch_bounds__(kThreads, 1)
gbm_simt_kernel(const ProblemDev P, const MlpDev N, const float* __restrict__ tx, int M_I, uint64_t seed,
                int64_t point_base, const float* __restrict__ u_s, const float* __restrict__ zS,
                const int32_t* __restrict__ idx_in, float bound, float* __restrict__ y) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  GbmSmem& S = *reinterpret_cast<GbmSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = P.nx, v = P.sdgd_v;
  const int64_t p = blockIdx.x;
  const uint64_t gp = (uint64_t)(point_base + p);
  const bool det = zS != nullptr;
  const float* txp = tx + p * (int64_t)(1 + d);
  const float t = __ldg(txp), dT = horizon(P, t);
  const float fd = (float)d;
  float bout = 0.f;

  // ---- stage the network and the point ----
  for (int i = tid; i < 128; i += kThreads) {
    S.xs[i] = (i < d) ? __ldg(txp + 1 + i) : 0.f;
    S.zs[i] = 0.f;
    for (int k = 0; k < kMaxGbmNodes; ++k) S.gw[k][i] = (k < P.gbm_nodes && i < d) ? P.gbm_w[k * (1 + d) + 1 + i] : 0.f;
  }
  if (tid < kMaxGbmNodes) S.scal[8 + tid] = (tid < P.gbm_nodes) ? P.gbm_w[tid * (1 + d)] : 0.f;
  if (!ZERO_NET) {
    const int n_hidden = N.n_linear - 2;
    for (int i = tid; i < 128 * kH; i += kThreads) {
      const int j = i / kH, h = i - j * kH;
      S.Wt1[j * kLDW1 + h] = N.Wt1[i];  // packed [d_pad=128][HP=64]
    }
    for (int l = 0; l < n_hidden; ++l) {
      for (int i = tid; i < kH * kH; i += kThreads) S.Wh[l][i] = N.Wth[l][i];
      if (tid < kH) S.bh[l][tid] = N.bh[l][tid];
    }
    if (tid < kH) {
      S.wt[tid] = N.wt[tid];
      S.wout[tid] = N.wout[tid];
    }
    bout = __ldg(N.bout);
  }
  __syncthreads();
  if (!ZERO_NET) {
    // c0 = b1 + W_x x
    const int col = tid & 63, part = tid >> 6;
    float acc = 0.f;
    for (int j = part; j < d; j += 4) acc = fmaf(S.Wt1[j * kLDW1 + col], S.xs[j], acc);
    S.part[part][col] = acc;
    __syncthreads();
    if (tid < kH) S.c0[tid] = N.b1[tid] + S.part[0][tid] + S.part[1][tid] + S.part[2][tid] + S.part[3][tid];
    __syncthreads();
  }

  // ---- baseline: source(t,x) and the full Hessian diagonal of u_k at (t,x) (data.py:1276-1278) ----
  const float src_b = gbm_source(S, P, t, 0.f, false);
  if (!ZERO_NET) {
    if (tid < kDirs) S.idx[tid] = tid;
    forward_row(S, N, d, t, 0.f, bout, false);
    jet_tile(S, N, d);
    if (tid < 128) S.diag[tid] = (tid < d) ? S.red[0][tid] : 0.f;
  } else {
    if (tid < 128) S.diag[tid] = 0.f;
  }
  __syncthreads();

  float G = 0.f, fsum = 0.f;  // thread j < d owns gradient component j; fsum is uniform
  const float alpha_term = 0.5f * (1.f - P.alpha) * fd, abs_term = 0.25f * fd, inv_v = 1.f / (float)v;

  for (int m = 0; m < M_I; ++m) {
    // ---- draws of this path row ----
    float uu;
    if (det)
      uu = __ldg(u_s + p * M_I + m);
    else
      uu = u01(philox_at(seed, gp, kStreamS, (uint32_t)m, 0u).x);
    const float s = uu * dT + t, a = sqrtf(uu * dT) * P.sigma;
    if (warp == 0) {
      const int j0 = 4 * lane;
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 < d) {
        if (det)
          z = load_row4(zS + ((int64_t)p * M_I + m) * d, j0, d, (d & 3) == 0);
        else
          z = normal4(seed, gp, kStreamIntegral, (uint32_t)m, (uint32_t)lane);
      }
      S.zs[j0] = (j0 < d) ? z.x : 0.f;
      S.zs[j0 + 1] = (j0 + 1 < d) ? z.y : 0.f;
      S.zs[j0 + 2] = (j0 + 2 < d) ? z.z : 0.f;
      S.zs[j0 + 3] = (j0 + 3 < d) ? z.w : 0.f;
    } else if (warp == 1) {
      // SDGD coordinates: v draws with replacement from {0..d-1} (data.py:497-502)
      int jj[4];
      if (det) {
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
    float mean_jj = 0.f, mean_abs = 0.f;
    if (!ZERO_NET) {
      forward_row(S, N, d, s, a, bout, true);
      jet_tile(S, N, v);
    }
    // sampled second derivatives of the path row, and the baseline's diagonal gathered at the same indices
    float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f, z0 = 0.f;
    if (tid < v) {
      const float ujj = ZERO_NET ? 0.f : S.red[0][tid];
      const float dj = S.diag[S.idx[tid]];
      a1 = ujj, a2 = fabsf(ujj), b1 = dj, b2 = fabsf(dj);
    }
    block_sum3(S, a1, a2, b1);
    block_sum3(S, b2, z0, z0);
    mean_jj = a1 * inv_v, mean_abs = a2 * inv_v;
    const float f = alpha_term * mean_jj + abs_term * mean_abs + src;
    const float fb = alpha_term * (b1 * inv_v) + abs_term * (b2 * inv_v) + src_b;
    fsum += dT * f;
    const float coef = dT * (f - fb) / a;
    if (tid < d) G = fmaf(coef, S.zs[tid], G);
    __syncthreads();  // zs / idx are rewritten by the next row
  }

  // ---- y += integral part; clip (data.py:222) ----
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
}


// Hessian diagonal of the MLP u_k at the rows of tx (the evaluation callback's TEST_HESSIAN metric on its diagonal,
// picard/utils.py:329-478,579-591): the baseline tile of the label kernel on its own -- one CTA per point, forward
// value chain at (t, x), then the 2-jets of all d coordinate directions.
__global__ void __launch_bounds__(kThreads, 1)
mlp_hess_diag_kernel(const ProblemDev P, const MlpDev N, const float* __restrict__ tx, float* __restrict__ diag) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  GbmSmem& S = *reinterpret_cast<GbmSmem*>(smem_raw);
  const int tid = threadIdx.x, d = P.nx;
  const int64_t p = blockIdx.x;
  const float* txp = tx + p * (int64_t)(1 + d);
  const float t = __ldg(txp);
  for (int i = tid; i < 128; i += kThreads) {
    S.xs[i] = (i < d) ? __ldg(txp + 1 + i) : 0.f;
    S.zs[i] = 0.f;
  }
  const int n_hidden = N.n_linear - 2;
  for (int i = tid; i < 128 * kH; i += kThreads) {
    const int j = i / kH, h = i - j * kH;
    S.Wt1[j * kLDW1 + h] = N.Wt1[i];
  }
  for (int l = 0; l < n_hidden; ++l) {
    for (int i = tid; i < kH * kH; i += kThreads) S.Wh[l][i] = N.Wth[l][i];
    if (tid < kH) S.bh[l][tid] = N.bh[l][tid];
  }
  if (tid < kH) {
    S.wt[tid] = N.wt[tid];
    S.wout[tid] = N.wout[tid];
  }
  const float bout = __ldg(N.bout);
  __syncthreads();
  {
    const int col = tid & 63, part = tid >> 6;
    float acc = 0.f;
    for (int j = part; j < d; j += 4) acc = fmaf(S.Wt1[j * kLDW1 + col], S.xs[j], acc);
    S.part[part][col] = acc;
    __syncthreads();
    if (tid < kH) S.c0[tid] = N.b1[tid] + S.part[0][tid] + S.part[1][tid] + S.part[2][tid] + S.part[3][tid];
    __syncthreads();
  }
  if (tid < kDirs) S.idx[tid] = tid;
  forward_row(S, N, d, t, 0.f, bout, false);
  jet_tile(S, N, d);
  if (tid < d) diag[p * (int64_t)d + tid] = S.red[0][tid];
}

// Full Hessian of u_k in x (EvalCallback with TEST_HESSIAN, picard/utils.py:198-200,579-591: d backward passes upstream).
// One CTA per point: the value chain once, the d coordinate directions (diagonal), then the d (d - 1) / 2 pair directions
// e_i + e_j in tiles of 128 jets; H_ij = (D2[e_i + e_j] - H_ii - H_jj) / 2.  H [n][d][d], symmetric.
__global__ void __launch_bounds__(kThreads, 1)
mlp_hess_full_kernel(const ProblemDev P, const MlpDev N, const float* __restrict__ tx, float* __restrict__ H) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  GbmSmem& S = *reinterpret_cast<GbmSmem*>(smem_raw);
  const int tid = threadIdx.x, d = P.nx;
  const int64_t p = blockIdx.x;
  const float* txp = tx + p * (int64_t)(1 + d);
  const float t = __ldg(txp);
  for (int i = tid; i < 128; i += kThreads) {
    S.xs[i] = (i < d) ? __ldg(txp + 1 + i) : 0.f;
    S.zs[i] = 0.f;
  }
  const int n_hidden = N.n_linear - 2;
  for (int i = tid; i < 128 * kH; i += kThreads) {
    const int j = i / kH, h = i - j * kH;
    S.Wt1[j * kLDW1 + h] = N.Wt1[i];
  }
  for (int l = 0; l < n_hidden; ++l) {
    for (int i = tid; i < kH * kH; i += kThreads) S.Wh[l][i] = N.Wth[l][i];
    if (tid < kH) S.bh[l][tid] = N.bh[l][tid];
  }
  if (tid < kH) {
    S.wt[tid] = N.wt[tid];
    S.wout[tid] = N.wout[tid];
  }
  const float bout = __ldg(N.bout);
  __syncthreads();
  {
    const int col = tid & 63, part = tid >> 6;
    float acc = 0.f;
    for (int j = part; j < d; j += 4) acc = fmaf(S.Wt1[j * kLDW1 + col], S.xs[j], acc);
    S.part[part][col] = acc;
    __syncthreads();
    if (tid < kH) S.c0[tid] = N.b1[tid] + S.part[0][tid] + S.part[1][tid] + S.part[2][tid] + S.part[3][tid];
    __syncthreads();
  }
  if (tid < kDirs) S.idx[tid] = tid, S.idx2[tid] = -1;
  forward_row(S, N, d, t, 0.f, bout, false);
  jet_tile(S, N, d, true);
  float* Hp = H + p * (int64_t)d * d;
  if (tid < d) {
    S.diag[tid] = S.red[0][tid];
    Hp[tid * d + tid] = S.red[0][tid];
  }
  __syncthreads();
  const int npairs = d * (d - 1) / 2;
  for (int q0 = 0; q0 < npairs; q0 += kDirs) {
    // pair q = (i, j), i < j, enumerated row by row: q = i d - i (i + 1) / 2 + (j - i - 1)
    int pi = -1, pj = -1;
    if (tid < kDirs && q0 + tid < npairs) {
      const int q = q0 + tid;
      int i = (int)floorf(((2.f * d - 1.f) - sqrtf((2.f * d - 1.f) * (2.f * d - 1.f) - 8.f * (float)q)) * 0.5f);
      while (i > 0 && i * d - i * (i + 1) / 2 > q) --i;                      // (guards the float estimate)
      while ((i + 1) * d - (i + 1) * (i + 2) / 2 <= q) ++i;
      pi = i, pj = q - (i * d - i * (i + 1) / 2) + i + 1;
    }
    if (tid < kDirs) S.idx[tid] = pi, S.idx2[tid] = pj;
    __syncthreads();
    jet_tile(S, N, min(kDirs, npairs - q0), true);
    if (tid < kDirs && pi >= 0) {
      const float hij = 0.5f * (S.red[0][tid] - S.diag[pi] - S.diag[pj]);
      Hp[pi * d + pj] = hij;
      Hp[pj * d + pi] = hij;
    }
    __syncthreads();
  }
}

}  // namespace

int launch_gbm_simt(const ProblemDev& P, const MlpDev& N, bool zero_net, const float* tx, int64_t n, int M_I,
                    uint64_t seed, int64_t base, const float* u_s, const float* zS, const int32_t* idx, float bound,
                    float* y, cudaStream_t st) {
  if (n <= 0) return DPI_OK;
  if (P.nx > 128 || P.sdgd_v < 1 || P.sdgd_v > 128) {
    set_error("GBM/SDGD kernel needs nx <= 128 and 1 <= v <= 128 (got nx=%d v=%d)", P.nx, P.sdgd_v);
    return DPI_ERR_UNSUPPORTED;
  }
  if (!zero_net && (N.HP != 64 || N.n_linear - 2 > kMaxHidden || N.n_linear < 3 || N.d_pad != 128)) {
    set_error("GBM/SDGD kernel needs an MLP with 1..%d hidden GEMM layers of width <= 64", kMaxHidden);
    return DPI_ERR_UNSUPPORTED;
  }
  if (zS && !idx && P.sdgd_v > 0) {
    set_error("deterministic mode with SDGD needs the idx draws");
    return DPI_ERR_INVALID;
  }
  const size_t smem = sizeof(GbmSmem);
  auto kz = gbm_simt_kernel<true>;
  auto kn = gbm_simt_kernel<false>;
  DPI_CUDA_CHECK(cudaFuncSetAttribute(kz, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DPI_CUDA_CHECK(cudaFuncSetAttribute(kn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (zero_net)
    kz<<<(unsigned)n, kThreads, smem, st>>>(P, N, tx, M_I, seed, base, u_s, zS, idx, bound, y);
  else
    kn<<<(unsigned)n, kThreads, smem, st>>>(P, N, tx, M_I, seed, base, u_s, zS, idx, bound, y);
  DPI_CUDA_CHECK(cudaGetLastError());
  return DPI_OK;
}

}  // namespace dpi

namespace dpi {
int launch_mlp_hess_full(const ProblemDev& P, const MlpDev& N, const float* tx, int64_t n, float* H, cudaStream_t st) {
  if (n <= 0) return DPI_OK;
  if (P.nx > 128 || N.HP != 64 || N.n_linear - 2 > kMaxHidden || N.n_linear < 3 || N.d_pad != 128) {
    set_error("dpi_net_hessian: needs nx <= 128 and an MLP with 1..%d hidden GEMM layers of width <= 64", kMaxHidden);
    return DPI_ERR_UNSUPPORTED;
  }
  const size_t smem = sizeof(GbmSmem);
  DPI_CUDA_CHECK(cudaFuncSetAttribute(mlp_hess_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mlp_hess_full_kernel<<<(unsigned)n, kThreads, smem, st>>>(P, N, tx, H);
  DPI_CUDA_CHECK(cudaGetLastError());
  return DPI_OK;
}

int launch_mlp_hess_diag(const ProblemDev& P, const MlpDev& N, const float* tx, int64_t n, float* diag, cudaStream_t st) {
  if (n <= 0) return DPI_OK;
  if (P.nx > 128 || N.HP != 64 || N.n_linear - 2 > kMaxHidden || N.n_linear < 3 || N.d_pad != 128) {
    set_error("Hessian diagonal: built for MLPs with 1..%d hidden GEMM layers of width <= 64 and nx <= 128", kMaxHidden);
    return DPI_ERR_UNSUPPORTED;
  }
  const size_t smem = sizeof(GbmSmem);
  DPI_CUDA_CHECK(cudaFuncSetAttribute(mlp_hess_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mlp_hess_diag_kernel<<<(unsigned)n, kThreads, smem, st>>>(P, N, tx, diag);
  DPI_CUDA_CHECK(cudaGetLastError());
  return DPI_OK;
}
}  // namespace dpi
