// Regression step of a Picard iteration for a scalar-output MLP u_theta: the value + gradient
// matching loss of `PicardSolutionGradientWrapper.training_step` (picard/solution_jac.py:172-216) with
// `FixedLossScaler` (picard/solution_jac.py:77-89), `loss_fn = torch.square` (picard/solution.py:66-68), the time weight
// exp(beta t) (picard/solution_jac.py:189), and `torch.optim.Adam` (picard/solution.py:95-97).
//
// The reference differentiates vmap(jacrev(u)) by reverse mode.  Here, with r = dL/d(grad_x u) held fixed,
//     d/dtheta <r, grad_x u> = d/dtheta JVP_r(u),
// so the parameter gradient is ONE reverse sweep through the (value, tangent) pair of networks - the same pair the label
// kernels evaluate - with ELU'' entering through the tangent branch:
//     zbar_l = hbar_l . ELU'(z_l) + hdbar_l . ELU''(z_l) . zd_l ,   zdbar_l = hdbar_l . ELU'(z_l),
//     dW_l   = zbar_l^T h_{l-1} + zdbar_l^T hd_{l-1} ,   db_l = sum_b zbar_l ,   (hbar, hdbar)_{l-1} = (zbar, zdbar)_l W_l.
// Value rows and tangent rows are stacked ([2B, H] buffers), so every layer of the sweep is one GEMM of the fused-epilogue
// fp32 GEMM of dpi_hjb.cu.  fp32 master weights are updated in place in the caller's tensors (torch layout).
#include <algorithm>
#include <cmath>

#include "dpi_common.cuh"
#include "dpi_kernels.h"

namespace dpi {
namespace {

// loss terms, dL/du and dL/d(grad_x u).  One warp per row writes the row's two loss terms; a second, single-CTA
// kernel adds the B row terms in a fixed order (deterministic; a single 1024-thread CTA over all B (1+d) elements
// took 52 us, a fifth of the whole regression step).
__global__ void __launch_bounds__(256)
train_loss_kernel(const float* __restrict__ tx, const float* __restrict__ y, const float* __restrict__ hlast /* [B][H] */,
                  const float* __restrict__ wout, const float* __restrict__ bout, int H,
                  const float* __restrict__ g /* [B][1+d] */, int B, int d, float beta, float fw,
                  float* __restrict__ a /* [B] */, float* __restrict__ r /* [B][1+d] */,
                  float* __restrict__ rowl /* [B][2] */) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  // output layer u_b = w_out . h_b + b_out (a GEMV: not worth a launch of its own on the step's critical chain)
  float ub = 0.f;
  for (int k = lane; k < H; k += 32) ub = fmaf(__ldg(hlast + (int64_t)b * H + k), __ldg(wout + k), ub);
  ub = warp_sum(ub) + __ldg(bout);
  const int ld = 1 + d;
  const float invB = 1.f / (float)B;
  const float w = expf(beta * __ldg(tx + (int64_t)b * ld));
  const int64_t row = (int64_t)b * ld;
  float lg = 0.f;
  for (int j = 1 + lane; j < ld; j += 32) {
    const float dg = __ldg(g + row + j) - __ldg(y + row + j);
    lg += w * dg * dg;
    r[row + j] = 2.f * fw * w * dg * invB;
  }
  lg = warp_sum(lg);
  if (lane == 0) {
    const float dv = ub - __ldg(y + row);
    a[b] = 2.f * w * dv * invB;
    r[row] = 0.f;
    rowl[2 * b] = w * dv * dv;
    rowl[2 * b + 1] = lg;
  }
}
__global__ void __launch_bounds__(256)
train_loss_finish_kernel(const float* __restrict__ rowl, int B, float fw, float* __restrict__ losses) {
  __shared__ float red[2][8];
  float lv = 0.f, lg = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) lv += rowl[2 * b], lg += rowl[2 * b + 1];
  lv = warp_sum(lv), lg = warp_sum(lg);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = lv, red[1][threadIdx.x >> 5] = lg;
  __syncthreads();
  if (threadIdx.x == 0) {
    lv = lg = 0.f;
    for (int i = 0; i < 8; ++i) lv += red[0][i], lg += red[1][i];
    const float invB = 1.f / (float)B;
    losses[1] = lv * invB;
    losses[2] = lg * invB;
    losses[0] = lv * invB + fw * lg * invB;
  }
}

// the same for the last hidden layer, whose incoming adjoints are hbar = a_b w_out (value rows), hdbar = w_out
// (tangent rows): no buffer, no separate kernel
__global__ void train_adjoint_top_kernel(const float* __restrict__ a, const float* __restrict__ wout,
                                         const float* __restrict__ e, const float* __restrict__ zd, int B, int H,
                                         float* __restrict__ zb /* [2B][H] */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, k = i - b * H;
  const float w = __ldg(wout + k);
  const float ee = __ldg(e + i), hbar = __ldg(a + b) * w, hdbar = w;
  const float e2 = ee < 1.f ? ee : 0.f;
  zb[i] = fmaf(hdbar * e2, __ldg(zd + i), hbar * ee);
  zb[(int64_t)B * H + i] = hdbar * ee;
}

// adjoints through one ELU pair: zb = [hbar . e + hdbar . e2 . zd ; hdbar . e], e2 = ELU'' = (e < 1 ? e : 0)
__global__ void train_adjoint_kernel(const float* __restrict__ hb /* [2B][H] */, const float* __restrict__ e,
                                     const float* __restrict__ zd, int B, int H, float* __restrict__ zb /* [2B][H] */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const float ee = __ldg(e + i), hbar = __ldg(hb + i), hdbar = __ldg(hb + (int64_t)B * H + i);
  const float e2 = ee < 1.f ? ee : 0.f;
  zb[i] = fmaf(hdbar * e2, __ldg(zd + i), hbar * ee);
  zb[(int64_t)B * H + i] = hdbar * ee;
}

// column sums over the first B rows: out[k] = sum_b w_b x[b][k] (w = null: plain sum).  One warp per column.
__global__ void train_colsum_kernel(const float* __restrict__ x, int ld, const float* __restrict__ w, int B, int H,
                                    const float* __restrict__ x2 /* optional second [B][ld] summand, unweighted */,
                                    float* __restrict__ out) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= H) return;
  float s = 0.f;
  for (int b = threadIdx.x & 31; b < B; b += 32) {
    const float v = __ldg(x + (int64_t)b * ld + k);
    s += w ? __ldg(w + b) * v : v;
    if (x2) s += __ldg(x2 + (int64_t)b * ld + k);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) out[k] = s;
}

struct AdamSeg {
  float* p;
  int64_t off, n;
};
struct AdamTable {
  AdamSeg seg[2 * DPI_MAX_LINEAR];
  int n_seg;
};

// torch.optim.Adam (no weight decay, no amsgrad) over all parameter tensors in one launch
__global__ void train_adam_kernel(const AdamTable T, const float* __restrict__ grad, float* __restrict__ m,
                                  float* __restrict__ v, float lr, float b1, float b2, float eps, float bc1, float bc2,
                                  int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int s = 0;
  while (s + 1 < T.n_seg && i >= T.seg[s + 1].off) ++s;
  float* p = T.seg[s].p + (i - T.seg[s].off);
  const float gi = grad[i];
  const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
  const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
  m[i] = mi;
  v[i] = vi;
  *p -= (lr / bc1) * mi / (sqrtf(vi / bc2) + eps);
}

// c[b][k] = e[b][k] * w[k]
__global__ void train_bcast_mul_kernel(const float* __restrict__ e, const float* __restrict__ w, int B, int H,
                                       float* __restrict__ c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  c[i] = __ldg(e + i) * __ldg(w + (i % H));
}

// Adam with the (1-based) step read from device memory, for replay from a CUDA graph; train_step_inc_kernel advances it
__global__ void train_adam_dev_kernel(const AdamTable T, const float* __restrict__ grad, float* __restrict__ m,
                                      float* __restrict__ v, float lr, float b1, float b2, float eps,
                                      const long long* __restrict__ step, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float st = (float)(*step);
  const float bc1 = 1.f - powf(b1, st), bc2 = 1.f - powf(b2, st);
  int s = 0;
  while (s + 1 < T.n_seg && i >= T.seg[s + 1].off) ++s;
  float* p = T.seg[s].p + (i - T.seg[s].off);
  const float gi = grad[i];
  const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
  const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
  m[i] = mi;
  v[i] = vi;
  *p -= (lr / bc1) * mi / (sqrtf(vi / bc2) + eps);
}
__global__ void train_step_inc_kernel(long long* step) { *step += 1; }

// ---- data-parallel step over NVLink peer memory: gradient all-reduce fused with Adam -----------------------------
// Every rank's flat gradient lives in a cudaMalloc'd block that its peers have mapped (CUDA IPC).  One kernel per
// optimiser step and rank:
//   barrier 1  own flag <- 2e+1 (release, system scope); wait until every peer's flag >= 2e+1: all gradients written
//   reduce     g_i = (sum over ranks r = 0..W-1, in rank order, of grad_r[i]) / W, read through NVLink (ld.cg: remote
//              lines are never cached in the local L2) -- the same order on every rank, so the replicas stay BITWISE
//              identical without a broadcast -- followed by the Adam update of the local replica
//   barrier 2  last block: own flag <- 2e+2; wait for the peers' flags >= 2e+2: nobody still reads this rank's
//              gradient when the next step's kernels overwrite it
// e is a device-side launch counter (never reset, identical on every rank because all ranks run the same schedule),
// so the kernel replays from a CUDA graph.  Waits trap after ~4 s instead of hanging the box.
struct DpPeers {
  const float* grad[DPI_DP_MAX_WORLD];
  int* flag[DPI_DP_MAX_WORLD];
  int world, rank;
};
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void dp_wait_peers(const DpPeers& D, int target) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int r = 0; r < D.world; ++r) {
    if (r == D.rank) continue;
    unsigned spins = 0;
    while (ld_acquire_sys(D.flag[r]) - target < 0) {  // wrap-safe comparison of launch counters
      if ((++spins & 1023u) == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) __trap();
      }
    }
  }
}
__global__ void __launch_bounds__(256)
train_dp_reduce_adam_kernel(const AdamTable T, const DpPeers D, float* __restrict__ m, float* __restrict__ v, float lr,
                            float b1, float b2, float eps, const long long* __restrict__ step, int* __restrict__ epoch,
                            int* __restrict__ done, int64_t total) {
  const int e = *epoch;  // written only by the last block of the previous launch
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      __threadfence_system();
      st_release_sys(D.flag[D.rank], 2 * e + 1);
    }
    dp_wait_peers(D, 2 * e + 1);
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) {
    float gi = 0.f;
    for (int r = 0; r < D.world; ++r) gi += __ldcg(D.grad[r] + i);
    gi *= 1.f / (float)D.world;
    const float st = (float)(*step);
    const float bc1 = 1.f - powf(b1, st), bc2 = 1.f - powf(b2, st);
    int s = 0;
    while (s + 1 < T.n_seg && i >= T.seg[s + 1].off) ++s;
    float* p = T.seg[s].p + (i - T.seg[s].off);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    *p -= (lr / bc1) * mi / (sqrtf(vi / bc2) + eps);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(done, 1) == (int)gridDim.x - 1) {  // last block of this rank: every read of the peers' gradients is done
      *done = 0;
      __threadfence_system();
      st_release_sys(D.flag[D.rank], 2 * e + 2);
      dp_wait_peers(D, 2 * e + 2);
      *epoch = e + 1;
    }
  }
}

}  // namespace

int64_t train_param_count(const dpi_linear* L, int n_lin) {
  int64_t n = 0;
  for (int l = 0; l < n_lin; ++l) n += (int64_t)L[l].n_out * L[l].n_in + L[l].n_out;
  return n;
}

int64_t train_scratch_floats(const dpi_linear* L, int n_lin, int64_t B) {
  int hmax = 1 + L[0].n_in;
  for (int l = 0; l < n_lin - 1; ++l) hmax = std::max(hmax, L[l].n_out);
  // per layer: stacked activations [2B][H], ELU' [B][H], tangent pre-activations [B][H]; two stacked adjoint
  // buffers; stacked input [2B][1+d]; g [B][1+d]; u, a [B]
  return (int64_t)n_lin * 4 * B * hmax + (1 + (int64_t)n_lin) * 2 * B * hmax + 3 * B * (int64_t)(L[0].n_in) + 4 * B +
         16 * (int64_t)hmax * hmax + 64 * n_lin + 256;
}

// grad_flat [n_params] in torch parameters() order (W_0, b_0, W_1, b_1, ...); losses [3] = total, value, gradient
int launch_train_grad(const dpi_linear* L, int n_lin, const float* tx, const float* y, int64_t B64, float beta, float fw,
                      float* scratch, float* grad_flat, float* losses, cudaStream_t st, int64_t* launches,
                      const TrainAux* aux) {
  const int B = (int)B64;
  const int d1 = L[0].n_in;  // 1 + d
  const int nh = n_lin - 1;  // hidden (ELU) layers; layer nh is the scalar output layer
  int hmax = d1;
  for (int l = 0; l < nh; ++l) hmax = std::max(hmax, L[l].n_out);
  float* p = scratch;
  auto take = [&](int64_t n) {
    float* q = p;
    p += ((n + 3) / 4) * 4;
    return q;
  };
  float* HB[DPI_MAX_LINEAR + 1];  // HB[0] = [tx; r], HB[l] = [h_l; hd_l]  (value rows, then tangent rows)
  float *E[DPI_MAX_LINEAR + 1], *ZD[DPI_MAX_LINEAR + 1];
  HB[0] = take(2 * (int64_t)B * d1);
  for (int l = 1; l <= nh; ++l) {
    HB[l] = take(2 * (int64_t)B * hmax);
    E[l] = take((int64_t)B * hmax);
    ZD[l] = take((int64_t)B * hmax);
  }
  float* adjA = take(2 * (int64_t)B * hmax);
  float* adjZ[DPI_MAX_LINEAR];  // one (zbar, zdbar) buffer per layer: the weight-gradient branch reads it concurrently
  for (int l = 0; l < nh; ++l) adjZ[l] = take(2 * (int64_t)B * hmax);
  float* adjB = adjZ[0];
  float* g = take((int64_t)B * d1);
  float* a = take(B);
  float* rowl = take(2 * (int64_t)B);  // per-row loss terms
  float* ws = take(16 * (int64_t)hmax * hmax);  // split-K partials of the weight-gradient GEMMs
  int rc;
  auto wd = [&](int l) { return l == 0 ? d1 : L[l - 1].n_out; };  // width of HB[l]
  auto ew = [&](int threads_total) { return (unsigned)((threads_total + 255) / 256); };
  // offsets of the parameter tensors in the flat gradient
  int64_t offW[DPI_MAX_LINEAR], offb[DPI_MAX_LINEAR], off = 0;
  for (int l = 0; l < n_lin; ++l) {
    offW[l] = off, off += (int64_t)L[l].n_out * L[l].n_in;
    offb[l] = off, off += L[l].n_out;
  }

  // Everything that only feeds the gradient buffers or the reported losses runs on a side stream (a parallel branch
  // of the captured graph): the step is bound by its dependent chain of small kernels, not by their work.
  cudaStream_t sw = (aux && aux->side) ? aux->side : st;
  const bool forked = sw != st;
  auto fork_at = [&](cudaEvent_t ev) -> int {
    if (!forked) return DPI_OK;
    DPI_CUDA_CHECK(cudaEventRecord(ev, st));
    DPI_CUDA_CHECK(cudaStreamWaitEvent(sw, ev, 0));
    return DPI_OK;
  };
  // ---- value forward (keeps ELU'); the scalar output layer is evaluated inside the loss kernel ----
  // HB[0]'s value rows (a copy of tx) are only read by the first layer's weight gradient at the very end
  if ((rc = fork_at(forked ? aux->pre[0] : nullptr))) return rc;
  DPI_CUDA_CHECK(cudaMemcpyAsync(HB[0], tx, sizeof(float) * (size_t)B * d1, cudaMemcpyDeviceToDevice, sw));
  for (int l = 0; l < nh; ++l) {
    GemmArgs G{};
    G.A = (l == 0) ? tx : HB[l], G.lda = wd(l), G.B = L[l].W, G.ldb = L[l].n_in, G.bias = L[l].b;
    G.C = HB[l + 1], G.ldc = L[l].n_out, G.D = E[l + 1], G.ldd = L[l].n_out;
    G.R = B, G.N = L[l].n_out, G.K = L[l].n_in, G.epi = kEpiBiasEluD;
    if ((rc = gemm(G, st, launches))) return rc;
  }
  // ---- input gradient by one reverse sweep: c = E_nh . w_out ; c = (c W_l) . E_l ; g = c W_0 ----
  {
    const int H = L[nh].n_in;
    train_bcast_mul_kernel<<<ew(B * H), 256, 0, st>>>(E[nh], L[nh].W, B, H, adjA);
    DPI_CUDA_CHECK(cudaGetLastError());
    *launches += 1;
    const float* c = adjA;
    float* nxt = adjB;
    for (int l = nh - 1; l >= 1; --l) {  // through Linear l (n_out = wd(l+1), n_in = wd(l)), then ELU' of layer l
      GemmArgs G{};
      G.A = c, G.lda = L[l].n_out, G.B = L[l].W, G.ldb = L[l].n_in, G.b_kn = 1, G.E = E[l], G.lde = L[l].n_in;
      G.C = nxt, G.ldc = L[l].n_in, G.R = B, G.N = L[l].n_in, G.K = L[l].n_out, G.epi = kEpiMulE;
      if ((rc = gemm(G, st, launches))) return rc;
      c = nxt;
      nxt = (nxt == adjA) ? adjB : adjA;
    }
    GemmArgs G{};
    G.A = c, G.lda = L[0].n_out, G.B = L[0].W, G.ldb = d1, G.b_kn = 1;
    G.C = g, G.ldc = d1, G.R = B, G.N = d1, G.K = L[0].n_out, G.epi = kEpiNone;
    if ((rc = gemm(G, st, launches))) return rc;
  }
  // ---- loss, a = dL/du, r = dL/d(grad_x u) (tangent rows of HB[0]) ----
  train_loss_kernel<<<(B + 7) / 8, 256, 0, st>>>(tx, y, HB[nh], L[nh].W, L[nh].b, L[nh].n_in, g, B, d1 - 1, beta, fw, a,
                                                 HB[0] + (int64_t)B * d1, rowl);
  if ((rc = fork_at(forked ? aux->pre[1] : nullptr))) return rc;
  train_loss_finish_kernel<<<1, 256, 0, sw>>>(rowl, B, fw, losses);
  DPI_CUDA_CHECK(cudaGetLastError());
  *launches += 2;
  // ---- tangent forward along r: zd_l = hd_{l-1} W_l^T (kept), hd_l = ELU'(z_l) . zd_l ----
  for (int l = 0; l < nh; ++l) {
    GemmArgs G{};
    G.A = HB[l] + (int64_t)B * wd(l), G.lda = wd(l), G.B = L[l].W, G.ldb = L[l].n_in;
    G.E = E[l + 1], G.lde = L[l].n_out;
    G.C = ZD[l + 1], G.ldc = L[l].n_out, G.D = HB[l + 1] + (int64_t)B * L[l].n_out, G.ldd = L[l].n_out;
    G.R = B, G.N = L[l].n_out, G.K = L[l].n_in, G.epi = kEpiKeepMulE;
    if ((rc = gemm(G, st, launches))) return rc;
  }
  // ---- output layer gradients: dW_out = sum_b a_b h_nh + sum_b hd_nh ; db_out = sum_b a_b ----
  {
    const int H = L[nh].n_in;
    if ((rc = fork_at(forked ? aux->pre[2] : nullptr))) return rc;
    train_colsum_kernel<<<(H + 7) / 8, 256, 0, sw>>>(HB[nh], H, a, B, H, HB[nh] + (int64_t)B * H, grad_flat + offW[nh]);
    train_colsum_kernel<<<1, 256, 0, sw>>>(a, 1, nullptr, B, 1, nullptr, grad_flat + offb[nh]);
    DPI_CUDA_CHECK(cudaGetLastError());
    *launches += 2;
  }
  // ---- reverse sweep through the (value, tangent) hidden layers ----
  // The main chain is adjoint kernel -> (hbar, hdbar) GEMM per layer; the weight gradients (split-K GEMM, its
  // reduction, the bias column sums) only read what the chain has already produced, so they run on a side stream
  // (a parallel branch of the captured graph): the step is bound by the dependent chain of small kernels.
  {
    float* hb = adjA;   // [2B][wd(l+1)] adjoints of (h_{l+1}, hd_{l+1})
    for (int l = nh - 1; l >= 0; --l) {
      const int Ho = L[l].n_out, Hi = L[l].n_in;
      float* zb = adjZ[l];  // [2B][wd(l+1)] adjoints of (z_{l+1}, zd_{l+1})
      if (l == nh - 1)  // top adjoints hbar = a w_out (value rows), hdbar = w_out (tangent rows) formed on the fly
        train_adjoint_top_kernel<<<ew(B * Ho), 256, 0, st>>>(a, L[nh].W, E[l + 1], ZD[l + 1], B, Ho, zb);
      else
        train_adjoint_kernel<<<ew(B * Ho), 256, 0, st>>>(hb, E[l + 1], ZD[l + 1], B, Ho, zb);
      DPI_CUDA_CHECK(cudaGetLastError());
      *launches += 1;
      if (forked) {
        DPI_CUDA_CHECK(cudaEventRecord(aux->fork[l], st));
        DPI_CUDA_CHECK(cudaStreamWaitEvent(sw, aux->fork[l], 0));
      }
      // dW_l [Ho][Hi] = [zbar; zdbar]^T [h_l; hd_l]   (A stored [K = 2B][M = Ho])
      GemmArgs G{};
      G.A = zb, G.lda = Ho, G.a_km = 1, G.B = HB[l], G.ldb = Hi, G.b_kn = 1;
      G.C = grad_flat + offW[l], G.ldc = Hi, G.R = Ho, G.N = Hi, G.K = 2 * B, G.epi = kEpiNone;
      G.ws = ws, G.ws_splits = 16;
      if ((rc = gemm(G, sw, launches))) return rc;
      train_colsum_kernel<<<(Ho + 7) / 8, 256, 0, sw>>>(zb, Ho, nullptr, B, Ho, nullptr, grad_flat + offb[l]);
      DPI_CUDA_CHECK(cudaGetLastError());
      *launches += 1;
      if (l > 0) {
        // (hbar, hdbar)_l = (zbar, zdbar)_{l+1} W_l : one GEMM over the 2B stacked rows
        GemmArgs H{};
        H.A = zb, H.lda = Ho, H.B = L[l].W, H.ldb = Hi, H.b_kn = 1;
        H.C = hb, H.ldc = Hi, H.R = 2 * (int64_t)B, H.N = Hi, H.K = Ho, H.epi = kEpiNone;
        if ((rc = gemm(H, st, launches))) return rc;
      }
    }
    if (forked) {
      DPI_CUDA_CHECK(cudaEventRecord(aux->join, sw));
      DPI_CUDA_CHECK(cudaStreamWaitEvent(st, aux->join, 0));
    }
  }
  return DPI_OK;
}

int launch_train_adam(const dpi_linear* L, int n_lin, const float* grad_flat, float* m, float* v, float lr, float b1,
                      float b2, float eps, int64_t step, cudaStream_t st, int64_t* launches) {
  AdamTable T{};
  int64_t off = 0;
  T.n_seg = 0;
  for (int l = 0; l < n_lin; ++l) {
    T.seg[T.n_seg++] = AdamSeg{const_cast<float*>(L[l].W), off, (int64_t)L[l].n_out * L[l].n_in};
    off += (int64_t)L[l].n_out * L[l].n_in;
    T.seg[T.n_seg++] = AdamSeg{const_cast<float*>(L[l].b), off, (int64_t)L[l].n_out};
    off += L[l].n_out;
  }
  const float bc1 = 1.f - (float)std::pow((double)b1, (double)step), bc2 = 1.f - (float)std::pow((double)b2, (double)step);
  train_adam_kernel<<<(unsigned)((off + 255) / 256), 256, 0, st>>>(T, grad_flat, m, v, lr, b1, b2, eps, bc1, bc2, off);
  DPI_CUDA_CHECK(cudaGetLastError());
  *launches += 1;
  return DPI_OK;
}

int launch_train_adam_dev(const dpi_linear* L, int n_lin, const float* grad_flat, float* m, float* v, float lr, float b1,
                          float b2, float eps, long long* step_dev, cudaStream_t st, int64_t* launches) {
  AdamTable T{};
  int64_t off = 0;
  T.n_seg = 0;
  for (int l = 0; l < n_lin; ++l) {
    T.seg[T.n_seg++] = AdamSeg{const_cast<float*>(L[l].W), off, (int64_t)L[l].n_out * L[l].n_in};
    off += (int64_t)L[l].n_out * L[l].n_in;
    T.seg[T.n_seg++] = AdamSeg{const_cast<float*>(L[l].b), off, (int64_t)L[l].n_out};
    off += L[l].n_out;
  }
  train_adam_dev_kernel<<<(unsigned)((off + 255) / 256), 256, 0, st>>>(T, grad_flat, m, v, lr, b1, b2, eps, step_dev, off);
  train_step_inc_kernel<<<1, 1, 0, st>>>(step_dev);
  DPI_CUDA_CHECK(cudaGetLastError());
  *launches += 2;
  return DPI_OK;
}

int launch_train_dp_reduce_adam(const dpi_linear* L, int n_lin, const DpLinks& dp, float* m, float* v, float lr, float b1,
                                float b2, float eps, long long* step_dev, cudaStream_t st, int64_t* launches) {
  AdamTable T{};
  int64_t off = 0;
  T.n_seg = 0;
  for (int l = 0; l < n_lin; ++l) {
    T.seg[T.n_seg++] = AdamSeg{const_cast<float*>(L[l].W), off, (int64_t)L[l].n_out * L[l].n_in};
    off += (int64_t)L[l].n_out * L[l].n_in;
    T.seg[T.n_seg++] = AdamSeg{const_cast<float*>(L[l].b), off, (int64_t)L[l].n_out};
    off += L[l].n_out;
  }
  DpPeers D{};
  D.world = dp.world, D.rank = dp.rank;
  for (int r = 0; r < dp.world; ++r) D.grad[r] = dp.grad[r], D.flag[r] = dp.flag[r];
  train_dp_reduce_adam_kernel<<<(unsigned)((off + 255) / 256), 256, 0, st>>>(T, D, m, v, lr, b1, b2, eps, step_dev, dp.epoch,
                                                                          dp.done, off);
  train_step_inc_kernel<<<1, 1, 0, st>>>(step_dev);
  DPI_CUDA_CHECK(cudaGetLastError());
  *launches += 2;
  return DPI_OK;
}

}  // namespace dpi
