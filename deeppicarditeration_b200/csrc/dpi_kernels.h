// Internal launch interface between the C ABI (dpi_api.cu) and the kernel translation units.
#pragma once
#include <vector>

#include "dpi_common.cuh"

namespace dpi {

// ---- fp32 packed MLP (SIMT path): k-major, zero-padded to HP columns --------------------------------
struct MlpDev {
  int n_linear;  // L: Linear(1+nx,H) , L-2 x Linear(H,H), Linear(H,1)
  int nx;
  int HP;     // padded hidden width: 64 or 128
  int d_pad;  // nx rounded up to a multiple of 128
  const float* Wt1;   // [d_pad][HP]  Wt1[j][h] = W1[h][1+j]
  const float* wt;    // [HP]         W1[:,0]   (time column)
  const float* b1;    // [HP]
  const float* wsum;  // [HP]         sum_j W1[:,1+j]  == tangent of layer 1 along (0,1,...,1)
  const float* Wth[DPI_MAX_LINEAR];  // hidden layer l: [HP][HP]  Wth[k][h] = W_l[h][k]
  const float* Wth_bf[DPI_MAX_LINEAR];  // same, every entry rounded to fp16 (control variate of the tensor-core path)
  const float* bh[DPI_MAX_LINEAR];   // [HP]
  const float* wout;  // [HP]
  const float* bout;  // [1] device
};

// n_factors = 0: t = (T - 2 eps)(1 - U) + eps (data.py:161-167); n_factors >= 1: t = T (1 - prod of n_factors uniforms) (data.py:149-159)
int launch_sample_tx(const ProblemDev& P, int64_t n, int64_t base, uint64_t seed, float eps, int n_factors, float* tx,
                     cudaStream_t st);
// u_td (nullable): [n * M_T] values u_k(t + delta, X) for the TD estimator (used by the points whose step ends before T)
int launch_terminal(const ProblemDev& P, const float* tx, int64_t n, int M_T, uint64_t seed, int64_t base,
                    const float* zT, float* y, cudaStream_t st, const float* u_td = nullptr);
// rows [np * M_T][1 + nx] = (t_next, x + sqrt(t_next - t) sigma Z): the inputs of those values
int launch_td_rows(const ProblemDev& P, const float* tx, int64_t np, int M_T, uint64_t seed, int64_t base, const float* zT,
                   float* rows, cudaStream_t st);

// Pack torch-layout weights into MlpDev buffers (buffers preallocated by the handle; `bout` read back by caller).
int launch_pack_mlp(const dpi_network& net, int nx, int HP, int d_pad, float* Wt1, float* wt, float* b1, float* wsum,
                    float* const* Wth, float* const* Wth_bf, float* const* bh, float* wout, float* bout_dev,
                    cudaStream_t st);

// Burgers (Cha) integral estimator, fp32 SIMT.  base pass: fb[n], c0[n][HP]; path pass: y += integral part, then clip.
// emulate_bf16: hidden layers use the bf16-rounded weights and round the activations between layers, i.e. the
// arithmetic of the tensor-core path, so that f_m - f_b cancels its rounding error as s -> t.
int launch_cha_simt_base(const ProblemDev& P, const MlpDev& net, const float* tx, int64_t n, float* fb, float* c0,
                         float* u_out, float* dsum_out, bool emulate_bf16, cudaStream_t st);
int launch_cha_simt_path(const ProblemDev& P, const MlpDev& net, const float* tx, int64_t n, int M_I, uint64_t seed,
                         int64_t base, const float* u_s, const float* zS, const float* fb, const float* c0,
                         float bound, float* y, cudaStream_t st);

// ---- fully nonlinear equation with SDGD (dpi_gbm_simt.cu): y += integral part, then clip ------------------
int launch_gbm_simt(const ProblemDev& P, const MlpDev& net, bool zero_net, const float* tx, int64_t n, int M_I,
                    uint64_t seed, int64_t base, const float* u_s, const float* zS, const int32_t* idx, float bound,
                    float* y, cudaStream_t st);

// Hessian diagonal [n, nx] of the MLP at the rows of tx (hidden width <= 64, nx <= 128)
int launch_mlp_hess_diag(const ProblemDev& P, const MlpDev& net, const float* tx, int64_t n, float* diag, cudaStream_t st);
// the full Hessian [n][nx][nx] (coordinate and pair directions, polarisation)
int launch_mlp_hess_full(const ProblemDev& P, const MlpDev& net, const float* tx, int64_t n, float* H, cudaStream_t st);

// same on the tensor cores (dpi_gbm_tc.cu): fp16 jets through the hidden layers, nx <= 127, v <= 127
int launch_gbm_tc(const ProblemDev& P, const MlpDev& net, const float* tx, int64_t n, int M_I, uint64_t seed, int64_t base,
                  const float* u_s, const float* zS, const int32_t* idx, float bound, float* y, cudaStream_t st);

// warp-specialised version (dpi_gbm_tc2.cu): value chain on a prep team, two path rows in flight; nx, v <= 128, <= 3 hidden GEMM layers
bool gbm_tc2_eligible(const ProblemDev& P, const MlpDev& net);
int launch_gbm_tc2(const ProblemDev& P, const MlpDev& net, const float* tx, int64_t n, int M_I, uint64_t seed, int64_t base,
                   const float* u_s, const float* zS, const int32_t* idx, float bound, float* y, cudaStream_t st);

// directions in natural order, value-chain weights in registers, four path rows in flight (dpi_gbm_tc3.cu); nx, v <= 128, 4 v >= nx
bool gbm_tc3_eligible(const ProblemDev& P, const MlpDev& net);
int launch_gbm_tc3(const ProblemDev& P, const MlpDev& net, const float* tx, int64_t n, int M_I, uint64_t seed, int64_t base,
                   const float* u_s, const float* zS, const int32_t* idx, float bound, float* y, cudaStream_t st);

// ---- 16-bit tcgen05 path (dpi_mlp_tc.cu): fp16 operands, fp32 accumulation ---------------------------------------------------------------
#ifndef DPI_TC_GROUP_TILES
#define DPI_TC_GROUP_TILES 16
#endif
// 128-sample tiles per work unit (fixed: the summation order must not depend on n).  8, 16 and 32 run equally fast at
// the bench size; 16 halves the unit-partial traffic of 8 and still gives 512 points (one of eight GPUs) 7 units per SM.
constexpr int kTcGroupTiles = DPI_TC_GROUP_TILES;
// per-column fp32 constants of the network, passed to the kernel by value (constant bank)
struct alignas(16) TcConsts {
  float wt[128];            // log2e W1[:,0]                time column (fp32: multiplies the per-row time)
  uint32_t wt16[64];        // the same as fp16x2 pairs (fp16-accumulator path)
  uint32_t wsum16[64];      // fp16x2 pairs: log2e sum_j W1[:,1+j]   first-layer tangent along (0,1,...,1)
  uint32_t b16[3][64];      // fp16x2 pairs: log2e biases of the three hidden GEMM layers
  uint32_t wout16[64];      // fp16x2 pairs: ln2 w_out
  float bout;
  float pad[3];
};
struct TcNetDev {
  int nchunks = 1;  // first-layer K chunks of 128 input dims (d <= 1024)
  const void* W;  // 4 x [128 x 128] fp16 operand images (K-major, SWIZZLE_128B): W1 x-columns, W2, W3, W4
  TcConsts consts;
  int nx, ksteps1, no, cls, num_sms;
};
int tc_pack_network(const dpi_network& net, const ProblemDev& P, const MlpDev& M, std::vector<void*>& owner,
                    TcNetDev* out, bool* ready, int64_t* launches, cudaStream_t st);
int64_t tc_units_per_point(int M_I);
constexpr int kTcMaxClasses = 32;  // row classes of the producer mapping (partial sums per unit)
constexpr int64_t kTcPartGFloatsPerUnit = kTcMaxClasses * 128;
constexpr int64_t kTcPartFFloatsPerUnit = 8;  // [slot][lane quarter]
// partG [n*units][32][128], partF [n*units][8]: scratch for the unit partial sums
int launch_cha_tc_path(const ProblemDev& P, const TcNetDev& net, const float* tx, int64_t n, int M_I, uint64_t seed,
                       int64_t base, const float* u_s, const float* zS, const float* fb, const float* c0, float bound,
                       float* partG, float* partF, int* unit_counter, void* zring, int* nonfinite, float* y,
                       cudaStream_t st);
// bytes of the increment ring of the chunked (d > 128) tensor-core kernel: 4 tiles x chunks x 32 KB per SM
inline size_t tc_zring_bytes(const TcNetDev& net) { return net.nchunks > 1 ? (size_t)net.num_sms * 4 * net.nchunks * 32768 : 0; }
// D[128x128] = A[128x128] . B[128x128]^T (fp32 row-major device buffers, rounded to bf16) through tcgen05
int tc_umma_selftest(const float* A, const float* B, float* D, cudaStream_t st);

// ---- generic fp32 GEMM with fused epilogues (dpi_hjb.cu): C[R,N] = epilogue(A . B) ---------------------------
// A [R][K] row-major (or [K][R] when a_km), B = W [N][K] (b_kn = 0: y = x W^T) or W [K][N] (b_kn = 1: c W)
enum GemmEpi : int {
  kEpiNone = 0,      // C = acc
  kEpiBias = 1,      // C = acc + bias
  kEpiBiasElu = 2,   // C = ELU(acc + bias)
  kEpiBiasEluD = 3,  // C = ELU(acc + bias), D = ELU'(acc + bias)
  kEpiMulE = 4,      // C = acc * E        (VJP through an ELU: cotangent times the stored ELU')
  kEpiKeepMulE = 5,  // C = acc, D = acc * E   (JVP through an ELU, keeping the pre-activation tangent)
  kEpiMulDElu = 6    // C = acc * ELU'(z) with E = the stored ACTIVATION h = ELU(z): ELU' = h > 0 ? 1 : h + 1
};
struct GemmArgs {
  const float* A;
  int lda;
  const float* B;
  int ldb;
  int b_kn;
  const float* bias;
  const float* E;
  int lde;
  float* C;
  int ldc;
  float* D;
  int ldd;
  int64_t R;
  int N, K, epi;
  int a_km;
  float* ws;      // optional split-K workspace [ws_splits][R][N] (epi == kEpiNone only)
  int ws_splits;
  int tf32;       // 1: use the TF32 tcgen05 GEMM when the shape is eligible
};
int gemm(const GemmArgs& g, cudaStream_t st, int64_t* launches);
// TF32 tcgen05 GEMM (dpi_gemm_tc.cu): same arguments and epilogues; K % 32 == 0, no a_km, R >= 128, N >= 64
bool gemm_tf32_eligible(const GemmArgs& g);
int gemm_tf32(const GemmArgs& g, cudaStream_t st, int64_t* launches);
// WT [cols][rows] = W [rows][cols]^T
int gemm_transpose(const float* W, int rows, int cols, float* WT, cudaStream_t st);

// ---- HJB / PISGradNet, fp32 batched-GEMM pipeline (dpi_hjb.cu) -------------------------------------------
// Weight pointers are the caller's torch-layout tensors (W [n_out, n_in] row-major), used in place.
struct PisDev {
  int n_main, n_t_encoder, n_smooth;
  int hmax;  // widest nn_module hidden layer
  dpi_linear main[DPI_MAX_LINEAR], t_encoder[DPI_MAX_LINEAR], smooth_net[DPI_MAX_LINEAR];
  const float* phase;  // [64]
  const float* coeff;  // [64]
  int tf32;            // 1: eligible GEMMs on the tensor cores (TF32 operands, fp32 accumulate; dpi_gemm_tc.cu)
  const float* main_t[DPI_MAX_LINEAR];  // main[l].W transposed ([n_in][n_out]): K-major operands for the VJP GEMMs
};
int64_t hjb_scratch_floats(const PisDev& net, int d, int64_t rows, int64_t points);
int64_t hjb_points_per_chunk(const PisDev& net, int d, int M_I, int64_t budget_bytes);
// y += integral part (paths + baseline row per point through the network), then clip
int launch_hjb_paths(const ProblemDev& P, const PisDev& net, const float* tx, int64_t n, int M_I, uint64_t seed,
                     int64_t base, const float* u_s, const float* zS, float bound, float* scratch,
                     int64_t points_per_chunk, float* y, cudaStream_t st, int64_t* launches);
// u [n] and grad_x u [n, d] at the rows of tx (either may be null)
int launch_hjb_eval(const ProblemDev& P, const PisDev& net, const float* tx, int64_t n, float* scratch,
                    int64_t rows_per_chunk, float* u, float* du, cudaStream_t st, int64_t* launches);

// ---- fused tcgen05 PISGradNet path (dpi_hjb_tc.cu): fp16 operands, activations on chip, weights streamed from L2 ----
struct HjbTcNet;
// packs the network when its shape qualifies (equal hidden widths of 128, 256 or 512; d <= 128); *out = null otherwise
int hjb_tc_pack(const PisDev& net, int d, std::vector<void*>& owner, HjbTcNet** out, int64_t* launches, cudaStream_t st);
void hjb_tc_free(HjbTcNet* net);
int64_t hjb_tc_part_floats(int64_t n, int M_I, int d);
// y += integral part, then clip; fb [n] and part [hjb_tc_part_floats] are scratch
int launch_hjb_tc_paths(const ProblemDev& P, const HjbTcNet* net, const float* tx, int64_t n, int M_I, uint64_t seed,
                        int64_t base, const float* u_s, const float* zS, float bound, float* fb, float* part,
                        int* nonfinite, float* y, cudaStream_t st, int64_t* launches);

// plain MLP: u [n] and grad_x u [n, d] (either may be null); weights in torch layout, used in place
int64_t mlp_eval_scratch_floats(const dpi_linear* L, int n_lin, int64_t rows);
int launch_mlp_eval(const dpi_linear* L, int n_lin, int d, const float* tx, int64_t n, float* scratch,
                    int64_t rows_per_chunk, float* u, float* du, cudaStream_t st, int64_t* launches);

// ---- regression step for a scalar-output MLP (dpi_train.cu) ---------------------------------------------------
int64_t train_param_count(const dpi_linear* L, int n_lin);
int64_t train_scratch_floats(const dpi_linear* L, int n_lin, int64_t B);
// Side stream and events for the weight-gradient branch of the reverse sweep (fork after each adjoint kernel, one
// join at the end); null side stream = everything on `st`.
struct TrainAux {
  cudaStream_t side = nullptr;
  cudaEvent_t fork[DPI_MAX_LINEAR] = {};
  cudaEvent_t pre[3] = {};  // start of the step, loss terms written, tangent forward done
  cudaEvent_t join = nullptr;
};
int launch_train_grad(const dpi_linear* L, int n_lin, const float* tx, const float* y, int64_t B, float beta, float fw,
                      float* scratch, float* grad_flat, float* losses, cudaStream_t st, int64_t* launches,
                      const TrainAux* aux = nullptr);
int launch_train_adam(const dpi_linear* L, int n_lin, const float* grad_flat, float* m, float* v, float lr, float b1,
                      float b2, float eps, int64_t step, cudaStream_t st, int64_t* launches);

int launch_train_adam_dev(const dpi_linear* L, int n_lin, const float* grad_flat, float* m, float* v, float lr, float b1,
                          float b2, float eps, long long* step_dev, cudaStream_t st, int64_t* launches);
// data-parallel ranks of one node: every rank's gradient block mapped through CUDA IPC (own block at index `rank`)
constexpr int DPI_DP_MAX_WORLD = 16;
struct DpLinks {
  int world = 1, rank = 0;
  const float* grad[DPI_DP_MAX_WORLD] = {};
  int* flag[DPI_DP_MAX_WORLD] = {};
  int* epoch = nullptr;  // device launch counter of the fused reduce + Adam kernel
  int* done = nullptr;   // device block counter
  void* mapped[DPI_DP_MAX_WORLD] = {};  // peer base pointers to cudaIpcCloseMemHandle
};
// gradient all-reduce over NVLink peer memory fused with the Adam update (one kernel per step, graph-capturable)
int launch_train_dp_reduce_adam(const dpi_linear* L, int n_lin, const DpLinks& dp, float* m, float* v, float lr, float b1,
                                float b2, float eps, long long* step_dev, cudaStream_t st, int64_t* launches);

// ---- draws and status (dpi_draws.cu) --------------------------------------------------------------------------
// the Philox draws of a call in the reference's layout (any pointer may be null = skip)
int launch_export_draws(const ProblemDev& P, int64_t n, int M_T, int M_I, uint64_t seed, int64_t base, float* zT,
                        float* u_s, float* zS, int32_t* idx, cudaStream_t st, int64_t* launches);
// *flag |= 1 if any of y[0..count) is inf / NaN
int launch_check_finite(const float* y, int64_t count, int* flag, cudaStream_t st);

// y = clip(y + add, +-bound) elementwise helpers used by the Zero-network shortcuts
int launch_finish_zero(const ProblemDev& P, const float* tx, int64_t n, float bound, float* y, cudaStream_t st);

}  // namespace dpi
