/*
 * dpi_b200.h — C ABI of the B200-native Deep-Picard-Iteration label generator.
 *
 * This is the drop-in boundary for ONE hot path of the reference (DeepOptimalControl/DeepPicardIteration):
 * the per-iteration Monte Carlo label generator.  The reference has no FFI — the path sits behind a
 * Python object protocol — so each entry point below cites the Python interface it replaces
 * (paths relative to the reference root):
 *
 *   dpi_sample_tx      <- _OnlineDataGenerator.sample_t_always_uniform   picard/data.py:161-167
 *                         + Equation.sample_x / sample_x0 / sample_x_ts   picard/equations.py:118-119,217-230,334-335,485-486,710-711
 *   dpi_label_gen      <- OnlineDataGenerator.generate_with_gradients     picard/data.py:1208-1218
 *                           = estimate_terminal_with_gradients            picard/data.py:899-926
 *                           + estimate_integral_with_gradients            picard/data.py:471-527
 *                           + get_f (u_k forward + derivative + f)        picard/data.py:1226-1325
 *                           + clip                                        picard/data.py:222
 *   dpi_label_gen_host <- OnlineDataGenerator.sample_with_gradients as seen by a host-memory caller
 *                         (IterableDatasetWithInternalBatch.refresh_buffer, picard/dataset.py:107-116)
 *   dpi_set_problem    <- Equation constructors: Cha / GBMEquationComplexExact / OUProcessEquation
 *                         picard/equations.py:283-289,401-419,504-558,610-620
 *   dpi_set_network    <- OnlineDataGenerator.__init__ freezing `solution`  picard/data.py:407-412
 *                         (construct_mlp / PISGradNet / ZeroSolution, picard/solution.py:123-135,138-289,330-337)
 *   dpi_net_eval       <- solution(tx) and autograd.grad(u.sum(), x)       picard/data.py:1230,1245-1251
 *                         (used by the evaluation callback, picard/utils.py:362-458)
 *
 * Conventions: plain pointers and sizes only (no torch types).  All functions return 0 on success or a
 * negative dpi_status; the message is available from dpi_last_error() (thread-local).  Nothing throws.
 * Device pointers are torch-owned; the library allocates only handle-internal scratch (packed weights,
 * per-point baseline scratch, staging for the *_host entry points).  Kernels are launched on the stream
 * passed in (a cudaStream_t cast to void*); calls are asynchronous like torch ops, except *_host variants
 * which synchronise the stream before returning.  One handle per device; a handle is not thread-safe.
 *
 * There is NO CPU fallback: every entry point fails with DPI_ERR_CUDA when no sm_100 device is present.
 */
#ifndef DPI_B200_H
#define DPI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPI_B200_VERSION 101 /* major*10000 + minor*100 + patch; 101: dpi_label_config.estimate_delta_t */
#define DPI_MAX_LINEAR 8     /* Linear layers per chain */

typedef struct dpi_handle dpi_handle;

typedef enum {
  DPI_OK = 0,
  DPI_ERR_INVALID = -1,     /* bad argument / unsupported combination */
  DPI_ERR_CUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
  DPI_ERR_UNSUPPORTED = -3, /* valid in the reference, not built yet in this library */
  DPI_ERR_STATE = -4,       /* problem / network not set */
  DPI_ERR_NONFINITE = -5    /* a label came out inf / NaN (dpi_check_labels, dpi_label_gen_host) */
} dpi_status;

typedef enum { DPI_EQ_CHA = 0, DPI_EQ_GBM = 1, DPI_EQ_OU = 2 } dpi_equation_kind;
typedef enum { DPI_NET_ZERO = 0, DPI_NET_MLP = 1, DPI_NET_PISGRAD = 2 } dpi_network_kind;
/* arithmetic of the u_k GEMM chain: fp32 FMA (labels within 1e-4 rel of the fp64 reference) or the 16-bit tcgen05
 * tensor-core path (within 1e-2 rel; measured 6e-5).  Accumulation: the Cha and GBM kernels keep fp16 accumulators in
 * TMEM for their hidden layers (K <= 128 per sum: 4.4e-4 from an fp32-accumulated product, against 2.8e-4 for the
 * rounding to fp16 that the next layer's operand needs anyway), fp32 for first layers that sum over more than 128
 * inputs and everywhere in the PISGradNet kernel.  The tensor-core operands are IEEE fp16,
 * not bf16: three more mantissa bits at the same operand width, and the bias/ELU/ELU' epilogue runs in packed
 * fp16x2.  Its range limit is |log2(e) x pre-activation| < 6.5e4, far above anything a trained u_k produces; a network
 * that does leave it yields inf / NaN, which the label kernels' final pass records in a status word:
 * dpi_label_gen_host returns DPI_ERR_NONFINITE and dpi_check_labels reports it for the asynchronous dpi_label_gen.
 * DPI_PREC_BF16 (bf16 operands: fp32's range, 8 mantissa bits) is NOT built: it is a distinct value that every entry
 * point rejects with DPI_ERR_UNSUPPORTED, so a caller asking for bf16's range is never silently given fp16's.
 * Which paths DPI_PREC_F16 selects: Cha with the MLP [1+d,128,128,128,128,1], d <= 1024 (chunked first layer above 128);
 * GBM/SDGD with hidden width 64, v <= 127 (fp16 2-jets); PISGradNet/OU with equal hidden widths of 128, 256
 * or 512 and nx <= 128: the fused kernel of dpi_hjb_tc.cu (fp16 operands, activations on chip, weights streamed from
 * L2).  DPI_PREC_TF32 (PISGradNet only): the layer-by-layer pipeline with its wide GEMMs on the tensor cores in TF32
 * (fp32 operands in memory, 10-bit mantissa in the multiply, fp32 accumulate; also what DPI_PREC_F16 falls back to for
 * PISGradNet shapes the fused kernel does not take).  Anything else: DPI_ERR_UNSUPPORTED. */
typedef enum { DPI_PREC_FP32 = 0, DPI_PREC_F16 = 1, DPI_PREC_BF16 = 2, DPI_PREC_TF32 = 3 } dpi_precision;

/* Equation parameters.  All array pointers are HOST pointers (tiny), copied by dpi_set_problem. */
typedef struct {
  int32_t kind; /* dpi_equation_kind */
  int32_t nx;
  float T;
  float alpha;
  /* Cha (equations.py:283-289): k as configured (the library divides by sqrt(nx)) */
  float cha_k;
  /* GBMEquationComplexExact (equations.py:401-419): w [nodes, 1+nx] row-major, v [nodes] */
  int32_t gbm_nodes;
  const double* gbm_w;
  const double* gbm_v;
  /* OUProcessEquation (equations.py:504-558,610-620): mean [K,nx], diagvar [K,nx], pi [K] */
  int32_t ou_K;
  const double* ou_mean;
  const double* ou_diagvar;
  const double* ou_pi;
  float ou_theta;
  float ou_mu;
  float ou_alpha_init; /* alpha_scale * alpha: variance of x0 ~ N(0, alpha_init I) */
  /* SDGD (data.py:497-502): number of sampled Hessian-diagonal coordinates per path; 0 = off */
  int32_t sdgd_v;
} dpi_problem;

/* One torch.nn.Linear: W [n_out, n_in] row-major fp32 and b [n_out], DEVICE pointers. */
typedef struct {
  const float* W;
  const float* b;
  int32_t n_in;
  int32_t n_out;
} dpi_linear;

/* Frozen previous iterate u_k.  ELU between Linear layers (the only activation the configs use). */
typedef struct {
  int32_t kind; /* dpi_network_kind */
  /* MLP: `main` = Linear(1+nx,H1) .. Linear(HL,1).  PISGradNet: `main` = nn_module (nx+64 -> .. -> nx). */
  int32_t n_main;
  dpi_linear main[DPI_MAX_LINEAR];
  /* PISGradNet only (solution.py:170-207): */
  int32_t n_t_encoder;
  dpi_linear t_encoder[DPI_MAX_LINEAR];
  int32_t n_smooth;
  dpi_linear smooth_net[DPI_MAX_LINEAR];
  const float* timestep_phase; /* [64] device */
  const float* timestep_coeff; /* [64] device */
} dpi_network;

/* Deterministic mode: caller-supplied draws laid out exactly as the reference draws them
 * (row = point*M + sample).  DEVICE pointers; idx may be NULL unless sdgd_v > 0. */
typedef struct {
  const float* zT;    /* [n*M_T, nx]  terminal increments      data.py:914  */
  const float* u_s;   /* [n*M_I]      uniform integral times   data.py:359  */
  const float* zS;    /* [n*M_I, nx]  integral increments      data.py:360  */
  const int32_t* idx; /* [n*M_I, v]   SDGD coordinates         data.py:501  */
} dpi_draws;

typedef struct {
  int32_t M_T;            /* n_estimate_terminal */
  int32_t M_I;            /* n_estimate_integral */
  int32_t precision;      /* dpi_precision */
  float sample_bound;     /* clip(y, +-bound); <=0 or inf = off (data.py:123-128,222) */
  uint64_t seed;          /* Philox key (ignored when draws != NULL) */
  int64_t point_index_base; /* global index of tx row 0: Philox counters are keyed by the GLOBAL point
                               index so labels do not depend on the shard size or GPU count;
                               0 <= point_index_base, point_index_base + n <= 2^48 */
  float estimate_delta_t;   /* DATA.ESTIMATE_DELTA_T (data.py:529-575, :928-952): 0 = the estimators integrate to T;
                               > 0 = temporal-difference form: both integrate to t_next = min(t + delta, T) and the
                               terminal estimator takes u_k(t_next, X) in place of g(X) where t_next < T.  Cha and
                               OUProcessEquation only (as upstream asserts) */
  int32_t reserved0;
} dpi_label_config;

int dpi_version(void);
const char* dpi_last_error(void);

int dpi_create(int device, dpi_handle** out);
int dpi_destroy(dpi_handle* h);

int dpi_set_problem(dpi_handle* h, const dpi_problem* p);
/* Freeze u_k: every tensor of `net` is copied / repacked into handle-owned memory on `stream`, and the stream is
 * synchronised before the call returns.  Afterwards the label path and dpi_net_eval never read the caller's tensors
 * again: they may be trained in place (the next Picard iterate) or freed (picard/data.py:407-412 semantics). */
int dpi_set_network(dpi_handle* h, const dpi_network* net, void* stream);

/* tx [n, 1+nx] device, written.  t ~ U[eps, T-eps] (eps=0.01), x = x0 + sqrt(t*alpha) Z. */
int dpi_sample_tx(dpi_handle* h, int64_t n, int64_t point_index_base, uint64_t seed, float* tx, void* stream);

/* General form of the point sampler: n_factors = 0 is dpi_sample_tx with the margin eps (the reference uses 0.01 when
 * ESTIMATE_TERMINAL contains "ByGx", else 0: data.py:134-137); n_factors = k >= 1 is the reference's default sampler
 * `sample_t` (data.py:149-159): t = T (1 - U_1 U_2 ... U_k) with k = N - i + 1 uniforms, which concentrates the points
 * of late Picard iterations near t = 0. */
int dpi_sample_tx_ex(dpi_handle* h, int64_t n, int64_t point_index_base, uint64_t seed, int32_t n_factors, float eps,
                     float* tx, void* stream);

/* y [n, 1+nx] device, written: column 0 the value label, columns 1.. the gradient label. */
int dpi_label_gen(dpi_handle* h, const float* tx, int64_t n, const dpi_label_config* cfg,
                  const dpi_draws* draws /* nullable: Philox mode */, float* y, void* stream);

/* Synchronises `stream` and returns DPI_ERR_NONFINITE if any label written by dpi_label_gen on this handle since the
 * last check was inf / NaN (the status word is then cleared), DPI_OK otherwise. */
int dpi_check_labels(dpi_handle* h, void* stream);

/* The Philox draws a dpi_label_gen call with the same (cfg->seed, cfg->point_index_base, M_T, M_I) consumes, written
 * in the reference's layout (row = point * M + sample): zT [n*M_T, nx], u_s [n*M_I], zS [n*M_I, nx],
 * idx [n*M_I, sdgd_v] (int32).  DEVICE pointers, any of them NULL = skip.  Feeding them back through `draws`
 * reproduces the Philox-mode labels; feeding them to the reference reproduces them there (deterministic mode,
 * picard/data.py:914,359-360,501).  Not on the label path: the label kernels never materialise their increments. */
int dpi_export_draws(dpi_handle* h, int64_t n, const dpi_label_config* cfg, float* zT, float* u_s, float* zS,
                     int32_t* idx, void* stream);

/* Same, with HOST buffers (pinned for full speed): copies tx in, y out, synchronises, and checks the labels
 * (DPI_ERR_NONFINITE). If tx_in_host is NULL the points are drawn on the device (dpi_sample_tx) and returned in
 * tx_out_host.  Calls of 256 MB or more per direction are pipelined in ~64 MB chunks of points (H2D of chunk k + 1 and
 * D2H of chunk k - 1 on two copy streams of the handle under the kernels of chunk k); the labels are bitwise those of
 * the one-shot call, because Philox counters are keyed by the global point index. */
int dpi_label_gen_host(dpi_handle* h, const float* tx_in_host, int64_t n, const dpi_label_config* cfg,
                       float* tx_out_host /* nullable */, float* y_host, void* stream);

/* u_k and its input gradient on rows of tx: u [n], du [n, nx] (either may be NULL). fp32 path. */
int dpi_net_eval(dpi_handle* h, const float* tx, int64_t n, float* u, float* du, void* stream);

/* Diagonal of the Hessian of u_k with respect to x on rows of tx: d2u [n, nx] (forward 2-jets along the coordinate
 * directions, the construction the SDGD label kernel uses).  For the evaluation callback's Hessian metric
 * (picard/utils.py:329-478 with get_hessian :579-591, restricted to the diagonal).  Plain MLPs, hidden width <= 64. */
int dpi_net_hessian_diag(dpi_handle* h, const float* tx, int64_t n, float* d2u, void* stream);
/* The full Hessian of u_k in x at the rows of tx: d2u [n, nx, nx] (symmetric), what EvalCallback's TEST_HESSIAN compares
 * (picard/utils.py:198-200, get_hessian :579-591: nx backward passes upstream; here forward 2-jets along the
 * coordinate directions and the pairs e_i + e_j).  Plain MLPs of the fully nonlinear config's shape. */
int dpi_net_hessian(dpi_handle* h, const float* tx, int64_t n, float* d2u, void* stream);

/* ---- regression step (the trainer's inner loop) -------------------------------------------
 *   dpi_train_init  <- PicardSolutionGradientWrapper.__init__ + configure_optimizers   picard/solution_jac.py:125-146,
 *                                                                                       picard/solution.py:93-97
 *   dpi_train_grad  <- PicardSolutionGradientWrapper.training_step + loss.backward()   picard/solution_jac.py:172-216
 *   dpi_train_adam  <- torch.optim.Adam.step()                                          picard/solution.py:95-97
 * The network is a scalar-output MLP (NETWORK.TYPE = Value) whose Linear tensors are DEVICE fp32 pointers in torch
 * layout; they are the master weights and are updated in place.  dpi_train_grad writes the gradient of
 *   loss = mean_b w_b (u - y_0)^2 + fixed_weight * sum_j mean_b w_b (d_j u - y_j)^2 ,   w_b = exp(loss_beta * t_b)
 * flat in torch parameters() order (W_0, b_0, W_1, b_1, ...) and losses[3] = {total, value, gradient} (device).  Between
 * dpi_train_grad and dpi_train_adam a data-parallel caller all-reduces (averages) the flat gradient over its ranks. */
typedef struct {
  float lr, beta1, beta2, eps; /* torch.optim.Adam defaults: 1e-3, 0.9, 0.999, 1e-8 */
  float loss_beta;             /* TRAIN.LOSS.beta */
  float fixed_weight;          /* TRAIN.LOSS.SCALER.kwargs.fixed_weight (FixedLossScaler) */
} dpi_train_config;
int dpi_train_init(dpi_handle* h, const dpi_network* mlp, int64_t max_batch, void* stream);
int64_t dpi_train_param_count(const dpi_handle* h);
/* Fresh optimiser state (Adam moments zeroed): the reference builds a new Trainer/optimiser for every Picard iteration
 * (picard_iteration.py:279-289) while the network weights carry over (NETWORK.RELOAD). */
int dpi_train_reset(dpi_handle* h, void* stream);
int dpi_train_grad(dpi_handle* h, const float* tx, const float* y, int64_t batch, const dpi_train_config* cfg,
                   float* grad_flat, float* losses, void* stream);
int dpi_train_adam(dpi_handle* h, const float* grad_flat, const dpi_train_config* cfg, int64_t step /* 1-based */,
                   void* stream);
/* The trainer's inner loop over one iteration's label set on ONE device, without returning to the host between steps
 * (picard_iteration.py: Trainer.fit over dataset_with_gradients): n_epochs passes over the n rows of (tx, y) in
 * unshuffled mini-batches of `batch` (config.py:90), Adam step counter continuing from first_step (1-based).
 * losses_last [3] (device, nullable) receives the loss terms of the last step.  Asynchronous on `stream`. */
int dpi_train_fit(dpi_handle* h, const float* tx, const float* y, int64_t n, int64_t batch, int64_t n_epochs,
                  const dpi_train_config* cfg, int64_t first_step, float* losses_last, void* stream);

/* ---- data-parallel regression over NVLink (one process per GPU of one node) ------------------------------------------
 * The reference's trainer would all-reduce the MLP gradients between backward and optimiser step (DDP under
 * pl.Trainer).  Here the exchange is part of the optimiser kernel: every rank's gradient block is mapped
 * into its peers (CUDA IPC), and ONE kernel per step waits until all ranks' gradients are written, sums them over peer
 * memory in rank order (bitwise identical replicas, no broadcast), applies Adam, and holds until all peers have read.
 * It replaces the Adam kernel inside dpi_train_fit's captured step, so the inner loop stays native and graph-replayed.
 *   dpi_train_dp_export : handle_out receives DPI_DP_HANDLE_BYTES bytes naming this rank's block (after dpi_train_init);
 *   dpi_train_dp_connect: `handles` = world x DPI_DP_HANDLE_BYTES bytes, rank order (all-gathered by the caller, e.g.
 *                         over torch.distributed); every rank then calls dpi_train_fit with its share of each global
 *                         mini-batch (batch / world rows) and the same n, n_epochs, first_step. */
#define DPI_DP_HANDLE_BYTES 64
int dpi_train_dp_export(dpi_handle* h, void* handle_out);
int dpi_train_dp_connect(dpi_handle* h, int32_t rank, int32_t world, const void* handles);

/* Kernels launched by this handle since creation (for bench.py's gpu_launches). */
int64_t dpi_launch_count(const dpi_handle* h);

/* Per-stage device timing (replaces the reference's PROFILE_CUDA / count_cuda_time_wrapper switch,
 * picard/utils.py:50-66).  When enabled, dpi_label_gen brackets each stage with CUDA events on the launching
 * stream; dpi_stage_ms waits for the stage's end event of the most recent dpi_label_gen call and returns its
 * duration.  Stages: the terminal estimator, the per-point baseline pass, the integral (u_k over paths) kernel. */
typedef enum { DPI_STAGE_TERMINAL = 0, DPI_STAGE_BASELINE = 1, DPI_STAGE_INTEGRAL = 2, DPI_STAGE_COUNT = 3 } dpi_stage;
int dpi_set_stage_timing(dpi_handle* h, int enable);
int dpi_stage_ms(dpi_handle* h, int stage, float* ms);

/* Hardware self-test of the tcgen05 building blocks (shared-memory descriptors, SWIZZLE_128B operand layout,
 * instruction descriptor, TMEM readback): D[128x128] = fp16(A[128x128]) . fp16(B[128x128])^T with fp32
 * accumulation; A, B, D are fp32 row-major DEVICE buffers.  Synchronises the stream. */
int dpi_selftest_umma(dpi_handle* h, const float* A, const float* B, float* D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DPI_B200_H */
