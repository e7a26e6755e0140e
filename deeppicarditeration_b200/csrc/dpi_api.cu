// C ABI (include/dpi_b200.h): handle, problem/network packing and the label-generation driver.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dpi_common.cuh"
#include "dpi_kernels.h"

namespace dpi {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace dpi

using namespace dpi;

struct dpi_handle {
  int device = 0;
  bool has_problem = false;
  bool has_net = false;
  ProblemDev P{};
  int net_kind = DPI_NET_ZERO;
  MlpDev mlp{};
  TcNetDev tc{};
  bool tc_ready = false;
  PisDev pis{};
  HjbTcNet* hjb_tc = nullptr;   // fused tensor-core packing of the PISGradNet (null: shape not eligible)
  float* hjb_part = nullptr;    // per-tile partial sums of the fused path
  int64_t hjb_part_floats = 0;
  // regression step: trainable MLP (caller's tensors), Adam moments, activation scratch
  dpi_linear train_lin[DPI_MAX_LINEAR] = {};
  int n_train_lin = 0;
  int64_t train_params = 0, train_max_batch = 0;
  float *train_m = nullptr, *train_v = nullptr, *train_scratch = nullptr, *train_grad = nullptr;
  // dpi_train_fit replays one optimiser step as a CUDA graph on an internal stream
  cudaStream_t train_stream = nullptr;
  TrainAux train_aux;  // side stream + fork/join events of the weight-gradient branch
  cudaEvent_t train_ev_in = nullptr, train_ev_out = nullptr;
  float *train_stage_tx = nullptr, *train_stage_y = nullptr;
  long long* train_step_dev = nullptr;
  cudaGraphExec_t train_graph = nullptr;
  int64_t train_graph_batch = 0, train_graph_nodes = 0;
  dpi_train_config train_graph_cfg{};
  DpLinks dp;  // data-parallel peers (dpi_train_dp_connect); world == 1: single device
  dpi_linear mlp_raw[DPI_MAX_LINEAR] = {};  // the caller's MLP tensors (torch layout), for dpi_net_eval
  int n_mlp_raw = 0;
  // activations of one chunk of rows of the batched-GEMM (PISGradNet) path
  float* hjb_scratch = nullptr;
  int64_t hjb_scratch_floats = 0;
  void* td_scratch = nullptr;  // TD estimator: path-end rows and their u_k values of one chunk of points
  int64_t td_scratch_bytes = 0;
  std::vector<void*> prob_bufs;  // device allocations owned by the problem
  std::vector<void*> net_bufs;   // device allocations owned by the packed network
  // per-point scratch: control variate f_b [n], first-layer constant c0 [n][HP]
  float* fb = nullptr;
  float* c0 = nullptr;
  int64_t scratch_points = 0;
  int scratch_hp = 0;
  // unit partial sums of the tensor-core path
  float* partG = nullptr;
  float* partF = nullptr;
  void* tc_zring = nullptr;  // increment ring of the chunked tensor-core kernel (tc_zring_bytes)
  size_t tc_zring_size = 0;
  int64_t part_units = 0;
  // staging for the *_host entry points
  float* stage_tx = nullptr;
  float* stage_y = nullptr;
  int64_t stage_floats = 0;
  // copy streams + events of the pipelined host path (dpi_label_gen_host on large calls), created on first use
  cudaStream_t host_in = nullptr, host_out = nullptr;
  std::vector<cudaEvent_t> host_ev;
  // side stream of dpi_label_gen: the per-point baseline pass (control variate f(t,x), first-layer constants) has no
  // dependence on the terminal estimator, so it runs beside it (fork / join events; capturable)
  cudaStream_t aux = nullptr;
  cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
  int64_t launches = 0;
  int* nonfinite_flag = nullptr;  // device: set by the label kernels' final check when a label is inf / NaN
  // optional per-stage timing (dpi_set_stage_timing)
  bool timing = false;
  cudaEvent_t ev[DPI_STAGE_COUNT][2] = {};
  bool ev_valid[DPI_STAGE_COUNT] = {};
};

namespace {

int free_all(std::vector<void*>& v) {
  for (void* p : v) cudaFree(p);
  v.clear();
  return 0;
}

template <typename T>
int dev_alloc(std::vector<void*>& owner, size_t count, T** out) {
  void* p = nullptr;
  DPI_CUDA_CHECK(cudaMalloc(&p, sizeof(T) * (count ? count : 1)));
  owner.push_back(p);
  *out = reinterpret_cast<T*>(p);
  return DPI_OK;
}

int upload_f32(std::vector<void*>& owner, const std::vector<float>& host, float** out) {
  int rc = dev_alloc(owner, host.size(), out);
  if (rc) return rc;
  DPI_CUDA_CHECK(cudaMemcpy(*out, host.data(), sizeof(float) * host.size(), cudaMemcpyHostToDevice));
  return DPI_OK;
}

int ensure_scratch(dpi_handle* h, int64_t n, int hp) {
  if (n <= h->scratch_points && hp <= h->scratch_hp) return DPI_OK;
  if (h->fb) cudaFree(h->fb);
  if (h->c0) cudaFree(h->c0);
  h->fb = h->c0 = nullptr;
  h->scratch_points = 0;
  const int64_t cap = n > h->scratch_points ? n : h->scratch_points;
  const int chp = hp > h->scratch_hp ? hp : h->scratch_hp;
  DPI_CUDA_CHECK(cudaMalloc((void**)&h->fb, sizeof(float) * cap));
  DPI_CUDA_CHECK(cudaMalloc((void**)&h->c0, sizeof(float) * cap * chp));
  h->scratch_points = cap;
  h->scratch_hp = chp;
  return DPI_OK;
}

int ensure_parts(dpi_handle* h, int64_t units) {
  if (units <= h->part_units) return DPI_OK;
  if (h->partG) cudaFree(h->partG);
  if (h->partF) cudaFree(h->partF);
  h->partG = h->partF = nullptr;
  h->part_units = 0;
  DPI_CUDA_CHECK(cudaMalloc((void**)&h->partG, sizeof(float) * units * kTcPartGFloatsPerUnit));
  // + the unit hand-out counter of the tensor-core kernel (one int behind the partials)
  DPI_CUDA_CHECK(cudaMalloc((void**)&h->partF, sizeof(float) * (units * kTcPartFFloatsPerUnit + 4)));
  h->part_units = units;
  return DPI_OK;
}

int ensure_stage(dpi_handle* h, int64_t floats) {
  if (floats <= h->stage_floats) return DPI_OK;
  if (h->stage_tx) cudaFree(h->stage_tx);
  if (h->stage_y) cudaFree(h->stage_y);
  h->stage_tx = h->stage_y = nullptr;
  h->stage_floats = 0;
  DPI_CUDA_CHECK(cudaMalloc((void**)&h->stage_tx, sizeof(float) * floats));
  DPI_CUDA_CHECK(cudaMalloc((void**)&h->stage_y, sizeof(float) * floats));
  h->stage_floats = floats;
  return DPI_OK;
}

int ensure_hjb_scratch(dpi_handle* h, int64_t floats) {
  if (floats <= h->hjb_scratch_floats) return DPI_OK;
  if (h->hjb_scratch) cudaFree(h->hjb_scratch);
  h->hjb_scratch = nullptr;
  h->hjb_scratch_floats = 0;
  DPI_CUDA_CHECK(cudaMalloc((void**)&h->hjb_scratch, sizeof(float) * floats));
  h->hjb_scratch_floats = floats;
  return DPI_OK;
}
// activation scratch of the batched-GEMM path: at most this many bytes per chunk of rows
constexpr int64_t kHjbScratchBudget = 6LL << 30;

int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Private copy of one Linear layer (picard/data.py:407-412 freezes `solution`: the label path must not see the
// caller's tensors again -- they may be trained in place or freed once dpi_set_network returns).
int freeze_linear(std::vector<void*>& owner, const dpi_linear& src, dpi_linear* dst, cudaStream_t st) {
  float *W = nullptr, *b = nullptr;
  const size_t nw = (size_t)src.n_out * src.n_in;
  int rc;
  if ((rc = dev_alloc(owner, nw, &W))) return rc;
  if ((rc = dev_alloc(owner, (size_t)src.n_out, &b))) return rc;
  DPI_CUDA_CHECK(cudaMemcpyAsync(W, src.W, sizeof(float) * nw, cudaMemcpyDeviceToDevice, st));
  DPI_CUDA_CHECK(cudaMemcpyAsync(b, src.b, sizeof(float) * src.n_out, cudaMemcpyDeviceToDevice, st));
  *dst = src;
  dst->W = W;
  dst->b = b;
  return DPI_OK;
}

// RAII bracket: records the stage's start event now and its end event when it goes out of scope.
struct StageTimer {
  dpi_handle* h;
  int stage;
  cudaStream_t st;
  StageTimer(dpi_handle* h_, int stage_, cudaStream_t st_) : h(h_), stage(stage_), st(st_) {
    if (h->timing) cudaEventRecord(h->ev[stage][0], st);
  }
  ~StageTimer() {
    if (h->timing) {
      cudaEventRecord(h->ev[stage][1], st);
      h->ev_valid[stage] = true;
    }
  }
};

}  // namespace

extern "C" {

int dpi_version(void) { return DPI_B200_VERSION; }
const char* dpi_last_error(void) { return g_err; }

int dpi_create(int device, dpi_handle** out) {
  if (!out) {
    set_error("dpi_create: out is NULL");
    return DPI_ERR_INVALID;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("dpi_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    return DPI_ERR_CUDA;
  }
  if (device < 0 || device >= count) {
    set_error("dpi_create: device %d out of range (%d devices)", device, count);
    return DPI_ERR_INVALID;
  }
  cudaDeviceProp prop;
  DPI_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("dpi_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
              prop.minor);
    return DPI_ERR_CUDA;
  }
  DPI_CUDA_CHECK(cudaSetDevice(device));
  dpi_handle* h = new dpi_handle();
  h->device = device;
  if (cudaMalloc((void**)&h->nonfinite_flag, sizeof(int)) != cudaSuccess ||
      cudaMemset(h->nonfinite_flag, 0, sizeof(int)) != cudaSuccess) {
    set_error("dpi_create: cannot allocate the status word: %s", cudaGetErrorString(cudaGetLastError()));
    delete h;
    return DPI_ERR_CUDA;
  }
  *out = h;
  return DPI_OK;
}

int dpi_destroy(dpi_handle* h) {
  if (!h) return DPI_OK;
  cudaSetDevice(h->device);
  free_all(h->prob_bufs);
  free_all(h->net_bufs);
  if (h->fb) cudaFree(h->fb);
  if (h->c0) cudaFree(h->c0);
  if (h->stage_tx) cudaFree(h->stage_tx);
  if (h->stage_y) cudaFree(h->stage_y);
  if (h->aux) cudaStreamDestroy(h->aux);
  if (h->aux_fork) cudaEventDestroy(h->aux_fork);
  if (h->aux_join) cudaEventDestroy(h->aux_join);
  if (h->host_in) cudaStreamDestroy(h->host_in);
  if (h->host_out) cudaStreamDestroy(h->host_out);
  for (cudaEvent_t e : h->host_ev) cudaEventDestroy(e);
  if (h->partG) cudaFree(h->partG);
  if (h->tc_zring) cudaFree(h->tc_zring);
  if (h->partF) cudaFree(h->partF);
  if (h->hjb_scratch) cudaFree(h->hjb_scratch);
  if (h->hjb_part) cudaFree(h->hjb_part);
  if (h->td_scratch) cudaFree(h->td_scratch);
  if (h->hjb_tc) hjb_tc_free(h->hjb_tc);
  if (h->nonfinite_flag) cudaFree(h->nonfinite_flag);
  if (h->train_m) cudaFree(h->train_m);
  if (h->train_v) cudaFree(h->train_v);
  if (h->train_scratch) cudaFree(h->train_scratch);
  for (int r = 0; r < DPI_DP_MAX_WORLD; ++r)
    if (h->dp.mapped[r]) cudaIpcCloseMemHandle(h->dp.mapped[r]);
  if (h->train_grad) cudaFree(h->train_grad);
  if (h->train_graph) cudaGraphExecDestroy(h->train_graph);
  if (h->train_stage_tx) cudaFree(h->train_stage_tx);
  if (h->train_stage_y) cudaFree(h->train_stage_y);
  if (h->train_step_dev) cudaFree(h->train_step_dev);
  if (h->train_ev_in) cudaEventDestroy(h->train_ev_in);
  if (h->train_ev_out) cudaEventDestroy(h->train_ev_out);
  if (h->train_stream) cudaStreamDestroy(h->train_stream);
  if (h->train_aux.side) cudaStreamDestroy(h->train_aux.side);
  for (int l = 0; l < DPI_MAX_LINEAR; ++l)
    if (h->train_aux.fork[l]) cudaEventDestroy(h->train_aux.fork[l]);
  if (h->train_aux.join) cudaEventDestroy(h->train_aux.join);
  for (int i = 0; i < 3; ++i)
    if (h->train_aux.pre[i]) cudaEventDestroy(h->train_aux.pre[i]);
  for (int s = 0; s < DPI_STAGE_COUNT; ++s)
    for (int e = 0; e < 2; ++e)
      if (h->ev[s][e]) cudaEventDestroy(h->ev[s][e]);
  delete h;
  return DPI_OK;
}

int dpi_set_problem(dpi_handle* h, const dpi_problem* p) {
  if (!h || !p) {
    set_error("dpi_set_problem: NULL argument");
    return DPI_ERR_INVALID;
  }
  if (p->nx < 1 || !(p->alpha > 0.f) || !(p->T > 0.f)) {
    set_error("dpi_set_problem: need nx>=1, alpha>0, T>0 (got nx=%d alpha=%g T=%g)", p->nx, p->alpha, p->T);
    return DPI_ERR_INVALID;
  }
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  DPI_CUDA_CHECK(cudaDeviceSynchronize());  // nothing in flight may still read the old problem buffers
  free_all(h->prob_bufs);
  h->has_problem = false;
  ProblemDev P{};
  P.eq = p->kind;
  P.nx = p->nx;
  P.T = p->T;
  P.alpha = p->alpha;
  P.sigma = std::sqrt(p->alpha);
  P.sdgd_v = p->sdgd_v;
  const int d = p->nx;
  if (p->kind == DPI_EQ_CHA) {
    // equations.py:283-289: k = k_cfg/sqrt(nx); c = (2 + k^2 alpha d) / (2 k alpha d)
    const double k = (double)p->cha_k / std::sqrt((double)d);
    const double kad = k * (double)p->alpha * d;
    P.cha_k = (float)k;
    P.cha_c = (float)((2.0 + k * kad) / (2.0 * kad));
  } else if (p->kind == DPI_EQ_GBM) {
    if (p->gbm_nodes < 1 || p->gbm_nodes > kMaxGbmNodes || !p->gbm_w || !p->gbm_v) {
      set_error("dpi_set_problem: GBM needs 1..%d nodes with w and v", kMaxGbmNodes);
      return DPI_ERR_INVALID;
    }
    P.gbm_nodes = p->gbm_nodes;
    std::vector<float> w((size_t)p->gbm_nodes * (1 + d));
    for (size_t i = 0; i < w.size(); ++i) w[i] = (float)p->gbm_w[i];
    for (int k = 0; k < p->gbm_nodes; ++k) P.gbm_v[k] = (float)p->gbm_v[k];
    float* dw = nullptr;
    int rc = upload_f32(h->prob_bufs, w, &dw);
    if (rc) return rc;
    P.gbm_w = dw;
  } else if (p->kind == DPI_EQ_OU) {
    if (p->ou_K < 1 || p->ou_K > kMaxOuK || !p->ou_mean || !p->ou_diagvar || !p->ou_pi) {
      set_error("dpi_set_problem: OU needs 1..%d mixture components with mean, diagvar, pi", kMaxOuK);
      return DPI_ERR_INVALID;
    }
    P.ou_K = p->ou_K;
    P.ou_theta = p->ou_theta;
    P.ou_mu = p->ou_mu;
    P.ou_x0_std = std::sqrt(p->ou_alpha_init);
    std::vector<float> mean((size_t)p->ou_K * d), ivar((size_t)p->ou_K * d);
    for (int k = 0; k < p->ou_K; ++k) {
      double logdet = 0.0;
      for (int j = 0; j < d; ++j) {
        mean[(size_t)k * d + j] = (float)p->ou_mean[(size_t)k * d + j];
        ivar[(size_t)k * d + j] = (float)(1.0 / p->ou_diagvar[(size_t)k * d + j]);
        logdet += std::log(p->ou_diagvar[(size_t)k * d + j]);
      }
      // utils.py:866-874: log pi_k - 0.5 (d log 2pi + log det Sigma_k)
      P.ou_lognorm[k] = (float)(std::log(p->ou_pi[k]) - 0.5 * (d * std::log(2.0 * M_PI) + logdet));
    }
    float *dm = nullptr, *di = nullptr;
    int rc = upload_f32(h->prob_bufs, mean, &dm);
    if (rc) return rc;
    rc = upload_f32(h->prob_bufs, ivar, &di);
    if (rc) return rc;
    P.ou_mean = dm;
    P.ou_ivar = di;
  } else {
    set_error("dpi_set_problem: unknown equation kind %d", p->kind);
    return DPI_ERR_INVALID;
  }
  h->P = P;
  h->has_problem = true;
  return DPI_OK;
}

int dpi_set_network(dpi_handle* h, const dpi_network* net, void* stream) {
  if (!h || !net) {
    set_error("dpi_set_network: NULL argument");
    return DPI_ERR_INVALID;
  }
  if (!h->has_problem) {
    set_error("dpi_set_network: call dpi_set_problem first");
    return DPI_ERR_STATE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  DPI_CUDA_CHECK(cudaDeviceSynchronize());
  free_all(h->net_bufs);
  if (h->hjb_tc) hjb_tc_free(h->hjb_tc);
  h->hjb_tc = nullptr;
  h->has_net = false;
  h->tc_ready = false;
  h->net_kind = net->kind;
  const int nx = h->P.nx;
  if (net->kind == DPI_NET_ZERO) {
    h->has_net = true;
    return DPI_OK;
  }
  if (net->kind == DPI_NET_MLP) {
    const int L = net->n_main;
    if (L < 2 || L > DPI_MAX_LINEAR) {
      set_error("dpi_set_network: MLP needs 2..%d Linear layers (got %d)", DPI_MAX_LINEAR, L);
      return DPI_ERR_INVALID;
    }
    int maxw = 0;
    for (int l = 0; l < L; ++l) {
      const dpi_linear& ll = net->main[l];
      const int want_in = (l == 0) ? 1 + nx : net->main[l - 1].n_out;
      if (!ll.W || !ll.b || ll.n_in != want_in || ll.n_out < 1) {
        set_error("dpi_set_network: layer %d has shape [%d,%d], expected n_in=%d", l, ll.n_out, ll.n_in, want_in);
        return DPI_ERR_INVALID;
      }
      if (l < L - 1 && ll.n_out > maxw) maxw = ll.n_out;
    }
    if (net->main[L - 1].n_out != 1) {
      // TYPE "ValueGradient"/"OnlyGradient" nets (solution.py:300-307) are not used by any BASELINE config
      set_error("dpi_set_network: output dim %d != 1 (only NETWORK.TYPE=Value is built)", net->main[L - 1].n_out);
      return DPI_ERR_UNSUPPORTED;
    }
    if (maxw > 128) {
      set_error("dpi_set_network: hidden width %d > 128 is not built for the plain-MLP kernels", maxw);
      return DPI_ERR_UNSUPPORTED;
    }
    MlpDev M{};
    M.n_linear = L;
    M.nx = nx;
    // five-layer Cha networks keep the 128-wide layout whatever their widths: the tensor-core path (tc_pack_network)
    // pads narrower layers with zeros and reads the baseline pass's per-point rows with that stride
    M.HP = (maxw <= 64 && !(L == 5 && h->P.eq == DPI_EQ_CHA)) ? 64 : 128;
    M.d_pad = round_up(nx, 128);
    float *Wt1, *wt, *b1, *wsum, *wout, *bout;
    float* Wth[DPI_MAX_LINEAR] = {nullptr};
    float* Wth_bf[DPI_MAX_LINEAR] = {nullptr};
    float* bh[DPI_MAX_LINEAR] = {nullptr};
    int rc;
    if ((rc = dev_alloc(h->net_bufs, (size_t)M.d_pad * M.HP, &Wt1))) return rc;
    if ((rc = dev_alloc(h->net_bufs, M.HP, &wt))) return rc;
    if ((rc = dev_alloc(h->net_bufs, M.HP, &b1))) return rc;
    if ((rc = dev_alloc(h->net_bufs, M.HP, &wsum))) return rc;
    if ((rc = dev_alloc(h->net_bufs, M.HP, &wout))) return rc;
    if ((rc = dev_alloc(h->net_bufs, 1, &bout))) return rc;
    for (int l = 0; l < L - 2; ++l) {
      if ((rc = dev_alloc(h->net_bufs, (size_t)M.HP * M.HP, &Wth[l]))) return rc;
      if ((rc = dev_alloc(h->net_bufs, (size_t)M.HP * M.HP, &Wth_bf[l]))) return rc;
      if ((rc = dev_alloc(h->net_bufs, M.HP, &bh[l]))) return rc;
    }
    if ((rc = launch_pack_mlp(*net, nx, M.HP, M.d_pad, Wt1, wt, b1, wsum, Wth, Wth_bf, bh, wout, bout, st))) return rc;
    h->launches += 6 + 3 * (L - 2);
    M.Wt1 = Wt1;
    M.wt = wt;
    M.b1 = b1;
    M.wsum = wsum;
    M.wout = wout;
    M.bout = bout;
    for (int l = 0; l < L - 2; ++l) {
      M.Wth[l] = Wth[l];
      M.Wth_bf[l] = Wth_bf[l];
      M.bh[l] = bh[l];
    }
    h->mlp = M;
    // dpi_net_eval (value + full input gradient) reads torch-layout weights: keep a private copy, not the caller's
    for (int l = 0; l < L; ++l)
      if ((rc = freeze_linear(h->net_bufs, net->main[l], &h->mlp_raw[l], st))) return rc;
    h->n_mlp_raw = L;
    // bf16 tensor-core packing (tcgen05 path) when the shape qualifies
    rc = tc_pack_network(*net, h->P, M, h->net_bufs, &h->tc, &h->tc_ready, &h->launches, st);
    if (rc) return rc;
    DPI_CUDA_CHECK(cudaStreamSynchronize(st));  // the caller's tensors are not read after this call returns
    h->has_net = true;
    return DPI_OK;
  }
  if (net->kind == DPI_NET_PISGRAD) {
    // solution.py:138-207: channels = 64; t_encoder 128->64->64; smooth_net 128->64->..->64->dim; nn_module (64+nx)->..->nx
    if (h->P.eq != DPI_EQ_OU) {
      set_error("dpi_set_network: PISGradNet is built for OUProcessEquation only (its residual uses the GMM terminal)");
      return DPI_ERR_UNSUPPORTED;
    }
    auto chain_ok = [&](const dpi_linear* c, int n, int n_in, int n_out, int mid /* 0 = free */) {
      if (n < 2 || n > DPI_MAX_LINEAR) return false;
      for (int l = 0; l < n; ++l) {
        const int want_in = l == 0 ? n_in : c[l - 1].n_out;
        if (!c[l].W || !c[l].b || c[l].n_in != want_in || c[l].n_out < 1) return false;
        if (mid && l < n - 1 && c[l].n_out != mid) return false;
      }
      return c[n - 1].n_out == n_out;
    };
    if (!chain_ok(net->t_encoder, net->n_t_encoder, 128, 64, 64) || net->n_t_encoder != 2 ||
        !chain_ok(net->smooth_net, net->n_smooth, 128, net->smooth_net[net->n_smooth > 0 ? net->n_smooth - 1 : 0].n_out, 64) ||
        !chain_ok(net->main, net->n_main, 64 + nx, nx, 0) || !net->timestep_phase || !net->timestep_coeff) {
      set_error("dpi_set_network: PISGradNet chains do not have the shapes of picard/solution.py:170-207 for nx=%d", nx);
      return DPI_ERR_INVALID;
    }
    PisDev N{};
    N.n_main = net->n_main, N.n_t_encoder = net->n_t_encoder, N.n_smooth = net->n_smooth;
    N.hmax = 0;
    int rc_f;
    for (int l = 0; l < net->n_main; ++l) {
      if ((rc_f = freeze_linear(h->net_bufs, net->main[l], &N.main[l], st))) return rc_f;
      if (l < net->n_main - 1 && net->main[l].n_out > N.hmax) N.hmax = net->main[l].n_out;
    }
    for (int l = 0; l < net->n_t_encoder; ++l)
      if ((rc_f = freeze_linear(h->net_bufs, net->t_encoder[l], &N.t_encoder[l], st))) return rc_f;
    for (int l = 0; l < net->n_smooth; ++l)
      if ((rc_f = freeze_linear(h->net_bufs, net->smooth_net[l], &N.smooth_net[l], st))) return rc_f;
    {
      float* pc = nullptr;  // [phase 64 | coeff 64]
      if ((rc_f = dev_alloc(h->net_bufs, 128, &pc))) return rc_f;
      DPI_CUDA_CHECK(cudaMemcpyAsync(pc, net->timestep_phase, 64 * sizeof(float), cudaMemcpyDeviceToDevice, st));
      DPI_CUDA_CHECK(cudaMemcpyAsync(pc + 64, net->timestep_coeff, 64 * sizeof(float), cudaMemcpyDeviceToDevice, st));
      N.phase = pc, N.coeff = pc + 64;
    }
    // K-major ([n_in][n_out]) copies of the nn_module weights: the VJP GEMMs of the tensor-core path then load both
    // operands with 16-byte copies instead of transposing a [K][N] weight element by element
    for (int l = 0; l < net->n_main; ++l) {
      float* wt = nullptr;
      DPI_CUDA_CHECK(cudaMalloc((void**)&wt, sizeof(float) * (size_t)net->main[l].n_out * net->main[l].n_in));
      h->net_bufs.push_back(wt);
      if (int rc_t = gemm_transpose(N.main[l].W, net->main[l].n_out, net->main[l].n_in, wt, st)) return rc_t;
      N.main_t[l] = wt;
    }
    h->launches += net->n_main;
    if (int rc_p = hjb_tc_pack(N, nx, h->net_bufs, &h->hjb_tc, &h->launches, st)) return rc_p;
    h->pis = N;
    DPI_CUDA_CHECK(cudaStreamSynchronize(st));  // the caller's tensors are not read after this call returns
    h->has_net = true;
    return DPI_OK;
  }
  set_error("dpi_set_network: unknown network kind %d", net->kind);
  return DPI_ERR_INVALID;
}

int dpi_sample_tx(dpi_handle* h, int64_t n, int64_t point_index_base, uint64_t seed, float* tx, void* stream) {
  return dpi_sample_tx_ex(h, n, point_index_base, seed, 0, 0.01f, tx, stream);
}

int dpi_sample_tx_ex(dpi_handle* h, int64_t n, int64_t point_index_base, uint64_t seed, int32_t n_factors, float eps,
                     float* tx, void* stream) {
  if (!h || !h->has_problem) {
    set_error("dpi_sample_tx: problem not set");
    return DPI_ERR_STATE;
  }
  if (n < 0 || (n > 0 && !tx)) {
    set_error("dpi_sample_tx: bad n/tx");
    return DPI_ERR_INVALID;
  }
  if (point_index_base < 0 || point_index_base > (1LL << 48) - n) {
    set_error("dpi_sample_tx: point_index_base %lld + n %lld leaves the 48-bit range of the Philox counter",
              (long long)point_index_base, (long long)n);
    return DPI_ERR_INVALID;
  }
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  if (n_factors < 0 || n_factors > 4096 || !(eps >= 0.f) || !(2.f * eps < h->P.T)) {
    set_error("dpi_sample_tx: bad sampler parameters (n_factors=%d eps=%g)", n_factors, eps);
    return DPI_ERR_INVALID;
  }
  int rc = launch_sample_tx(h->P, n, point_index_base, seed, eps, n_factors, tx, reinterpret_cast<cudaStream_t>(stream));
  if (!rc && n > 0) h->launches += 1;
  return rc;
}

static int label_gen_body(dpi_handle* h, const float* tx, int64_t n, const dpi_label_config* cfg, const dpi_draws* draws,
                          float* y, void* stream);

// Philox counters hold 48 bits of the global point index (dpi_common.cuh: philox_counter)
static bool point_range_ok(int64_t base, int64_t n) { return base >= 0 && n >= 0 && base <= (1LL << 48) - n; }

int dpi_label_gen(dpi_handle* h, const float* tx, int64_t n, const dpi_label_config* cfg, const dpi_draws* draws,
                  float* y, void* stream) {
  if (!h || !cfg) {
    set_error("dpi_label_gen: NULL argument");
    return DPI_ERR_INVALID;
  }
  if (cfg->precision != DPI_PREC_FP32 && cfg->precision != DPI_PREC_F16 &&
      !(cfg->precision == DPI_PREC_TF32 && h->net_kind == DPI_NET_PISGRAD)) {
    set_error("dpi_label_gen: precision %d is not built (DPI_PREC_BF16 operands are not implemented: use DPI_PREC_F16, "
              "same operand width with three more mantissa bits, or DPI_PREC_FP32)", cfg->precision);
    return DPI_ERR_UNSUPPORTED;
  }
  if (!draws && !point_range_ok(cfg->point_index_base, n)) {
    set_error("dpi_label_gen: point_index_base %lld + n %lld leaves the 48-bit range of the Philox counter",
              (long long)cfg->point_index_base, (long long)n);
    return DPI_ERR_INVALID;
  }
  int rc = label_gen_body(h, tx, n, cfg, draws, y, stream);
  if (rc || n <= 0) return rc;
  // the tensor-core finish kernel checks its own output; every other path gets one small sweep over y
  const bool tc_cha = h->net_kind == DPI_NET_MLP && h->P.eq == DPI_EQ_CHA && cfg->precision == DPI_PREC_F16;
  const bool tc_hjb = h->net_kind == DPI_NET_PISGRAD && cfg->precision == DPI_PREC_F16 && h->hjb_tc;
  if (!tc_cha && !tc_hjb) {
    rc = launch_check_finite(y, n * (int64_t)(1 + h->P.nx), h->nonfinite_flag, reinterpret_cast<cudaStream_t>(stream));
    if (!rc) h->launches += 1;
  }
  return rc;
}

int dpi_check_labels(dpi_handle* h, void* stream) {
  if (!h) {
    set_error("dpi_check_labels: NULL handle");
    return DPI_ERR_INVALID;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  int flag = 0;
  DPI_CUDA_CHECK(cudaMemcpyAsync(&flag, h->nonfinite_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  DPI_CUDA_CHECK(cudaStreamSynchronize(st));
  if (flag) {
    DPI_CUDA_CHECK(cudaMemsetAsync(h->nonfinite_flag, 0, sizeof(int), st));
    set_error("non-finite labels: the frozen network leaves the range of the selected arithmetic (fp16 operands hold "
              "|log2(e) * pre-activation| < 6.5e4) or the inputs were not finite; re-run with DPI_PREC_FP32");
    return DPI_ERR_NONFINITE;
  }
  return DPI_OK;
}

int dpi_export_draws(dpi_handle* h, int64_t n, const dpi_label_config* cfg, float* zT, float* u_s, float* zS,
                     int32_t* idx, void* stream) {
  if (!h || !cfg || !h->has_problem) {
    set_error("dpi_export_draws: problem not set / NULL argument");
    return DPI_ERR_STATE;
  }
  if (n < 0 || cfg->M_T < 1 || cfg->M_I < 1 || !point_range_ok(cfg->point_index_base, n)) {
    set_error("dpi_export_draws: bad arguments (n=%lld M_T=%d M_I=%d base=%lld)", (long long)n, cfg->M_T, cfg->M_I,
              (long long)cfg->point_index_base);
    return DPI_ERR_INVALID;
  }
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  return launch_export_draws(h->P, n, cfg->M_T, cfg->M_I, cfg->seed, cfg->point_index_base, zT, u_s, zS, idx,
                             reinterpret_cast<cudaStream_t>(stream), &h->launches);
}

static int label_gen_body(dpi_handle* h, const float* tx, int64_t n, const dpi_label_config* cfg, const dpi_draws* draws,
                          float* y, void* stream) {
  if (!h || !cfg) {
    set_error("dpi_label_gen: NULL argument");
    return DPI_ERR_INVALID;
  }
  if (!h->has_problem || !h->has_net) {
    set_error("dpi_label_gen: problem/network not set");
    return DPI_ERR_STATE;
  }
  if (n < 0 || (n > 0 && (!tx || !y)) || cfg->M_T < 1 || cfg->M_I < 1) {
    set_error("dpi_label_gen: bad arguments (n=%lld M_T=%d M_I=%d)", (long long)n, cfg->M_T, cfg->M_I);
    return DPI_ERR_INVALID;
  }
  if (n == 0) return DPI_OK;
  if (n > 0x7fffffffLL) {
    set_error("dpi_label_gen: n=%lld exceeds the grid limit; shard the call", (long long)n);
    return DPI_ERR_INVALID;
  }
  if (draws && (!draws->zT || !draws->u_s || !draws->zS)) {
    set_error("dpi_label_gen: deterministic mode needs zT, u_s and zS");
    return DPI_ERR_INVALID;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  if (!(cfg->estimate_delta_t >= 0.f)) {
    set_error("dpi_label_gen: estimate_delta_t must be >= 0");
    return DPI_ERR_INVALID;
  }
  h->P.td_delta = cfg->estimate_delta_t;  // the horizon of both estimators (dpi_common.cuh: horizon)
  const ProblemDev& P = h->P;
  const float bound = cfg->sample_bound;
  int rc;

  for (int s = 0; s < DPI_STAGE_COUNT; ++s) h->ev_valid[s] = false;

  // 0) Cha + MLP: the baseline pass goes to the side stream and runs under the terminal estimator.  It is a chain of
  //    small per-point GEMMs (62 us whether 512 or 4096 points are labelled): 1 % of a 4096-point call, 6 % of the
  //    512-point calls of an 8-GPU Picard iteration.
  const bool cha_mlp = h->net_kind == DPI_NET_MLP && P.eq == DPI_EQ_CHA;
  bool baseline_forked = false;
  // (not with the TD estimators: their terminal stage evaluates u_k through the same per-point scratch)
  static const bool no_aux = getenv("DPI_NO_AUX_STREAM") != nullptr;  // developer switch: baseline pass on `stream`
  if (cha_mlp && !no_aux && cfg->estimate_delta_t == 0.f && !(cfg->precision == DPI_PREC_F16 && !h->tc_ready)) {
    if ((rc = ensure_scratch(h, n, h->mlp.HP))) return rc;
    if (!h->aux) {
      DPI_CUDA_CHECK(cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking));
      DPI_CUDA_CHECK(cudaEventCreateWithFlags(&h->aux_fork, cudaEventDisableTiming));
      DPI_CUDA_CHECK(cudaEventCreateWithFlags(&h->aux_join, cudaEventDisableTiming));
    }
    DPI_CUDA_CHECK(cudaEventRecord(h->aux_fork, st));  // tx is ready, fb / c0 of the previous call are consumed
    DPI_CUDA_CHECK(cudaStreamWaitEvent(h->aux, h->aux_fork, 0));
    {
      StageTimer tm(h, DPI_STAGE_BASELINE, h->aux);
      rc = launch_cha_simt_base(P, h->mlp, tx, n, h->fb, h->c0, nullptr, nullptr, cfg->precision == DPI_PREC_F16, h->aux);
    }
    DPI_CUDA_CHECK(cudaEventRecord(h->aux_join, h->aux));  // (joined below even when the launch failed)
    baseline_forked = true;
    if (rc) {
      cudaStreamWaitEvent(st, h->aux_join, 0);
      return rc;
    }
    h->launches += 1;
  }
  // every return below joins the side stream first
  struct Join {
    dpi_handle* h;
    cudaStream_t st;
    bool armed;
    void now() {
      if (armed) cudaStreamWaitEvent(st, h->aux_join, 0);
      armed = false;
    }
    ~Join() { now(); }
  } join{h, st, baseline_forked};

  // 1) terminal estimator -> y
  if (cfg->estimate_delta_t > 0.f) {
    // TD form (data.py:928-952): u_k at the path ends of every point, chunk by chunk -- rows, value pass of the frozen
    // network, then the terminal kernel reading those values where the step ends before T.  Not a tuned path (no
    // BASELINE config sets ESTIMATE_DELTA_T): the rows go through HBM.
    if (P.eq == DPI_EQ_GBM) {
      set_error("dpi_label_gen: ESTIMATE_DELTA_T > 0 is defined for Cha and OUProcessEquation (data.py:533)");
      return DPI_ERR_UNSUPPORTED;
    }
    StageTimer tm(h, DPI_STAGE_TERMINAL, st);
    const int d1 = 1 + P.nx;
    const int64_t per_point = (int64_t)cfg->M_T * (d1 + 1) * (int64_t)sizeof(float);
    const int64_t cp = std::max<int64_t>(1, std::min<int64_t>(n, (256LL << 20) / per_point));
    const int64_t need = cp * per_point;
    if (h->td_scratch_bytes < need) {
      if (h->td_scratch) cudaFree(h->td_scratch);
      h->td_scratch = nullptr, h->td_scratch_bytes = 0;
      DPI_CUDA_CHECK(cudaMalloc(&h->td_scratch, (size_t)need));
      h->td_scratch_bytes = need;
    }
    float* rows = reinterpret_cast<float*>(h->td_scratch);
    float* uvals = rows + cp * cfg->M_T * d1;
    for (int64_t p0 = 0; p0 < n; p0 += cp) {
      const int64_t np = std::min(cp, n - p0);
      const float* zTc = draws ? draws->zT + p0 * cfg->M_T * (int64_t)P.nx : nullptr;
      rc = launch_td_rows(P, tx + p0 * d1, np, cfg->M_T, cfg->seed, cfg->point_index_base + p0, zTc, rows, st);
      if (rc) return rc;
      rc = dpi_net_eval(h, rows, np * cfg->M_T, uvals, nullptr, stream);
      if (rc) return rc;
      rc = launch_terminal(P, tx + p0 * d1, np, cfg->M_T, cfg->seed, cfg->point_index_base + p0, zTc, y + p0 * d1, st, uvals);
      if (rc) return rc;
      h->launches += 2;
    }
  } else {
    StageTimer tm(h, DPI_STAGE_TERMINAL, st);
    rc = launch_terminal(P, tx, n, cfg->M_T, cfg->seed, cfg->point_index_base, draws ? draws->zT : nullptr, y, st);
    if (rc) return rc;
    h->launches += 1;
  }

  // 2) integral estimator: y += ..., then clip
  if (P.eq == DPI_EQ_GBM && (h->net_kind == DPI_NET_ZERO || h->net_kind == DPI_NET_MLP)) {
    if (P.sdgd_v < 1) {
      // has_hessian_term without an approximation context takes the full-Hessian branch (data.py:1262-1272),
      // which no BASELINE config selects
      set_error("dpi_label_gen: GBMEquationComplexExact needs HESSIAN_APPROXIMATION.method=SDGD (sdgd_v >= 1)");
      return DPI_ERR_UNSUPPORTED;
    }
    StageTimer tm(h, DPI_STAGE_INTEGRAL, st);
    if (cfg->precision != DPI_PREC_FP32) {
      // 16-bit tensor-core path: hidden layers of the 2-jets as tcgen05 GEMMs (u_k == 0 has no network to evaluate)
      if (h->net_kind == DPI_NET_MLP) {
        static const bool force_v1 = getenv("DPI_GBM_TC_V1") != nullptr;  // developer switch: the bulk-synchronous kernel
        static const bool force_v2 = getenv("DPI_GBM_TC_V2") != nullptr;  // developer switch: gathered directions, value chain from shared memory
        if (!force_v1 && !force_v2 && gbm_tc3_eligible(P, h->mlp)) {
          if (draws && !draws->idx) {
            set_error("deterministic mode with SDGD needs the idx draws");
            return DPI_ERR_INVALID;
          }
          rc = launch_gbm_tc3(P, h->mlp, tx, n, cfg->M_I, cfg->seed, cfg->point_index_base, draws ? draws->u_s : nullptr,
                              draws ? draws->zS : nullptr, draws ? draws->idx : nullptr, bound, y, st);
          if (!rc) h->launches += 1;
          return rc;
        }
        if (!force_v1 && gbm_tc2_eligible(P, h->mlp)) {
          if (draws && !draws->idx) {
            set_error("deterministic mode with SDGD needs the idx draws");
            return DPI_ERR_INVALID;
          }
          rc = launch_gbm_tc2(P, h->mlp, tx, n, cfg->M_I, cfg->seed, cfg->point_index_base, draws ? draws->u_s : nullptr,
                              draws ? draws->zS : nullptr, draws ? draws->idx : nullptr, bound, y, st);
          if (!rc) h->launches += 1;
          return rc;
        }
        rc = launch_gbm_tc(P, h->mlp, tx, n, cfg->M_I, cfg->seed, cfg->point_index_base, draws ? draws->u_s : nullptr,
                           draws ? draws->zS : nullptr, draws ? draws->idx : nullptr, bound, y, st);
        if (!rc) h->launches += 1;
        return rc;
      }
    }
    rc = launch_gbm_simt(P, h->mlp, h->net_kind == DPI_NET_ZERO, tx, n, cfg->M_I, cfg->seed, cfg->point_index_base,
                         draws ? draws->u_s : nullptr, draws ? draws->zS : nullptr, draws ? draws->idx : nullptr,
                         bound, y, st);
    if (!rc) h->launches += 1;
    return rc;
  }
  if (h->net_kind == DPI_NET_ZERO) {
    rc = launch_finish_zero(P, tx, n, bound, y, st);
    if (!rc) h->launches += 1;
    return rc;
  }
  if (h->net_kind == DPI_NET_MLP && P.eq == DPI_EQ_CHA) {
    const bool bf16 = cfg->precision == DPI_PREC_F16;  // 16-bit tensor-core path
    if (bf16 && !h->tc_ready) {
      set_error("dpi_label_gen: the 16-bit tensor-core path needs the MLP [1+d,128,128,128,128,1] with d <= 1024");
      return DPI_ERR_UNSUPPORTED;
    }
    if (baseline_forked) {
      join.now();  // fb / c0 of the baseline pass (side stream) are complete
    } else {
      if ((rc = ensure_scratch(h, n, h->mlp.HP))) return rc;
      {
        StageTimer tm(h, DPI_STAGE_BASELINE, st);
        rc = launch_cha_simt_base(P, h->mlp, tx, n, h->fb, h->c0, nullptr, nullptr, bf16, st);
      }
      if (rc) return rc;
      h->launches += 1;
    }
    StageTimer tm(h, DPI_STAGE_INTEGRAL, st);
    if (bf16) {
      // chunk the points so that the unit-partial scratch stays below ~256 MB
      const int64_t upp = tc_units_per_point(cfg->M_I);
      int64_t chunk = (int64_t)(256u << 20) / (int64_t)(sizeof(float) * kTcPartGFloatsPerUnit * upp);
      if (chunk < 1) chunk = 1;
      if (chunk > n) chunk = n;
      if ((rc = ensure_parts(h, chunk * upp))) return rc;
      if (tc_zring_bytes(h->tc) > h->tc_zring_size) {
        if (h->tc_zring) cudaFree(h->tc_zring);
        h->tc_zring = nullptr, h->tc_zring_size = 0;
        DPI_CUDA_CHECK(cudaMalloc(&h->tc_zring, tc_zring_bytes(h->tc)));
        h->tc_zring_size = tc_zring_bytes(h->tc);
      }
      const int d = P.nx;
      for (int64_t p0 = 0; p0 < n; p0 += chunk) {
        const int64_t np = (n - p0 < chunk) ? n - p0 : chunk;
        rc = launch_cha_tc_path(P, h->tc, tx + p0 * (1 + d), np, cfg->M_I, cfg->seed, cfg->point_index_base + p0,
                                draws ? draws->u_s + p0 * cfg->M_I : nullptr,
                                draws ? draws->zS + p0 * (int64_t)cfg->M_I * d : nullptr, h->fb + p0,
                                h->c0 + p0 * h->mlp.HP, bound, h->partG, h->partF,
                                reinterpret_cast<int*>(h->partF + h->part_units * kTcPartFFloatsPerUnit), h->tc_zring,
                                h->nonfinite_flag, y + p0 * (1 + d), st);
        if (rc) return rc;
        h->launches += 2;
      }
      return DPI_OK;
    }
    rc = launch_cha_simt_path(P, h->mlp, tx, n, cfg->M_I, cfg->seed, cfg->point_index_base,
                              draws ? draws->u_s : nullptr, draws ? draws->zS : nullptr, h->fb, h->c0, bound, y, st);
    if (!rc) h->launches += 1;
    return rc;
  }
  if (h->net_kind == DPI_NET_PISGRAD && P.eq == DPI_EQ_OU) {
    if (cfg->precision == DPI_PREC_F16 && h->hjb_tc) {
      // fused tcgen05 kernel: fp16 operands, activations in shared memory / TMEM, weights streamed from L2
      if ((rc = ensure_scratch(h, n, 64))) return rc;
      const int64_t pf = hjb_tc_part_floats(n, cfg->M_I, P.nx);
      if (pf > h->hjb_part_floats) {
        if (h->hjb_part) cudaFree(h->hjb_part);
        h->hjb_part = nullptr, h->hjb_part_floats = 0;
        DPI_CUDA_CHECK(cudaMalloc((void**)&h->hjb_part, sizeof(float) * pf));
        h->hjb_part_floats = pf;
      }
      StageTimer tm(h, DPI_STAGE_INTEGRAL, st);
      return launch_hjb_tc_paths(P, h->hjb_tc, tx, n, cfg->M_I, cfg->seed, cfg->point_index_base,
                                 draws ? draws->u_s : nullptr, draws ? draws->zS : nullptr, bound, h->fb, h->hjb_part,
                                 h->nonfinite_flag, y, st, &h->launches);
    }
    // DPI_PREC_TF32 (or an ineligible shape): the batched-GEMM pipeline with its wide GEMMs on tcgen05, TF32 operands
    PisDev pis = h->pis;
    pis.tf32 = cfg->precision != DPI_PREC_FP32;
    const int64_t ppc = std::min<int64_t>(n, hjb_points_per_chunk(h->pis, P.nx, cfg->M_I, kHjbScratchBudget));
    if ((rc = ensure_hjb_scratch(h, hjb_scratch_floats(h->pis, P.nx, ppc * (cfg->M_I + 1), ppc)))) return rc;
    StageTimer tm(h, DPI_STAGE_INTEGRAL, st);
    return launch_hjb_paths(P, pis, tx, n, cfg->M_I, cfg->seed, cfg->point_index_base, draws ? draws->u_s : nullptr,
                            draws ? draws->zS : nullptr, bound, h->hjb_scratch, ppc, y, st, &h->launches);
  }
  set_error("dpi_label_gen: equation %d with network kind %d is not built yet", P.eq, h->net_kind);
  return DPI_ERR_UNSUPPORTED;
}

int dpi_label_gen_host(dpi_handle* h, const float* tx_in_host, int64_t n, const dpi_label_config* cfg,
                       float* tx_out_host, float* y_host, void* stream) {
  if (!h || !cfg || !y_host || n < 0) {
    set_error("dpi_label_gen_host: bad arguments");
    return DPI_ERR_INVALID;
  }
  if (!h->has_problem) {
    set_error("dpi_label_gen_host: problem not set");
    return DPI_ERR_STATE;
  }
  if (n == 0) return DPI_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  const int64_t floats = n * (1 + h->P.nx);
  int rc = ensure_stage(h, floats);
  if (rc) return rc;
  // Large calls (the 2^20-point sweep moves 4.2 GB each way) are pipelined: the points go through in chunks of
  // ~kHostPipeChunkBytes, chunk k's kernels on `stream` overlapping the H2D copy of chunk k + 1 and the D2H copy of
  // chunk k - 1 on two copy streams of the handle.  Labels do not depend on how a call is split (Philox counters are
  // keyed by the global point index), so the result is bitwise the one-shot result.
  constexpr int64_t kHostPipeChunkBytes = 64ll << 20, kHostPipeMinBytes = 256ll << 20;
  if ((int64_t)sizeof(float) * floats >= kHostPipeMinBytes && !h->timing) {
    const int64_t ld = 1 + h->P.nx;
    const int64_t cp = std::max<int64_t>(1024, kHostPipeChunkBytes / ((int64_t)sizeof(float) * ld));
    const int64_t nchunk = (n + cp - 1) / cp;
    if (!h->host_in) DPI_CUDA_CHECK(cudaStreamCreateWithFlags(&h->host_in, cudaStreamNonBlocking));
    if (!h->host_out) DPI_CUDA_CHECK(cudaStreamCreateWithFlags(&h->host_out, cudaStreamNonBlocking));
    while ((int64_t)h->host_ev.size() < 2 * nchunk + 2) {
      cudaEvent_t e;
      DPI_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->host_ev.push_back(e);
    }
    cudaEvent_t ev_start = h->host_ev[2 * nchunk], ev_done = h->host_ev[2 * nchunk + 1];
    DPI_CUDA_CHECK(cudaEventRecord(ev_start, st));  // the staging buffers are free once earlier work on `stream` is done
    DPI_CUDA_CHECK(cudaStreamWaitEvent(h->host_in, ev_start, 0));
    DPI_CUDA_CHECK(cudaStreamWaitEvent(h->host_out, ev_start, 0));
    if (!tx_in_host) {
      rc = dpi_sample_tx(h, n, cfg->point_index_base, cfg->seed, h->stage_tx, stream);
      if (rc) return rc;
    }
    for (int64_t k = 0; k < nchunk; ++k) {
      const int64_t p0 = k * cp, np = std::min(cp, n - p0);
      const size_t off = (size_t)(p0 * ld), bytes = sizeof(float) * (size_t)(np * ld);
      if (tx_in_host) {
        DPI_CUDA_CHECK(cudaMemcpyAsync(h->stage_tx + off, tx_in_host + off, bytes, cudaMemcpyHostToDevice, h->host_in));
        DPI_CUDA_CHECK(cudaEventRecord(h->host_ev[2 * k], h->host_in));
        DPI_CUDA_CHECK(cudaStreamWaitEvent(st, h->host_ev[2 * k], 0));
      }
      dpi_label_config ck = *cfg;
      ck.point_index_base = cfg->point_index_base + p0;
      rc = dpi_label_gen(h, h->stage_tx + off, np, &ck, nullptr, h->stage_y + off, stream);
      if (rc) {  // nothing of this call may still be writing the caller's buffers when the error is reported
        cudaStreamSynchronize(h->host_in);
        cudaStreamSynchronize(st);
        cudaStreamSynchronize(h->host_out);
        return rc;
      }
      DPI_CUDA_CHECK(cudaEventRecord(h->host_ev[2 * k + 1], st));
      DPI_CUDA_CHECK(cudaStreamWaitEvent(h->host_out, h->host_ev[2 * k + 1], 0));
      DPI_CUDA_CHECK(cudaMemcpyAsync(y_host + off, h->stage_y + off, bytes, cudaMemcpyDeviceToHost, h->host_out));
      if (tx_out_host)
        DPI_CUDA_CHECK(cudaMemcpyAsync(tx_out_host + off, h->stage_tx + off, bytes, cudaMemcpyDeviceToHost, h->host_out));
    }
    DPI_CUDA_CHECK(cudaEventRecord(ev_done, h->host_out));
    DPI_CUDA_CHECK(cudaStreamWaitEvent(st, ev_done, 0));  // `stream` (synchronised below) covers the last copies
    return dpi_check_labels(h, stream);
  }
  if (tx_in_host) {
    DPI_CUDA_CHECK(cudaMemcpyAsync(h->stage_tx, tx_in_host, sizeof(float) * floats, cudaMemcpyHostToDevice, st));
  } else {
    rc = dpi_sample_tx(h, n, cfg->point_index_base, cfg->seed, h->stage_tx, stream);
    if (rc) return rc;
  }
  rc = dpi_label_gen(h, h->stage_tx, n, cfg, nullptr, h->stage_y, stream);
  if (rc) return rc;
  DPI_CUDA_CHECK(cudaMemcpyAsync(y_host, h->stage_y, sizeof(float) * floats, cudaMemcpyDeviceToHost, st));
  if (tx_out_host)
    DPI_CUDA_CHECK(cudaMemcpyAsync(tx_out_host, h->stage_tx, sizeof(float) * floats, cudaMemcpyDeviceToHost, st));
  return dpi_check_labels(h, stream);  // synchronises the stream; DPI_ERR_NONFINITE instead of silent inf / NaN labels
}

int dpi_net_eval(dpi_handle* h, const float* tx, int64_t n, float* u, float* du, void* stream) {
  if (!h || !h->has_problem || !h->has_net) {
    set_error("dpi_net_eval: problem/network not set");
    return DPI_ERR_STATE;
  }
  if (n <= 0) return DPI_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  if (h->net_kind == DPI_NET_MLP && h->P.eq == DPI_EQ_CHA && du == nullptr) {
    int rc = ensure_scratch(h, n, h->mlp.HP);
    if (rc) return rc;
    rc = launch_cha_simt_base(h->P, h->mlp, tx, n, h->fb, h->c0, u, nullptr, false, st);
    if (!rc) h->launches += 1;
    return rc;
  }
  if (h->net_kind == DPI_NET_MLP) {
    // value and full input gradient through the batched-GEMM pipeline (any equation, any hidden widths)
    const int64_t per_row = mlp_eval_scratch_floats(h->mlp_raw, h->n_mlp_raw, 1) * (int64_t)sizeof(float);
    const int64_t rpc = std::max<int64_t>(1, std::min<int64_t>(n, kHjbScratchBudget / per_row));
    int rc = ensure_hjb_scratch(h, mlp_eval_scratch_floats(h->mlp_raw, h->n_mlp_raw, rpc));
    if (rc) return rc;
    return launch_mlp_eval(h->mlp_raw, h->n_mlp_raw, h->P.nx, tx, n, h->hjb_scratch, rpc, u, du, st, &h->launches);
  }
  if (h->net_kind == DPI_NET_ZERO) {
    if (u) DPI_CUDA_CHECK(cudaMemsetAsync(u, 0, sizeof(float) * n, st));
    if (du) DPI_CUDA_CHECK(cudaMemsetAsync(du, 0, sizeof(float) * n * h->P.nx, st));
    return DPI_OK;
  }
  if (h->net_kind == DPI_NET_PISGRAD && h->P.eq == DPI_EQ_OU) {
    const int64_t per_row = hjb_scratch_floats(h->pis, h->P.nx, 1, 0) * (int64_t)sizeof(float);
    const int64_t rpc = std::max<int64_t>(1, std::min<int64_t>(n, kHjbScratchBudget / per_row));
    int rc = ensure_hjb_scratch(h, hjb_scratch_floats(h->pis, h->P.nx, rpc, 0));
    if (rc) return rc;
    return launch_hjb_eval(h->P, h->pis, tx, n, h->hjb_scratch, rpc, u, du, st, &h->launches);
  }
  set_error("dpi_net_eval: full input gradient not built yet");
  return DPI_ERR_UNSUPPORTED;
}

int dpi_net_hessian_diag(dpi_handle* h, const float* tx, int64_t n, float* d2u, void* stream) {
  if (!h || !h->has_problem || !h->has_net) {
    set_error("dpi_net_hessian_diag: problem/network not set");
    return DPI_ERR_STATE;
  }
  if (n <= 0) return DPI_OK;
  if (!tx || !d2u) {
    set_error("dpi_net_hessian_diag: NULL buffer");
    return DPI_ERR_INVALID;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  if (h->net_kind == DPI_NET_ZERO) {
    DPI_CUDA_CHECK(cudaMemsetAsync(d2u, 0, sizeof(float) * n * h->P.nx, st));
    return DPI_OK;
  }
  if (h->net_kind != DPI_NET_MLP) {
    set_error("dpi_net_hessian_diag: built for plain MLPs (the fully nonlinear config)");
    return DPI_ERR_UNSUPPORTED;
  }
  int rc = launch_mlp_hess_diag(h->P, h->mlp, tx, n, d2u, st);
  if (!rc) h->launches += 1;
  return rc;
}

int dpi_net_hessian(dpi_handle* h, const float* tx, int64_t n, float* d2u, void* stream) {
  if (!h || !h->has_problem || !h->has_net) {
    set_error("dpi_net_hessian: problem/network not set");
    return DPI_ERR_STATE;
  }
  if (n <= 0) return DPI_OK;
  if (!tx || !d2u) {
    set_error("dpi_net_hessian: NULL buffer");
    return DPI_ERR_INVALID;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  if (h->net_kind == DPI_NET_ZERO) {
    DPI_CUDA_CHECK(cudaMemsetAsync(d2u, 0, sizeof(float) * n * h->P.nx * h->P.nx, st));
    return DPI_OK;
  }
  if (h->net_kind != DPI_NET_MLP) {
    set_error("dpi_net_hessian: built for plain MLPs (the fully nonlinear config)");
    return DPI_ERR_UNSUPPORTED;
  }
  int rc = launch_mlp_hess_full(h->P, h->mlp, tx, n, d2u, st);
  if (!rc) h->launches += 1;
  return rc;
}

int dpi_train_init(dpi_handle* h, const dpi_network* net, int64_t max_batch, void* stream) {
  if (!h || !net || max_batch < 1) {
    set_error("dpi_train_init: bad arguments");
    return DPI_ERR_INVALID;
  }
  if (net->kind != DPI_NET_MLP) {
    set_error("dpi_train_init: the regression step is built for plain MLPs (NETWORK.PISGRADNET false)");
    return DPI_ERR_UNSUPPORTED;
  }
  const int L = net->n_main;
  if (L < 2 || L > DPI_MAX_LINEAR || net->main[L - 1].n_out != 1) {
    set_error("dpi_train_init: need 2..%d Linear layers ending in a scalar output (NETWORK.TYPE = Value)", DPI_MAX_LINEAR);
    return DPI_ERR_INVALID;
  }
  for (int l = 0; l < L; ++l) {
    const dpi_linear& ll = net->main[l];
    if (!ll.W || !ll.b || ll.n_out < 1 || (l > 0 && ll.n_in != net->main[l - 1].n_out)) {
      set_error("dpi_train_init: layer %d has an inconsistent shape [%d,%d]", l, ll.n_out, ll.n_in);
      return DPI_ERR_INVALID;
    }
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  DPI_CUDA_CHECK(cudaStreamSynchronize(st));
  if (h->train_m) cudaFree(h->train_m);
  if (h->train_v) cudaFree(h->train_v);
  if (h->train_scratch) cudaFree(h->train_scratch);
  for (int r = 0; r < DPI_DP_MAX_WORLD; ++r)
    if (h->dp.mapped[r]) cudaIpcCloseMemHandle(h->dp.mapped[r]);
  h->dp = DpLinks{};
  if (h->train_grad) cudaFree(h->train_grad);
  if (h->train_graph) cudaGraphExecDestroy(h->train_graph);
  if (h->train_stage_tx) cudaFree(h->train_stage_tx);
  if (h->train_stage_y) cudaFree(h->train_stage_y);
  h->train_graph = nullptr;
  h->train_graph_batch = 0;
  h->train_stage_tx = h->train_stage_y = nullptr;
  h->train_m = h->train_v = h->train_scratch = h->train_grad = nullptr;
  for (int l = 0; l < L; ++l) h->train_lin[l] = net->main[l];
  h->n_train_lin = L;
  h->train_params = train_param_count(h->train_lin, L);
  h->train_max_batch = max_batch;
  DPI_CUDA_CHECK(cudaMalloc((void**)&h->train_m, sizeof(float) * h->train_params));
  DPI_CUDA_CHECK(cudaMalloc((void**)&h->train_v, sizeof(float) * h->train_params));
  DPI_CUDA_CHECK(cudaMalloc((void**)&h->train_scratch, sizeof(float) * train_scratch_floats(h->train_lin, L, max_batch)));
  // [gradient P | loss terms 4 | pad 4 | flag, epoch, done (ints) ...]: one block, so that peers map all of it at once
  DPI_CUDA_CHECK(cudaMalloc((void**)&h->train_grad, sizeof(float) * (h->train_params + 16)));
  DPI_CUDA_CHECK(cudaMemset(h->train_grad, 0, sizeof(float) * (h->train_params + 16)));
  h->dp = DpLinks{};
  DPI_CUDA_CHECK(cudaMalloc((void**)&h->train_stage_tx, sizeof(float) * max_batch * net->main[0].n_in));
  DPI_CUDA_CHECK(cudaMalloc((void**)&h->train_stage_y, sizeof(float) * max_batch * net->main[0].n_in));
  if (!h->train_step_dev) DPI_CUDA_CHECK(cudaMalloc((void**)&h->train_step_dev, sizeof(long long)));
  if (!h->train_stream) {
    DPI_CUDA_CHECK(cudaStreamCreateWithFlags(&h->train_stream, cudaStreamNonBlocking));
    DPI_CUDA_CHECK(cudaEventCreateWithFlags(&h->train_ev_in, cudaEventDisableTiming));
    DPI_CUDA_CHECK(cudaEventCreateWithFlags(&h->train_ev_out, cudaEventDisableTiming));
    DPI_CUDA_CHECK(cudaStreamCreateWithFlags(&h->train_aux.side, cudaStreamNonBlocking));
    for (int l = 0; l < DPI_MAX_LINEAR; ++l)
      DPI_CUDA_CHECK(cudaEventCreateWithFlags(&h->train_aux.fork[l], cudaEventDisableTiming));
    DPI_CUDA_CHECK(cudaEventCreateWithFlags(&h->train_aux.join, cudaEventDisableTiming));
    for (int i = 0; i < 3; ++i) DPI_CUDA_CHECK(cudaEventCreateWithFlags(&h->train_aux.pre[i], cudaEventDisableTiming));
  }
  DPI_CUDA_CHECK(cudaMemsetAsync(h->train_m, 0, sizeof(float) * h->train_params, st));
  DPI_CUDA_CHECK(cudaMemsetAsync(h->train_v, 0, sizeof(float) * h->train_params, st));
  return DPI_OK;
}

int64_t dpi_train_param_count(const dpi_handle* h) { return h ? h->train_params : 0; }

int dpi_train_reset(dpi_handle* h, void* stream) {
  if (!h || !h->n_train_lin) {
    set_error("dpi_train_reset: call dpi_train_init first");
    return DPI_ERR_STATE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  DPI_CUDA_CHECK(cudaMemsetAsync(h->train_m, 0, sizeof(float) * h->train_params, st));
  DPI_CUDA_CHECK(cudaMemsetAsync(h->train_v, 0, sizeof(float) * h->train_params, st));
  return DPI_OK;
}

int dpi_train_grad(dpi_handle* h, const float* tx, const float* y, int64_t batch, const dpi_train_config* cfg,
                   float* grad_flat, float* losses, void* stream) {
  if (!h || !h->n_train_lin) {
    set_error("dpi_train_grad: call dpi_train_init first");
    return DPI_ERR_STATE;
  }
  if (!tx || !y || !cfg || !grad_flat || !losses || batch < 1 || batch > h->train_max_batch) {
    set_error("dpi_train_grad: bad arguments (batch=%lld, max_batch=%lld)", (long long)batch, (long long)h->train_max_batch);
    return DPI_ERR_INVALID;
  }
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  return launch_train_grad(h->train_lin, h->n_train_lin, tx, y, batch, cfg->loss_beta, cfg->fixed_weight, h->train_scratch,
                           grad_flat, losses, reinterpret_cast<cudaStream_t>(stream), &h->launches, &h->train_aux);
}

int dpi_train_adam(dpi_handle* h, const float* grad_flat, const dpi_train_config* cfg, int64_t step, void* stream) {
  if (!h || !h->n_train_lin) {
    set_error("dpi_train_adam: call dpi_train_init first");
    return DPI_ERR_STATE;
  }
  if (!grad_flat || !cfg || step < 1) {
    set_error("dpi_train_adam: bad arguments");
    return DPI_ERR_INVALID;
  }
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  return launch_train_adam(h->train_lin, h->n_train_lin, grad_flat, h->train_m, h->train_v, cfg->lr, cfg->beta1, cfg->beta2,
                           cfg->eps, step, reinterpret_cast<cudaStream_t>(stream), &h->launches);
}

int dpi_train_fit(dpi_handle* h, const float* tx, const float* y, int64_t n, int64_t batch, int64_t n_epochs,
                  const dpi_train_config* cfg, int64_t first_step, float* losses_last, void* stream) {
  if (!h || !h->n_train_lin) {
    set_error("dpi_train_fit: call dpi_train_init first");
    return DPI_ERR_STATE;
  }
  if (!tx || !y || !cfg || n < 1 || batch < 1 || batch > h->train_max_batch || n_epochs < 0 || first_step < 1) {
    set_error("dpi_train_fit: bad arguments (n=%lld batch=%lld max_batch=%lld)", (long long)n, (long long)batch,
              (long long)h->train_max_batch);
    return DPI_ERR_INVALID;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  const int ld = h->train_lin[0].n_in;
  float* losses = h->train_grad + h->train_params;
  cudaStream_t ts = h->train_stream;
  // the internal stream picks up after everything queued on the caller's stream (labels, previous fit)
  DPI_CUDA_CHECK(cudaEventRecord(h->train_ev_in, st));
  DPI_CUDA_CHECK(cudaStreamWaitEvent(ts, h->train_ev_in, 0));
  const long long step0 = (long long)first_step;
  DPI_CUDA_CHECK(cudaMemcpyAsync(h->train_step_dev, &step0, sizeof(long long), cudaMemcpyHostToDevice, ts));
  DPI_CUDA_CHECK(cudaStreamSynchronize(ts));  // step0 lives on this call's stack
  auto same_cfg = [&](const dpi_train_config& c) { return std::memcmp(&c, cfg, sizeof(dpi_train_config)) == 0; };
  for (int64_t e = 0; e < n_epochs; ++e)
    for (int64_t i = 0; i < n; i += batch) {
      const int64_t b = std::min<int64_t>(batch, n - i);
      DPI_CUDA_CHECK(cudaMemcpyAsync(h->train_stage_tx, tx + i * ld, sizeof(float) * b * ld, cudaMemcpyDeviceToDevice, ts));
      DPI_CUDA_CHECK(cudaMemcpyAsync(h->train_stage_y, y + i * ld, sizeof(float) * b * ld, cudaMemcpyDeviceToDevice, ts));
      if (b == batch) {
        if (!h->train_graph || h->train_graph_batch != b || !same_cfg(h->train_graph_cfg)) {
          // (re)capture one optimiser step: ~45 small kernels whose launch cost would otherwise dominate
          if (h->train_graph) cudaGraphExecDestroy(h->train_graph);
          h->train_graph = nullptr;
          cudaGraph_t graph = nullptr;
          int64_t nodes = 0;
          DPI_CUDA_CHECK(cudaStreamBeginCapture(ts, cudaStreamCaptureModeRelaxed));
          int rc = launch_train_grad(h->train_lin, h->n_train_lin, h->train_stage_tx, h->train_stage_y, b, cfg->loss_beta,
                                     cfg->fixed_weight, h->train_scratch, h->train_grad, losses, ts, &nodes, &h->train_aux);
          if (!rc)
            rc = h->dp.world > 1
                     ? launch_train_dp_reduce_adam(h->train_lin, h->n_train_lin, h->dp, h->train_m, h->train_v, cfg->lr,
                                                   cfg->beta1, cfg->beta2, cfg->eps, h->train_step_dev, ts, &nodes)
                     : launch_train_adam_dev(h->train_lin, h->n_train_lin, h->train_grad, h->train_m, h->train_v, cfg->lr,
                                             cfg->beta1, cfg->beta2, cfg->eps, h->train_step_dev, ts, &nodes);
          cudaError_t ce = cudaStreamEndCapture(ts, &graph);
          if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
          }
          DPI_CUDA_CHECK(ce);
          DPI_CUDA_CHECK(cudaGraphInstantiate(&h->train_graph, graph, 0));
          cudaGraphDestroy(graph);
          h->train_graph_batch = b;
          h->train_graph_nodes = nodes;
          h->train_graph_cfg = *cfg;
        }
        DPI_CUDA_CHECK(cudaGraphLaunch(h->train_graph, ts));
        h->launches += h->train_graph_nodes;
      } else {  // ragged last mini-batch: plain launches
        int rc = launch_train_grad(h->train_lin, h->n_train_lin, h->train_stage_tx, h->train_stage_y, b, cfg->loss_beta,
                                   cfg->fixed_weight, h->train_scratch, h->train_grad, losses, ts, &h->launches, &h->train_aux);
        if (rc) return rc;
        rc = h->dp.world > 1
                 ? launch_train_dp_reduce_adam(h->train_lin, h->n_train_lin, h->dp, h->train_m, h->train_v, cfg->lr, cfg->beta1,
                                               cfg->beta2, cfg->eps, h->train_step_dev, ts, &h->launches)
                 : launch_train_adam_dev(h->train_lin, h->n_train_lin, h->train_grad, h->train_m, h->train_v, cfg->lr, cfg->beta1,
                                         cfg->beta2, cfg->eps, h->train_step_dev, ts, &h->launches);
        if (rc) return rc;
      }
    }
  if (losses_last)
    DPI_CUDA_CHECK(cudaMemcpyAsync(losses_last, losses, 3 * sizeof(float), cudaMemcpyDeviceToDevice, ts));
  DPI_CUDA_CHECK(cudaEventRecord(h->train_ev_out, ts));
  DPI_CUDA_CHECK(cudaStreamWaitEvent(st, h->train_ev_out, 0));
  return DPI_OK;
}

int dpi_train_dp_export(dpi_handle* h, void* handle_out) {
  if (!h || !h->n_train_lin || !handle_out) {
    set_error("dpi_train_dp_export: call dpi_train_init first");
    return DPI_ERR_STATE;
  }
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == DPI_DP_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t mh;
  DPI_CUDA_CHECK(cudaIpcGetMemHandle(&mh, h->train_grad));
  std::memcpy(handle_out, &mh, sizeof(mh));
  return DPI_OK;
}

int dpi_train_dp_connect(dpi_handle* h, int32_t rank, int32_t world, const void* handles) {
  if (!h || !h->n_train_lin || !handles) {
    set_error("dpi_train_dp_connect: call dpi_train_init first");
    return DPI_ERR_STATE;
  }
  if (world < 1 || world > DPI_DP_MAX_WORLD || rank < 0 || rank >= world) {
    set_error("dpi_train_dp_connect: rank %d of %d (at most %d ranks of one node)", rank, world, DPI_DP_MAX_WORLD);
    return DPI_ERR_INVALID;
  }
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  DPI_CUDA_CHECK(cudaDeviceSynchronize());
  for (int r = 0; r < DPI_DP_MAX_WORLD; ++r)
    if (h->dp.mapped[r]) cudaIpcCloseMemHandle(h->dp.mapped[r]);
  h->dp = DpLinks{};
  if (h->train_graph) cudaGraphExecDestroy(h->train_graph);  // the captured step names the optimiser kernel
  h->train_graph = nullptr;
  h->train_graph_batch = 0;
  const int64_t P = h->train_params;
  int* ctl = reinterpret_cast<int*>(h->train_grad + P + 8);
  DPI_CUDA_CHECK(cudaMemset(ctl, 0, 8 * sizeof(int)));
  h->dp.world = world, h->dp.rank = rank;
  h->dp.epoch = ctl + 1, h->dp.done = ctl + 2;
  for (int r = 0; r < world; ++r) {
    float* base = h->train_grad;
    if (r != rank) {
      cudaIpcMemHandle_t mh;
      std::memcpy(&mh, reinterpret_cast<const char*>(handles) + (size_t)r * sizeof(mh), sizeof(mh));
      void* p = nullptr;
      DPI_CUDA_CHECK(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
      h->dp.mapped[r] = p;
      base = reinterpret_cast<float*>(p);
    }
    h->dp.grad[r] = base;
    h->dp.flag[r] = reinterpret_cast<int*>(base + P + 8);
  }
  return DPI_OK;
}

int64_t dpi_launch_count(const dpi_handle* h) { return h ? h->launches : 0; }

int dpi_selftest_umma(dpi_handle* h, const float* A, const float* B, float* D, void* stream) {
  if (!h || !A || !B || !D) {
    set_error("dpi_selftest_umma: NULL argument");
    return DPI_ERR_INVALID;
  }
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  return tc_umma_selftest(A, B, D, reinterpret_cast<cudaStream_t>(stream));
}

int dpi_set_stage_timing(dpi_handle* h, int enable) {
  if (!h) {
    set_error("dpi_set_stage_timing: NULL handle");
    return DPI_ERR_INVALID;
  }
  DPI_CUDA_CHECK(cudaSetDevice(h->device));
  if (enable && !h->ev[0][0])
    for (int s = 0; s < DPI_STAGE_COUNT; ++s)
      for (int e = 0; e < 2; ++e) DPI_CUDA_CHECK(cudaEventCreate(&h->ev[s][e]));
  h->timing = enable != 0;
  return DPI_OK;
}

int dpi_stage_ms(dpi_handle* h, int stage, float* ms) {
  if (!h || !ms || stage < 0 || stage >= DPI_STAGE_COUNT) {
    set_error("dpi_stage_ms: bad arguments");
    return DPI_ERR_INVALID;
  }
  if (!h->timing || !h->ev_valid[stage]) {
    set_error("dpi_stage_ms: stage %d was not timed by the last dpi_label_gen call", stage);
    return DPI_ERR_STATE;
  }
  DPI_CUDA_CHECK(cudaEventSynchronize(h->ev[stage][1]));
  DPI_CUDA_CHECK(cudaEventElapsedTime(ms, h->ev[stage][0], h->ev[stage][1]));
  return DPI_OK;
}

}  // extern "C"
