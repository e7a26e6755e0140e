"""Torch-facing wrapper of the C-ABI handle, and the `picard_b200::label_gen` custom op.

PyTorch is plumbing here (device memory, streams); all arithmetic happens in libdpi_b200.so.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as nat

_ENGINES: Dict[int, "LabelEngine"] = {}
_NEXT_ID = [1]


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class LabelEngine:
    """One handle per device. Holds the packed equation and frozen network u_k."""

    def __init__(self, device="cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("deeppicarditeration_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"LabelEngine device must be cuda, got {device}")
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", self.index)
        self.lib = nat.load()
        h = C.c_void_p()
        nat.check(self.lib.dpi_create(self.index, C.byref(h)), "dpi_create")
        self.h = h
        self.nx = None
        self._keep: List[torch.Tensor] = []
        self._prob_keep = []
        self.id = _NEXT_ID[0]
        _NEXT_ID[0] += 1
        _ENGINES[self.id] = self

    def close(self):
        if getattr(self, "h", None):
            self.lib.dpi_destroy(self.h)
            self.h = None
            _ENGINES.pop(self.id, None)

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------------------------------
    def set_problem(self, kind: str, nx: int, T: float, alpha: float, *, cha_k: float = 0.0, gbm_w=None, gbm_v=None,
                    ou_mean=None, ou_diagvar=None, ou_pi=None, ou_theta: float = 1.0, ou_mu: float = 0.0,
                    ou_alpha_init: float = 0.0, sdgd_v: int = 0):
        p = nat.dpi_problem()
        p.kind = {"Cha": nat.DPI_EQ_CHA, "GBM": nat.DPI_EQ_GBM, "OU": nat.DPI_EQ_OU}[kind]
        p.nx, p.T, p.alpha, p.cha_k = int(nx), float(T), float(alpha), float(cha_k)
        p.sdgd_v = int(sdgd_v)
        keep = []

        def dptr(a):
            arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
            keep.append(arr)
            return arr.ctypes.data_as(C.POINTER(C.c_double))

        if kind == "GBM":
            w = np.asarray(gbm_w, dtype=np.float64)
            p.gbm_nodes = int(w.shape[0])
            p.gbm_w, p.gbm_v = dptr(w), dptr(np.asarray(gbm_v, dtype=np.float64).reshape(-1))
        if kind == "OU":
            m = np.asarray(ou_mean, dtype=np.float64)
            p.ou_K = int(m.shape[0])
            p.ou_mean, p.ou_diagvar, p.ou_pi = dptr(m), dptr(ou_diagvar), dptr(ou_pi)
            p.ou_theta, p.ou_mu, p.ou_alpha_init = float(ou_theta), float(ou_mu), float(ou_alpha_init)
        with torch.cuda.device(self.device):
            nat.check(self.lib.dpi_set_problem(self.h, C.byref(p)), "dpi_set_problem")
        self.nx = int(nx)
        self.kind = kind
        self.sdgd_v = int(sdgd_v)

    def _fill_chain(self, arr, lins: Sequence[Tuple[torch.Tensor, torch.Tensor]]) -> int:
        if len(lins) > nat.DPI_MAX_LINEAR:
            raise NotImplementedError(f"at most {nat.DPI_MAX_LINEAR} Linear layers per chain")
        for i, (W, b) in enumerate(lins):
            W, b = _f32c(W, self.device), _f32c(b, self.device)
            self._keep += [W, b]
            arr[i].W, arr[i].b = W.data_ptr(), b.data_ptr()
            arr[i].n_out, arr[i].n_in = int(W.shape[0]), int(W.shape[1])
        return len(lins)

    def set_network(self, kind: str, main=(), t_encoder=(), smooth_net=(), phase=None, coeff=None):
        """kind: 'zero' | 'mlp' | 'pisgrad'; chains are lists of (W[out,in], b[out]) tensors."""
        net = nat.dpi_network()
        self._keep = []
        net.kind = {"zero": nat.DPI_NET_ZERO, "mlp": nat.DPI_NET_MLP, "pisgrad": nat.DPI_NET_PISGRAD}[kind]
        net.n_main = self._fill_chain(net.main, main)
        net.n_t_encoder = self._fill_chain(net.t_encoder, t_encoder)
        net.n_smooth = self._fill_chain(net.smooth_net, smooth_net)
        if phase is not None:
            ph, co = _f32c(phase.reshape(-1), self.device), _f32c(coeff.reshape(-1), self.device)
            self._keep += [ph, co]
            net.timestep_phase, net.timestep_coeff = ph.data_ptr(), co.data_ptr()
        with torch.cuda.device(self.device):
            nat.check(self.lib.dpi_set_network(self.h, C.byref(net), self._stream()), "dpi_set_network")
        self.net_kind = kind

    # ------------------------------------------------------------------------------------------
    def sample_tx(self, n: int, seed: int = 0, point_index_base: int = 0, *, n_factors: int = 0,
                  eps: float = 0.01) -> torch.Tensor:
        """(t, x) points.  n_factors = 0: t ~ U[eps, T - eps] (data.py:161-167); n_factors = k >= 1: the reference's
        default `sample_t`, t = T (1 - product of k uniforms) with k = N - i + 1 (data.py:149-159)."""
        tx = torch.empty(n, 1 + self.nx, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            nat.check(self.lib.dpi_sample_tx_ex(self.h, n, point_index_base, C.c_uint64(seed & (2**64 - 1)),
                                                int(n_factors), float(eps), _ptr(tx), self._stream()), "dpi_sample_tx_ex")
        return tx

    @staticmethod
    def _cfg(M_T, M_I, precision, bound, seed, base, delta_t=0.0):
        cfg = nat.dpi_label_config()
        cfg.estimate_delta_t = float(delta_t or 0.0)
        cfg.M_T, cfg.M_I = int(M_T), int(M_I)
        # "tf32": PISGradNet's layer-by-layer pipeline with TF32 tensor-core GEMMs; "bf16" maps to the distinct
        # DPI_PREC_BF16 value, which the library rejects (bf16 operands are not built) instead of aliasing fp16
        cfg.precision = {"fp32": nat.DPI_PREC_FP32, "f16": nat.DPI_PREC_F16, "fp16": nat.DPI_PREC_F16,
                         "tf32": nat.DPI_PREC_TF32, "bf16": nat.DPI_PREC_BF16}[precision]
        cfg.sample_bound = float(bound) if bound is not None and math.isfinite(bound) else 0.0
        cfg.seed = seed & (2**64 - 1)
        cfg.point_index_base = int(base)
        return cfg

    def label_gen(self, tx: torch.Tensor, M_T: int, M_I: int, *, precision: str = "fp32", sample_bound=None,
                  seed: int = 0, point_index_base: int = 0, draws: Optional[dict] = None,
                  check: bool = False, estimate_delta_t: float = 0.0) -> torch.Tensor:
        """dpi_label_gen on device tensors (asynchronous).  estimate_delta_t > 0 selects the TD estimators
        (data.py:529-575, :928-952).  check=True additionally synchronises and raises
        FloatingPointError if a label is inf / NaN (dpi_check_labels), e.g. a network outside fp16's range."""
        tx = _f32c(tx, self.device)
        n = tx.shape[0]
        assert tx.shape[1] == 1 + self.nx
        y = torch.empty_like(tx)
        cfg = self._cfg(M_T, M_I, precision, sample_bound, seed, point_index_base, estimate_delta_t)
        dr = None
        keep = []
        if draws is not None:
            dr = nat.dpi_draws()
            zT, u_s, zS = (_f32c(draws[k], self.device) for k in ("zT", "u_s", "zS"))
            assert zT.numel() == n * M_T * self.nx and zS.numel() == n * M_I * self.nx and u_s.numel() == n * M_I
            keep += [zT, u_s, zS]
            dr.zT, dr.u_s, dr.zS = zT.data_ptr(), u_s.data_ptr(), zS.data_ptr()
            if draws.get("idx") is not None:
                idx = draws["idx"].to(device=self.device, dtype=torch.int32).contiguous()
                assert idx.numel() == 0 or (int(idx.min()) >= 0 and int(idx.max()) < self.nx), \
                    "SDGD coordinates must lie in [0, nx)"
                keep.append(idx)
                dr.idx = idx.data_ptr()
        with torch.cuda.device(self.device):
            nat.check(self.lib.dpi_label_gen(self.h, _ptr(tx), n, C.byref(cfg), C.byref(dr) if dr is not None else None,
                                             _ptr(y), self._stream()), "dpi_label_gen")
        if keep:  # draws must outlive the asynchronous launch
            torch.cuda.current_stream(self.device).synchronize()
        if check:
            self.check_labels()
        return y

    def check_labels(self) -> None:
        """Synchronise and raise FloatingPointError if a label written since the last check was inf / NaN."""
        with torch.cuda.device(self.device):
            nat.check(self.lib.dpi_check_labels(self.h, self._stream()), "dpi_check_labels")

    def export_draws(self, n: int, M_T: int, M_I: int, *, seed: int = 0, point_index_base: int = 0,
                     want=("zT", "u_s", "zS", "idx")) -> dict:
        """The Philox draws a label_gen(seed, point_index_base) call consumes, in the reference's layout
        (zT [n*M_T,nx], u_s [n*M_I,1], zS [n*M_I,nx], idx [n*M_I,v]); usable as `draws=` or by the reference itself."""
        cfg = self._cfg(M_T, M_I, "fp32", None, seed, point_index_base)
        out = {}
        if "zT" in want:
            out["zT"] = torch.empty(n * M_T, self.nx, device=self.device, dtype=torch.float32)
        if "u_s" in want:
            out["u_s"] = torch.empty(n * M_I, 1, device=self.device, dtype=torch.float32)
        if "zS" in want:
            out["zS"] = torch.empty(n * M_I, self.nx, device=self.device, dtype=torch.float32)
        if "idx" in want and getattr(self, "sdgd_v", 0) > 0:
            out["idx"] = torch.empty(n * M_I, self.sdgd_v, device=self.device, dtype=torch.int32)
        with torch.cuda.device(self.device):
            nat.check(self.lib.dpi_export_draws(self.h, n, C.byref(cfg), _ptr(out.get("zT")), _ptr(out.get("u_s")),
                                                _ptr(out.get("zS")), _ptr(out.get("idx")), self._stream()),
                      "dpi_export_draws")
        out.setdefault("idx", None)
        return out

    def label_gen_host(self, n: int, M_T: int, M_I: int, *, tx_host: Optional[torch.Tensor] = None,
                       y_host: Optional[torch.Tensor] = None, tx_out_host: Optional[torch.Tensor] = None,
                       precision: str = "fp32", sample_bound=None, seed: int = 0, point_index_base: int = 0):
        """Host-buffer entry point (H2D + kernels + D2H + sync inside the call)."""
        ld = 1 + self.nx
        if y_host is None:
            y_host = torch.empty(n, ld, dtype=torch.float32, pin_memory=True)
        if tx_host is None and tx_out_host is None:
            tx_out_host = torch.empty(n, ld, dtype=torch.float32, pin_memory=True)
        cfg = self._cfg(M_T, M_I, precision, sample_bound, seed, point_index_base)
        with torch.cuda.device(self.device):
            nat.check(self.lib.dpi_label_gen_host(
                self.h, _ptr(tx_host), n, C.byref(cfg), _ptr(tx_out_host), _ptr(y_host), self._stream()),
                "dpi_label_gen_host")
        return (tx_host if tx_host is not None else tx_out_host), y_host

    def net_eval(self, tx: torch.Tensor, want_grad: bool = False):
        tx = _f32c(tx, self.device)
        n = tx.shape[0]
        u = torch.empty(n, device=self.device, dtype=torch.float32)
        du = torch.empty(n, self.nx, device=self.device, dtype=torch.float32) if want_grad else None
        with torch.cuda.device(self.device):
            nat.check(self.lib.dpi_net_eval(self.h, _ptr(tx), n, _ptr(u), _ptr(du), self._stream()), "dpi_net_eval")
        return u, du

    def net_hessian_diag(self, tx: torch.Tensor) -> torch.Tensor:
        """diag of the Hessian of u_k in x at the rows of tx: [n, nx] (dpi_net_hessian_diag)."""
        tx = _f32c(tx, self.device)
        d2 = torch.empty(tx.shape[0], self.nx, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            nat.check(self.lib.dpi_net_hessian_diag(self.h, _ptr(tx), tx.shape[0], _ptr(d2), self._stream()),
                      "dpi_net_hessian_diag")
        return d2

    def net_hessian(self, tx: torch.Tensor) -> torch.Tensor:
        """The full Hessian of u_k in x at the rows of tx: [n, nx, nx] (dpi_net_hessian)."""
        tx = _f32c(tx, self.device)
        H = torch.empty(tx.shape[0], self.nx, self.nx, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            nat.check(self.lib.dpi_net_hessian(self.h, _ptr(tx), tx.shape[0], _ptr(H), self._stream()), "dpi_net_hessian")
        return H

    def launch_count(self) -> int:
        return int(self.lib.dpi_launch_count(self.h))

    def selftest_umma(self, A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
        """D = fp16(A) @ fp16(B).T on the tensor cores (128x128x128), for the hardware self-test."""
        A, B = _f32c(A, self.device), _f32c(B, self.device)
        assert A.shape == (128, 128) and B.shape == (128, 128)
        D = torch.empty(128, 128, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            nat.check(self.lib.dpi_selftest_umma(self.h, _ptr(A), _ptr(B), _ptr(D), self._stream()),
                      "dpi_selftest_umma")
        return D

    STAGES = {"terminal": 0, "baseline": 1, "integral": 2}

    def set_stage_timing(self, enable: bool = True) -> None:
        nat.check(self.lib.dpi_set_stage_timing(self.h, int(bool(enable))), "dpi_set_stage_timing")

    def stage_ms(self, stage: str) -> float:
        """Device time of one stage of the most recent label_gen call (CUDA events on the launching stream)."""
        ms = C.c_float()
        nat.check(self.lib.dpi_stage_ms(self.h, self.STAGES[stage], C.byref(ms)), "dpi_stage_ms")
        return float(ms.value)


# ------------------------------------------------------------------------------------------------
# torch custom op: thin shim so the label generator composes with torch code (streams, graphs capture-free).
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("picard_b200::label_gen", mutates_args=(), device_types="cuda")
def label_gen_op(tx: torch.Tensor, engine_id: int, M_T: int, M_I: int, precision: str, sample_bound: float, seed: int,
                 point_index_base: int, zT: Optional[torch.Tensor] = None, u_s: Optional[torch.Tensor] = None,
                 zS: Optional[torch.Tensor] = None, idx: Optional[torch.Tensor] = None,
                 estimate_delta_t: float = 0.0) -> torch.Tensor:
    """`torch.ops.picard_b200.label_gen`: what `OnlineDataGenerator.generate_with_gradients` calls.  `engine_id` names
    the LabelEngine (handle with the packed equation and frozen u_k); the body is the C-ABI call, nothing else."""
    eng = _ENGINES[engine_id]
    draws = None if zT is None else {"zT": zT, "u_s": u_s, "zS": zS, "idx": idx}
    return eng.label_gen(tx, M_T, M_I, precision=precision, sample_bound=sample_bound, seed=seed,
                         point_index_base=point_index_base, draws=draws, estimate_delta_t=estimate_delta_t)


@label_gen_op.register_fake
def _(tx, engine_id, M_T, M_I, precision, sample_bound, seed, point_index_base, zT=None, u_s=None, zS=None, idx=None,
      estimate_delta_t=0.0):
    return torch.empty_like(tx, dtype=torch.float32)
