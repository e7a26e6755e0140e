"""Host-side mirror of the reference's regression step for one Picard iteration.

`RegressionTrainer` plays the role of `PicardSolutionGradientWrapper.training_step` + `loss.backward()` +
`torch.optim.Adam.step()` (picard/solution_jac.py:172-216, picard/solution.py:93-97) for a scalar-output MLP with
`FixedLossScaler`: same loss, same optimiser, fp32 master weights held in the solution's own `nn.Linear` tensors and
updated in place by the library (no torch autograd on this path).

Data parallel (one process per GPU of a node, torch.distributed initialised): `enable_data_parallel()` exchanges CUDA
IPC handles of the ranks' gradient blocks once; from then on `fit` stays the native, graph-replayed inner loop and
the gradient all-reduce happens INSIDE the optimiser kernel over NVLink peer memory (`dpi_train_dp_*`, dpi_train.cu):
every rank sums the peers' gradients in rank order, so the replicas of `u_k` stay bitwise identical without a
broadcast.  `step()` (one mini-batch from Python) keeps the plain NCCL all-reduce between `dpi_train_grad` and
`dpi_train_adam`.

`AutogradRegressionTrainer` is the LIBRARY path for what the native step does not cover -- PISGradNet (the HJB
config: `jacrev` through the network inside the loss, solution_jac.py:167-213) and the `SimpleLossScaler` /
`DimensionLossScaler` weightings (solution_jac.py:40-68): same loss and optimiser, evaluated by torch autograd
(double backward) and `torch.optim.Adam`.  It exists so that every BASELINE config can be SOLVED end to end on the
label kernels; it is not a kernel of this repository and is not counted as one.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _native as nat
from .engine import LabelEngine
from .solution import _linears


class RegressionTrainer:
    def __init__(self, solution, *, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, loss_beta: float = 0.0,
                 fixed_weight: float = 1.0, max_batch: int = 4096, device="cuda", process_group=None,
                 data_parallel: Optional[bool] = None):
        model = getattr(solution, "model", solution)
        if not isinstance(model, torch.nn.Sequential):
            raise NotImplementedError("the B200 regression step is built for plain MLPs (Linear + ELU)")
        self.linears = [m for m in model if isinstance(m, torch.nn.Linear)]
        _linears(model, type(solution).__name__)  # raises for anything but Linear / ELU(alpha=1)
        for m in self.linears:
            if m.weight.dtype != torch.float32 or not m.weight.is_cuda or not m.weight.is_contiguous():
                raise ValueError("RegressionTrainer updates the solution's tensors in place: they must be contiguous "
                                 "float32 CUDA tensors (solution.float().cuda())")
        if self.linears[-1].out_features != 1:
            raise NotImplementedError("only NETWORK.TYPE=Value (scalar output) is built")
        self.engine = LabelEngine(device)
        self.device = self.engine.device
        self.cfg = nat.dpi_train_config(lr, betas[0], betas[1], eps, loss_beta, fixed_weight)
        net = nat.dpi_network()
        net.kind = nat.DPI_NET_MLP
        net.n_main = len(self.linears)
        for i, m in enumerate(self.linears):
            net.main[i].W, net.main[i].b = m.weight.data_ptr(), m.bias.data_ptr()
            net.main[i].n_out, net.main[i].n_in = m.out_features, m.in_features
        with torch.cuda.device(self.device):
            nat.check(self.engine.lib.dpi_train_init(self.engine.h, C.byref(net), max_batch, self.engine._stream()),
                      "dpi_train_init")
        self.n_params = int(self.engine.lib.dpi_train_param_count(self.engine.h))
        self.grad = torch.empty(self.n_params, device=self.device, dtype=torch.float32)
        self.losses = torch.zeros(3, device=self.device, dtype=torch.float32)
        self.step_count = 0
        self.pg = process_group
        self.dp_world = 1
        self._dp_wanted = data_parallel  # None: data parallel iff torch.distributed runs with more than one rank
        if self._distributed():
            self.enable_data_parallel()

    def _distributed(self) -> bool:
        if self._dp_wanted is False:
            return False
        d = torch.distributed
        return self.pg is not None or (d.is_available() and d.is_initialized() and d.get_world_size() > 1)

    def enable_data_parallel(self) -> None:
        """All-gather the ranks' IPC handles and map the peers' gradient blocks (ranks of one node)."""
        d = torch.distributed
        rank, world = d.get_rank(self.pg), d.get_world_size(self.pg)
        e = self.engine
        buf = (C.c_ubyte * nat.DPI_DP_HANDLE_BYTES)()
        with torch.cuda.device(self.device):
            nat.check(e.lib.dpi_train_dp_export(e.h, buf), "dpi_train_dp_export")
        mine = torch.tensor(list(buf), dtype=torch.uint8, device=self.device if d.get_backend(self.pg) == "nccl" else "cpu")
        allh = [torch.empty_like(mine) for _ in range(world)]
        d.all_gather(allh, mine, group=self.pg)
        raw = bytes(torch.cat(allh).cpu().tolist())
        with torch.cuda.device(self.device):
            nat.check(e.lib.dpi_train_dp_connect(e.h, rank, world, C.c_char_p(raw)), "dpi_train_dp_connect")
        self.dp_world = world

    def close(self):
        self.engine.close()

    def reset_optimizer(self):
        """Fresh Adam state, weights kept: what the reference gets from a new Trainer per Picard iteration."""
        e = self.engine
        with torch.cuda.device(self.device):
            nat.check(e.lib.dpi_train_reset(e.h, e._stream()), "dpi_train_reset")
        self.step_count = 0

    def backward(self, tx: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """Loss terms and the flat parameter gradient of one mini-batch; returns losses [total, value, gradient]."""
        tx = tx.detach().to(self.device, torch.float32).contiguous()
        y = y.detach().to(self.device, torch.float32).contiguous()
        assert tx.shape == y.shape and tx.shape[1] == self.linears[0].in_features
        e = self.engine
        with torch.cuda.device(self.device):
            nat.check(e.lib.dpi_train_grad(e.h, C.c_void_p(tx.data_ptr()), C.c_void_p(y.data_ptr()), tx.shape[0],
                                           C.byref(self.cfg), C.c_void_p(self.grad.data_ptr()),
                                           C.c_void_p(self.losses.data_ptr()), e._stream()), "dpi_train_grad")
        self._keep = (tx, y)
        return self.losses

    def step(self, tx: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """One optimiser step on a mini-batch (solution_jac.py:172-216 + Adam)."""
        losses = self.backward(tx, y)
        if self.pg is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                                   and torch.distributed.get_world_size() > 1):
            torch.distributed.all_reduce(self.grad, op=torch.distributed.ReduceOp.AVG, group=self.pg)
        self.step_count += 1
        e = self.engine
        with torch.cuda.device(self.device):
            nat.check(e.lib.dpi_train_adam(e.h, C.c_void_p(self.grad.data_ptr()), C.byref(self.cfg), self.step_count,
                                           e._stream()), "dpi_train_adam")
        return losses

    def fit(self, tx: torch.Tensor, y: torch.Tensor, batch_size: int, n_epochs: int) -> Optional[torch.Tensor]:
        """The reference's inner loop over one iteration's label set: unshuffled mini-batches (config.py:90), n_epochs passes."""
        n = tx.shape[0]
        if not self._distributed() or self.dp_world > 1:
            # the whole inner loop runs natively (no host round trip between the steps); with peers connected the
            # captured step ends in the fused NVLink reduce + Adam kernel and batch_size is this rank's share
            tx = tx.detach().to(self.device, torch.float32).contiguous()
            y = y.detach().to(self.device, torch.float32).contiguous()
            e = self.engine
            with torch.cuda.device(self.device):
                nat.check(e.lib.dpi_train_fit(e.h, C.c_void_p(tx.data_ptr()), C.c_void_p(y.data_ptr()), n, batch_size,
                                              n_epochs, C.byref(self.cfg), self.step_count + 1,
                                              C.c_void_p(self.losses.data_ptr()), e._stream()), "dpi_train_fit")
            self.step_count += n_epochs * ((n + batch_size - 1) // batch_size)
            self._keep = (tx, y)
            return self.losses
        last = None
        for _ in range(n_epochs):
            for i in range(0, n, batch_size):
                last = self.step(tx[i:i + batch_size], y[i:i + batch_size])
        return last



class AutogradRegressionTrainer:
    """Same interface as `RegressionTrainer` (reset_optimizer / step / fit / close, losses = [total, value, gradient]),
    any scalar-output torch module, loss of `PicardSolutionGradientWrapper.training_step` (solution_jac.py:167-213) with
    `scaler` in {"FixedLossScaler", "SimpleLossScaler", "DimensionLossScaler"} (solution_jac.py:40-78)."""

    def __init__(self, solution, *, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, loss_beta: float = 0.0,
                 scaler: Optional[str] = "FixedLossScaler", fixed_weight: float = 1.0, device="cuda", process_group=None):
        if scaler not in (None, "FixedLossScaler", "SimpleLossScaler", "DimensionLossScaler"):
            raise NotImplementedError(f"loss scaler {scaler}")
        self.solution, self.device = solution, torch.device(device)
        eq = getattr(solution, "equation", None)
        if eq is not None and hasattr(eq, "to"):
            eq.to(self.device)  # PISGradNet's residual g0 = equation.g reads the equation's tensors on every forward
        self.lr, self.betas, self.eps = lr, betas, eps
        self.beta, self.scaler, self.fixed_weight = float(loss_beta), scaler or "FixedLossScaler", float(fixed_weight)
        self.pg = process_group
        self.losses = torch.zeros(3, device=self.device, dtype=torch.float32)
        self.step_count = 0
        self.reset_optimizer()

    def _world(self) -> int:
        d = torch.distributed
        return d.get_world_size(self.pg) if d.is_available() and d.is_initialized() else 1

    def reset_optimizer(self):
        """Fresh Adam state, weights kept.  Once a step graph has been captured the state is cleared IN PLACE (the graph
        holds the state tensors), otherwise a new optimiser is built."""
        if getattr(self, "_graph", None) is not None:
            for st in self.opt.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
        else:
            params = [p for p in self.solution.parameters() if p.requires_grad]
            cuda = self.device.type == "cuda"
            self.opt = torch.optim.Adam(params, lr=self.lr, betas=self.betas, eps=self.eps, capturable=cuda)
        self.step_count = 0

    def loss_terms(self, tx: torch.Tensor, y: torch.Tensor):
        """v_loss [1], gd_loss_multi_dim [nx] and the scaled total (solution_jac.py:167-213)."""
        tx = tx.detach().clone().requires_grad_(True)
        u = self.solution(tx).reshape(-1, 1)
        (du,) = torch.autograd.grad(u.sum(), tx, create_graph=True)  # rows are independent: the row-wise jacrev
        w = torch.exp(tx[:, 0:1].detach() * self.beta)
        v_loss = torch.mean(w * (u - y[:, 0:1]) ** 2, dim=0)
        gd = torch.mean(w * (du[:, 1:] - y[:, 1:]) ** 2, dim=0)
        if self.scaler == "FixedLossScaler":
            g = gd.sum(-1, keepdim=True)
            total = v_loss + self.fixed_weight * g
        elif self.scaler == "SimpleLossScaler":
            # as upstream (solution_jac.py:42-47): the summed gradient loss is formed under no_grad, so this scaler's
            # second term shifts the reported loss but contributes nothing to the parameter gradient
            g = gd.sum(-1, keepdim=True).detach()
            a = torch.clamp(v_loss.detach() / g, min=0.0, max=1e3)
            total = v_loss + a * g
        else:  # DimensionLossScaler
            a = torch.clamp((v_loss / gd).detach(), min=0.0, max=1e3)
            g = (a * gd).sum(-1, keepdim=True)
            total = v_loss + g
        return total, v_loss, g

    def step(self, tx: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        p0 = next(self.solution.parameters())
        tx, y = tx.to(self.device, p0.dtype), y.to(self.device, p0.dtype)
        self.opt.zero_grad(set_to_none=True)
        total, v, g = self.loss_terms(tx, y)
        total.sum().backward()
        if self._world() > 1:
            for p in self.solution.parameters():
                if p.grad is not None:
                    torch.distributed.all_reduce(p.grad, op=torch.distributed.ReduceOp.AVG, group=self.pg)
        self.opt.step()
        self.step_count += 1
        self.losses = torch.stack([total.detach().sum(), v.detach().sum(), g.detach().sum()]).float()
        return self.losses

    def _capture(self, tx: torch.Tensor, y: torch.Tensor) -> None:
        """One optimiser step (loss, double backward, Adam) as a CUDA graph on static batch buffers: the step is a few
        hundred small autograd kernels, launch-bound when run eagerly."""
        self._gtx, self._gy = tx.clone(), y.clone()
        self._graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self._graph):
            total, v, g = self.loss_terms(self._gtx, self._gy)
            total.sum().backward()
            self.opt.step()
            self._glosses = torch.stack([total.detach().sum(), v.detach().sum(), g.detach().sum()]).float()

    def fit(self, tx: torch.Tensor, y: torch.Tensor, batch_size: int, n_epochs: int) -> Optional[torch.Tensor]:
        last = None
        n = tx.shape[0]
        graphable = (self.device.type == "cuda" and self._world() == 1 and n % batch_size == 0
                     and os.environ.get("DPI_AUTOGRAD_NO_GRAPH") is None)
        p0 = next(self.solution.parameters())
        for _ in range(n_epochs):
            for i in range(0, n, batch_size):
                xb, yb = tx[i:i + batch_size], y[i:i + batch_size]
                g = getattr(self, "_graph", None)
                if graphable and g is not None and tuple(self._gtx.shape) == tuple(xb.shape):
                    self._gtx.copy_(xb)
                    self._gy.copy_(yb)
                    g.replay()
                    self.step_count += 1
                    last = self.losses = self._glosses
                    continue
                last = self.step(xb, yb)
                if graphable and g is None and self.step_count >= 3:  # three eager steps first (allocator / autotune warm-up)
                    torch.cuda.synchronize(self.device)
                    self._capture(xb.to(self.device, p0.dtype), yb.to(self.device, p0.dtype))
        return last

    def close(self):
        self._graph = None
