"""Evaluation callback on the label kernels.

The reference evaluates every EVAL.FREQ steps on L2_N_POINTS points `t = linspace(0, T, n)`, `x = equation.sample_x(t)`
(`EvalCallback.on_validation_epoch_start`, picard/utils.py:362-478), taking u by a forward pass, grad_x u by
`autograd.grad` and, with TEST_HESSIAN, the Hessian by d further backward passes (`compute_at_t`, picard/utils.py:117-228;
`get_hessian`, :579-591).  Here the three quantities come from the library through the C ABI:

    u, grad_x u      dpi_net_eval           value chain + one VJP (plain MLP, PISGradNet closed form, ZeroSolution)
    diag Hess_x u    dpi_net_hessian_diag   forward 2-jets along the coordinate directions (the SDGD kernel's baseline tile)

    Hess_x u         dpi_net_hessian        the same jets along the pairs e_i + e_j (polarisation): all d x d entries

and the metrics are the reference's, name for name (MSE, rRMSE, rMAE, MArE, and the *g / *h variants computed per
dimension / per Hessian entry and then averaged, picard/utils.py:405-474).  With TEST_HESSIAN the *h metrics compare
all d x d entries against `equation.u_hessian` like upstream (picard/utils.py:198-200,461-470); an equation that only
offers `hessian_diagonal` is compared on the diagonal.

Exact solutions and the point sampler are the host mirror's torch code (equations.py): evaluation is not the hot path.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .engine import LabelEngine


def compute_at_t(engine: LabelEngine, t: torch.Tensor, equation, maximal_batch_size: Optional[int] = None,
                 eval_gradient: bool = False, eval_hessian: bool = False, generator: Optional[torch.Generator] = None):
    """picard/utils.py:117-228 with the network evaluated by the library.  t [n,1] on the engine's device.
    Returns numpy arrays (u, u_exact, u_x, u_x_exact, u_xx_diag, u_xx_diag_exact); entries not asked for are None."""
    dev = engine.device
    eq = equation.to(dev) if hasattr(equation, "to") else equation
    t = t.to(dev).reshape(-1, 1)
    outs = {k: [] for k in ("u", "ue", "g", "ge", "h", "he")}
    step = maximal_batch_size or t.shape[0]
    for i in range(0, t.shape[0], step):
        tb = t[i:i + step]
        z = torch.randn(tb.shape[0], eq.nx, device=dev, dtype=tb.dtype, generator=generator)
        xb = eq.sample_x0(tb.shape[0]).to(dev, tb.dtype) + torch.sqrt(tb) * eq.alpha_sqrt * z  # equations.py:118-119
        tx = torch.cat([tb, xb], -1)
        u, du = engine.net_eval(tx, want_grad=eval_gradient)
        outs["u"].append(u.view(-1, 1).double())
        outs["ue"].append(eq.exact_solution(tb.double(), xb.double()))
        if eval_gradient:
            outs["g"].append(du.double())
            outs["ge"].append(eq.u_x(tb.double(), xb.double()))
        if eval_hessian:
            if hasattr(eq, "u_hessian"):  # the reference's comparison: every entry of the Hessian
                outs["h"].append(engine.net_hessian(tx).double())
                outs["he"].append(eq.u_hessian(tb.double(), xb.double()))
            else:
                outs["h"].append(engine.net_hessian_diag(tx).double())
                outs["he"].append(eq.hessian_diagonal(tb.double(), xb.double()))
    cat = lambda k: torch.cat(outs[k]).cpu().numpy() if outs[k] else None
    return cat("u"), cat("ue"), cat("g"), cat("ge"), cat("h"), cat("he")


def error_metrics(u, u_exact, u_x=None, u_x_exact=None, u_xx=None, u_xx_exact=None) -> Dict[str, float]:
    """The metric block of EvalCallback (picard/utils.py:405-474): sums, not means, in the relative norms."""
    m: Dict[str, float] = {}

    def block(val, exact, suffix, axis):
        e = np.abs(val - exact)
        norm = lambda a, p: (np.abs(a) ** p).sum(axis) ** (1.0 / p)
        m["MSE" + suffix] = float(np.sqrt((e ** 2).mean(axis)).mean())
        m["rRMSE" + suffix] = float((norm(e, 2) / norm(exact, 2)).mean())
        m["rMAE" + suffix] = float((norm(e, 1) / norm(exact, 1)).mean())
        m["MArE" + suffix] = float((e / np.abs(exact)).mean())  # inf when an exact value is 0 (e.g. t = 0, x = 0), as upstream

    with np.errstate(divide="ignore", invalid="ignore"):
        block(u, u_exact, "", None)  # the value metrics pool all points
        if u_x is not None:
            block(u_x, u_x_exact, "g", 0)  # the derivative metrics are per dimension, then averaged
        if u_xx is not None:
            block(u_xx, u_xx_exact, "h", 0)
    return m


class EvalCallback:
    """`EvalCallback(eval_cfg, equation, writer, iter)` (picard/utils.py:329-360).  `eval_cfg` carries L2_N_POINTS,
    BATCH_SIZE, TEST_GRAD, TEST_HESSIAN (attributes or keys); `writer`, if given, needs `add_scalar(tag, value, step)`
    (tensorboardX's protocol).  Call it with the engine that holds the network to evaluate."""

    def __init__(self, eval_cfg, equation, writer=None, iter: int = 0, seed: Optional[int] = None):
        get = (lambda k, d=None: eval_cfg.get(k, d)) if isinstance(eval_cfg, dict) else (lambda k, d=None: getattr(eval_cfg, k, d))
        self.n_points = int(get("L2_N_POINTS", 10_000))
        self.batch_size = get("BATCH_SIZE")
        self.eval_gradient = bool(get("TEST_GRAD", False))
        self.eval_hessian = bool(get("TEST_HESSIAN", False))
        self.equation, self.writer, self.iter, self.seed = equation, writer, iter, seed
        if self.eval_hessian and not (hasattr(equation, "u_hessian") or hasattr(equation, "hessian_diagonal")):
            raise NotImplementedError(f"{type(equation).__name__} has no exact Hessian diagonal to compare against")

    def __call__(self, engine: LabelEngine) -> Dict[str, float]:
        dev = engine.device
        t = torch.linspace(0.0, float(self.equation.T), self.n_points, device=dev).reshape(-1, 1)  # utils.py:366-367
        gen = None
        if self.seed is not None:
            gen = torch.Generator(device=dev).manual_seed(self.seed)
        vals = compute_at_t(engine, t, self.equation, self.batch_size, self.eval_gradient, self.eval_hessian, gen)
        metrics = error_metrics(*vals)
        if self.writer is not None:
            for k, v in metrics.items():
                self.writer.add_scalar(k, v, self.iter)
        return metrics

    on_validation_epoch_start = __call__
