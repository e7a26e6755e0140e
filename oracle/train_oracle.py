"""CPU oracle of the regression step that follows label generation in a Picard iteration.

TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product path.

Restates `PicardSolutionGradientWrapper.training_step` (picard/solution_jac.py:172-216) for a scalar-output MLP with
`FixedLossScaler` (picard/solution_jac.py:77-89), `loss_fn = torch.square` (picard/solution.py:66-68) and the time
weight `exp(beta t)` (picard/solution_jac.py:189), followed by one `torch.optim.Adam` step
(`PicardBaseSolution.configure_optimizers`, picard/solution.py:95-97).  Two statements of the loss gradient:

* `loss_and_grads_autograd`: torch autograd with create_graph, i.e. what the reference executes (vmap(jacrev) + backward);
* `loss_and_grads_closed`: the forward-tangent form the CUDA pipeline implements.  With r = dL/d(grad_x u) held fixed,
  d/dtheta <r, grad_x u> = d/dtheta JVP_r(u), so the parameter gradient is ONE reverse sweep through the value + tangent
  network (the same pair the label kernels evaluate), with ELU'' entering through the tangent branch.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

Tensor = torch.Tensor


def _elu_d1(z):
    return torch.where(z > 0, torch.ones_like(z), torch.exp(z))


def _elu_d2(z):
    return torch.where(z > 0, torch.zeros_like(z), torch.exp(z))


def forward(params: Sequence[Tensor], tx: Tensor) -> Tensor:
    h = tx
    L = len(params) // 2
    for i in range(L):
        h = h @ params[2 * i].t() + params[2 * i + 1]
        if i < L - 1:
            h = torch.nn.functional.elu(h)
    return h


def loss_terms(u: Tensor, ux: Tensor, tx: Tensor, y: Tensor, beta: float, fixed_weight: float):
    """solution_jac.py:185-211: v_loss = mean_b w (u - y_u)^2 ; g_loss = sum_j mean_b w (u_xj - y_j)^2."""
    w = torch.exp(tx[:, :1] * beta)
    v_loss = torch.mean(w * (u - y[:, :1]) ** 2)
    g_loss = torch.mean(w * (ux - y[:, 1:]) ** 2, dim=0).sum()
    return v_loss + fixed_weight * g_loss, v_loss, g_loss


def loss_and_grads_autograd(params: Sequence[Tensor], tx: Tensor, y: Tensor, beta: float, fixed_weight: float):
    ps = [p.detach().clone().requires_grad_() for p in params]
    txg = tx.detach().clone().requires_grad_()
    u = forward(ps, txg)
    (utx,) = torch.autograd.grad(u.sum(), txg, create_graph=True)
    loss, v_loss, g_loss = loss_terms(u, utx[:, 1:], tx, y, beta, fixed_weight)
    grads = torch.autograd.grad(loss, ps)
    return loss.detach(), v_loss.detach(), g_loss.detach(), [g.detach() for g in grads]


def loss_and_grads_closed(params: Sequence[Tensor], tx: Tensor, y: Tensor, beta: float, fixed_weight: float):
    L = len(params) // 2
    W = [params[2 * i] for i in range(L)]
    b = [params[2 * i + 1] for i in range(L)]
    B = tx.shape[0]
    # value forward
    h, zs, hs, es = tx, [], [tx], []
    for l in range(L - 1):
        z = h @ W[l].t() + b[l]
        zs.append(z)
        es.append(_elu_d1(z))
        h = torch.nn.functional.elu(z)
        hs.append(h)
    u = h @ W[L - 1].t() + b[L - 1]
    # input gradient by one reverse sweep
    c = es[L - 2] * W[L - 1]  # [B,H]
    for l in range(L - 2, 0, -1):
        c = (c @ W[l]) * es[l - 1]
    g = c @ W[0]  # [B,1+d]
    w = torch.exp(tx[:, :1] * beta)
    loss, v_loss, g_loss = loss_terms(u, g[:, 1:], tx, y, beta, fixed_weight)
    a = 2.0 * w * (u - y[:, :1]) / B  # dL/du
    r = torch.zeros_like(tx)
    r[:, 1:] = 2.0 * fixed_weight * w * (g[:, 1:] - y[:, 1:]) / B  # dL/d(grad_x u), held fixed
    # tangent forward along r
    hd, zds, hds = r, [], [r]
    for l in range(L - 1):
        zd = hd @ W[l].t()
        zds.append(zd)
        hd = es[l] * zd
        hds.append(hd)
    # reverse sweep through (value, tangent)
    gW: List[Tensor] = [None] * L
    gb: List[Tensor] = [None] * L
    gW[L - 1] = a.t() @ hs[L - 1] + hds[L - 1].sum(0, keepdim=True)
    gb[L - 1] = a.sum(0)
    hbar = a * W[L - 1]          # [B,H]
    hdbar = W[L - 1].expand(B, -1)  # d(sum_b udot_b)/d hdot
    for l in range(L - 2, -1, -1):
        zbar = hbar * es[l] + hdbar * _elu_d2(zs[l]) * zds[l]
        zdbar = hdbar * es[l]
        gW[l] = zbar.t() @ hs[l] + zdbar.t() @ hds[l]
        gb[l] = zbar.sum(0)
        if l > 0:
            hbar = zbar @ W[l]
            hdbar = zdbar @ W[l]
    grads = []
    for l in range(L):
        grads += [gW[l], gb[l]]
    return loss, v_loss, g_loss, grads


def adam_step(params: List[Tensor], grads: List[Tensor], m: List[Tensor], v: List[Tensor], step: int, lr: float = 1e-3,
              beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.Adam (no weight decay, no amsgrad), in place; `step` is 1-based."""
    bc1, bc2 = 1.0 - beta1 ** step, 1.0 - beta2 ** step
    for p, g, mi, vi in zip(params, grads, m, v):
        mi.mul_(beta1).add_(g, alpha=1.0 - beta1)
        vi.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
        p.addcdiv_(mi, (vi / bc2).sqrt_().add_(eps), value=-lr / bc1)
