"""CPU restatement (torch, fp64 by default) of the reference's DPI label generator.

TEST INFRASTRUCTURE ONLY — this is the parity oracle and the `cpu_baseline` / `--impl reference`
arm of `bench.py`.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may import it;
the product path (`deeppicarditeration_b200/`) never does and has no CPU fallback.

Parity status: PINNED.  The reference ships no tests or golden vectors, so this
restatement is pinned against outputs of the *unmodified reference itself*, run in the build container
through `oracle/ref_shim.py` on replayed random draws; the resulting fixtures are committed under
`tests/golden/*.npz` together with the generating script `oracle/make_golden.py`, and
`tests/test_oracle_golden.py` checks every fixture (rel-L2 <= 1e-12 in fp64).

Every function cites the reference lines it restates (paths relative to /root/reference).  Two
evaluation modes of the integrand exist on purpose:

* ``mode="autograd"`` mirrors the reference op for op (forward + reverse-mode `autograd.grad`, and the
  v-fold double-backward loop for SDGD, data.py:1245-1251,1273-1292).  This is the arm timed as the CPU
  baseline, because it is what the reference executes.
* ``mode="closed"`` is the forward-mode / closed-form statement the CUDA kernels implement (one JVP
  along (0,1,...,1) for Burgers, a 2-jet per sampled coordinate for SDGD, one VJP for PISGradNet;
  SURVEY.md Appendix A).  Both modes are pinned by the same fixtures.

Row order everywhere is ``row = point * M + sample`` (repeat_interleave, data.py:336-348).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# Equations (equations.py)
# ----------------------------------------------------------------------------------------------
class Cha:
    """Burgers-type equation `Cha` (equations.py:266-338)."""

    name = "Cha"
    has_hessian_term = False

    def __init__(self, nx: int, alpha: float = 1.0, k: float = 1.0, T: float = 1.0):
        self.nx, self.alpha, self.T = int(nx), float(alpha), float(T)
        self.alpha_sqrt = math.sqrt(self.alpha)
        # equations.py:283-289
        self.k = float(k) / math.sqrt(self.nx)
        self.k_alpha_d = self.k * self.alpha * self.nx
        self.c = (2.0 + self.k * self.k_alpha_d) / (2.0 * self.k_alpha_d)

    def g(self, x: Tensor) -> Tensor:  # equations.py:304-305
        return torch.sigmoid(self.T + self.k * x.sum(-1, keepdim=True))

    def ff(self, t, x, y, w) -> Tensor:  # equations.py:199-200 + 297-302
        z = self.alpha_sqrt * w
        return self.alpha_sqrt * (self.k * y - self.c) * z.sum(-1, keepdim=True)

    def ff_from_sum(self, y: Tensor, dsum: Tensor) -> Tensor:
        """Same as `ff` but fed D = sum_j du/dx_j directly (what the JVP path produces)."""
        return self.alpha_sqrt * (self.k * y - self.c) * (self.alpha_sqrt * dsum)

    def exact_solution(self, t, x):  # equations.py:318-319
        return torch.sigmoid(t + self.k * x.sum(-1, keepdim=True))

    def u_x(self, t, x):  # equations.py:325-327
        u = self.exact_solution(t, x)
        return torch.ones_like(x) * (self.k * u * (1 - u))

    def x0_std(self) -> float:  # equations.py:334-335 (x0 == 0)
        return 0.0


class GBM:
    """`GBMEquationComplexExact` (equations.py:388-486); w [2,1+nx], v [2,1] come from the shipped .pt files."""

    name = "GBM"
    has_hessian_term = True

    def __init__(self, nx: int, w: Tensor, v: Tensor, alpha: float = 1.0, T: float = 1.0):
        self.nx, self.alpha, self.T = int(nx), float(alpha), float(T)
        self.alpha_sqrt = math.sqrt(self.alpha)
        self.w, self.v = w, v
        self.d = float(nx)

    def _arg(self, t, x):  # tx @ w.T  (equations.py:421-423)
        if not torch.is_tensor(t):
            t = torch.full((x.shape[0], 1), float(t), dtype=x.dtype)
        else:
            t = t * torch.ones(x.shape[0], 1, dtype=x.dtype)
        return torch.cat([t, x], dim=-1) @ self.w.t().to(x.dtype)

    def exact_solution(self, t, x):  # equations.py:421-424
        return torch.sin(self._arg(t, x)) @ self.v.to(x.dtype)

    def g(self, x):  # equations.py:415-416
        return self.exact_solution(self.T, x)

    def u_t(self, t, x):  # equations.py:426-428
        return torch.cos(self._arg(t, x)) @ (self.v * self.w[:, 0:1]).to(x.dtype)

    def u_x(self, t, x):  # equations.py:430-432
        return torch.cos(self._arg(t, x)) @ (self.v * self.w[:, 1:]).to(x.dtype)

    def laplacian(self, t, x):  # equations.py:452-455
        return -torch.sin(self._arg(t, x)) @ (self.v * (self.w[:, 1:] ** 2).sum(-1, keepdim=True)).to(x.dtype)

    def hess_diag(self, t, x):  # diagonal of equations.py:445-450
        # d2/dx_j^2 u* = -sum_k v_k w_kj^2 sin(a_k)
        return -torch.sin(self._arg(t, x)) @ (self.v * self.w[:, 1:] ** 2).to(x.dtype)

    def source(self, t, x):
        """The u_k-independent part of ffi (equations.py:457-466)."""
        return (
            -self.u_t(t, x)
            - 0.5 * self.laplacian(t, x)
            - 0.25 * self.hess_diag(t, x).abs().sum(-1, keepdim=True)
        )

    def ffi(self, t, x, u, u_ii):  # equations.py:457-466
        lap = self.d * u_ii.mean(-1, keepdim=True)
        non = self.d * u_ii.abs().mean(-1, keepdim=True)
        return 0.5 * (1.0 - self.alpha) * lap + 0.25 * non + self.source(t, x)

    def x0_std(self) -> float:  # equations.py:485-486
        return 0.0


class OU:
    """`OUProcessEquation` (equations.py:599-714) with diagonal GMM terminal (utils.py:792-880).

    mean [K,nx], diagvar [K,nx] (the diagonal of the reference's [K,nx,nx] `var`), pi [K].
    """

    name = "OU"
    has_hessian_term = False

    def __init__(self, nx, mean: Tensor, diagvar: Tensor, pi: Tensor, T=1.0, alpha=1.0, theta=1.0, mu=0.0,
                 alpha_scale=4.0):
        self.nx, self.T, self.alpha = int(nx), float(T), float(alpha)
        self.alpha_sqrt = math.sqrt(self.alpha)
        self.theta, self.mu = float(theta), float(mu)
        self.mean, self.diagvar, self.pi = mean, diagvar, pi
        self.alpha_init = alpha_scale * self.alpha  # equations.py:556
        self.d = float(nx)

    # --- GMM (utils.py:852-915) -----------------------------------------------------------
    @staticmethod
    def _gmm_log_terms(x, mean, diagvar, pi):
        diff = x.unsqueeze(1) - mean.unsqueeze(0)  # [B,K,n]
        exp_term = -0.5 * torch.einsum("bkn,kn->bk", diff ** 2, 1.0 / diagvar)
        norm = -0.5 * (mean.shape[-1] * math.log(2.0 * math.pi) + torch.log(torch.prod(diagvar, dim=1)))
        return torch.log(pi) + norm.unsqueeze(0) + exp_term, diff

    def gmm_log_prob(self, x, mean=None, diagvar=None):
        lp, _ = self._gmm_log_terms(x, self.mean if mean is None else mean,
                                    self.diagvar if diagvar is None else diagvar, self.pi)
        return torch.logsumexp(lp, dim=1).reshape(-1, 1)

    def gmm_grad_log_prob(self, x):
        lp, diff = self._gmm_log_terms(x, self.mean, self.diagvar, self.pi)
        resp = (lp - torch.logsumexp(lp, dim=1, keepdim=True)).exp()  # [B,K]
        return -(diff / self.diagvar.unsqueeze(0) * resp.unsqueeze(-1)).sum(1)

    def g(self, x):  # equations.py:592-593
        return -self.gmm_log_prob(x)

    def g_x(self, x):  # equations.py:595-596
        return -self.gmm_grad_log_prob(x)

    def ff(self, t, x, y, z):  # equations.py:657-666
        F = self.theta * (self.mu - x)
        return (-(F * z).sum(-1, keepdim=True) - self.alpha / 2 * (z ** 2).sum(-1, keepdim=True)
                - self.d * self.theta * torch.ones_like(y))

    def exact_solution(self, t, x):  # equations.py:625-655, row-wise GMM pushed through the OU flow
        out = []
        for n in range(t.shape[0]):
            lam = self.T - t[n]
            e1, e2 = torch.exp(-self.theta * lam), torch.exp(-2 * self.theta * lam)
            mean_t = self.mu + (self.mean - self.mu) * e1
            var_t = self.diagvar * e2 + (self.alpha / (2 * self.theta)) * (1 - e2)
            out.append(-self.gmm_log_prob(x[n:n + 1], mean_t, var_t))
        return torch.cat(out, 0)

    def x0_std(self) -> float:  # equations.py:710-711, utils.py:785-789
        return math.sqrt(self.alpha_init)


# ----------------------------------------------------------------------------------------------
# Networks u_k (solution.py)
# ----------------------------------------------------------------------------------------------
def _elu(z):
    return torch.where(z > 0, z, torch.expm1(z))


def _elu_d1(z):
    return torch.where(z > 0, torch.ones_like(z), torch.exp(z))


def _elu_d2(z):
    return torch.where(z > 0, torch.zeros_like(z), torch.exp(z))


@dataclass
class MLP:
    """`construct_mlp` (solution.py:123-135): Linear -> ELU -> ... -> Linear. weights[i] is [out,in]."""

    weights: List[Tensor]
    biases: List[Tensor]
    kind: str = "mlp"

    def forward(self, tx: Tensor) -> Tensor:
        h = tx
        L = len(self.weights)
        for i, (W, b) in enumerate(zip(self.weights, self.biases)):
            h = h @ W.t().to(h.dtype) + b.to(h.dtype)
            if i < L - 1:
                h = torch.nn.functional.elu(h)
        return h

    def jets(self, tx: Tensor, dirs: Optional[Tensor] = None, order: int = 1):
        """Forward-mode value + tangent (+ second-order) propagation.

        dirs: None -> the direction (0,1,...,1) (Burgers, d/dx summed); else int64 [rows, v] coordinate
        indices j (direction e_j in x, i.e. column 1+j of the first layer).
        Returns (u [rows,1], du [rows,v'], d2u [rows,v'] or None), v'=1 when dirs is None.
        """
        W1, b1 = self.weights[0].to(tx.dtype), self.biases[0].to(tx.dtype)
        z = tx @ W1.t() + b1  # [R,H]
        if dirs is None:
            zd = W1[:, 1:].sum(1).expand(z.shape[0], 1, -1)  # [R,1,H]
        else:
            zd = W1.t()[1 + dirs]  # [R,v,H]
        zdd = torch.zeros_like(zd) if order >= 2 else None
        L = len(self.weights)
        for i in range(1, L):
            W, b = self.weights[i].to(tx.dtype), self.biases[i].to(tx.dtype)
            e1 = _elu_d1(z).unsqueeze(1)
            if order >= 2:
                e2 = _elu_d2(z).unsqueeze(1)
                hdd = e2 * zd * zd + e1 * zdd
            hd = e1 * zd
            h = _elu(z)
            z = h @ W.t() + b
            zd = hd @ W.t()
            if order >= 2:
                zdd = hdd @ W.t()
        return z, zd[..., 0], (zdd[..., 0] if order >= 2 else None)


@dataclass
class ZeroNet:
    """`ZeroSolution` (solution.py:330-337): u == 0 and its autograd gradient is None -> zeros (data.py:1316)."""

    kind: str = "zero"

    def forward(self, tx):
        return torch.zeros(tx.shape[0], 1, dtype=tx.dtype)


@dataclass
class PISGradNet:
    """`PISGradNet` (solution.py:138-289).  `lin` lists are [(W[out,in], b[out])]."""

    phase: Tensor  # [1,64]   timestep_phase
    coeff: Tensor  # [1,64]   timestep_coeff = linspace(0.1,100,64)
    t_encoder: List[Tuple[Tensor, Tensor]]  # 2 linears, ELU between
    smooth_net: List[Tuple[Tensor, Tensor]]  # n_layers+2 linears, ELU between
    nn_module: List[Tuple[Tensor, Tensor]]  # len(hidden)+1 linears, ELU between
    T: float = 1.0
    kind: str = "pisgrad"

    @staticmethod
    def _chain(lins, h):
        for i, (W, b) in enumerate(lins):
            h = h @ W.t().to(h.dtype) + b.to(h.dtype)
            if i < len(lins) - 1:
                h = torch.nn.functional.elu(h)
        return h

    def _emb(self, lbd):  # solution.py:211-229
        arg = self.coeff.to(lbd.dtype) * lbd + self.phase.to(lbd.dtype)
        return torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)

    def smooth(self, lbd):  # solution.py:231-254
        return (self._chain(self.smooth_net, self._emb(lbd))[..., 0:1]
                - self._chain(self.smooth_net, self._emb(torch.zeros_like(lbd)))[..., 0:1])

    def forward(self, tx, g0=None):  # solution.py:256-289
        lbd, x = self.T - tx[..., 0:1], tx[..., 1:]
        S = self.smooth(lbd)
        t_emb = self._chain(self.t_encoder, self._emb(lbd))
        net_out = self._chain(self.nn_module, torch.cat([t_emb, x], dim=-1))
        sp = (net_out * x).sum(-1, keepdim=True)
        return S * sp + (1.0 - S) * g0(torch.exp(-0.5 * lbd) * x)

    def value_and_grad_closed(self, tx, eq: "OU"):
        """u and grad_x u without autograd:
        grad = S*(n + J_n^T x) + (1-S)*gamma*g_x(gamma x); J_n^T x is one VJP through nn_module."""
        lbd, x = self.T - tx[..., 0:1], tx[..., 1:]
        S = self.smooth(lbd)
        t_emb = self._chain(self.t_encoder, self._emb(lbd))
        h = torch.cat([t_emb, x], dim=-1)
        zs = []
        L = len(self.nn_module)
        for i, (W, b) in enumerate(self.nn_module):
            z = h @ W.t().to(h.dtype) + b.to(h.dtype)
            if i < L - 1:
                zs.append(z)
                h = _elu(z)
        n_out = z
        # VJP with cotangent x
        c = x @ self.nn_module[-1][0].to(x.dtype)  # [R,H]
        for i in range(L - 2, -1, -1):
            c = c * _elu_d1(zs[i])
            c = c @ self.nn_module[i][0].to(x.dtype)
        n_emb = t_emb.shape[-1]
        jtx = c[:, n_emb:]
        gamma = torch.exp(-0.5 * lbd)
        u = S * (n_out * x).sum(-1, keepdim=True) + (1.0 - S) * eq.g(gamma * x)
        grad = S * (n_out + jtx) + (1.0 - S) * gamma * eq.g_x(gamma * x)
        return u, grad


# ----------------------------------------------------------------------------------------------
# Draws, in the reference's consumption order (verified by tracing)
# ----------------------------------------------------------------------------------------------
@dataclass
class Draws:
    u_t: Tensor  # [n,1]          rand        data.py:166
    z_x0: Optional[Tensor]  # [n,nx]  randn (OU only)  utils.py:785-789
    z_x: Tensor  # [n,nx]         randn       equations.py:225 via sample_x
    zT: Tensor  # [n*M_T,nx]      randn       data.py:914
    u_s: Tensor  # [n*M_I,1]      rand        data.py:359
    zS: Tensor  # [n*M_I,nx]      randn       data.py:360
    idx: Optional[Tensor]  # [n*M_I,v] randint (SDGD only) data.py:501


def draw(eq, n: int, M_T: int, M_I: int, v: Optional[int] = None, generator: Optional[torch.Generator] = None,
         dtype=torch.float64, round_fp32: bool = False) -> Draws:
    g = generator
    rnd = (lambda t: t.float().to(dtype)) if round_fp32 else (lambda t: t)
    u_t = rnd(torch.rand(n, 1, generator=g, dtype=dtype))
    z_x0 = rnd(torch.randn(n, eq.nx, generator=g, dtype=dtype)) if eq.x0_std() > 0 else None
    z_x = rnd(torch.randn(n, eq.nx, generator=g, dtype=dtype))
    zT = rnd(torch.randn(n * M_T, eq.nx, generator=g, dtype=dtype))
    u_s = rnd(torch.rand(n * M_I, 1, generator=g, dtype=dtype))
    zS = rnd(torch.randn(n * M_I, eq.nx, generator=g, dtype=dtype))
    idx = torch.randint(0, eq.nx, (n * M_I, v), generator=g) if v else None
    return Draws(u_t, z_x0, z_x, zT, u_s, zS, idx)


def sample_tx(eq, d: Draws, eps: float = 0.01) -> Tensor:
    """data.py:161-167 (t), equations.py:118-119 + sample_x0 + 217-230 (x)."""
    t = (eq.T - 2 * eps) * (1 - d.u_t) + eps
    x0 = eq.x0_std() * d.z_x0 if d.z_x0 is not None else torch.zeros_like(d.z_x)
    x = x0 + torch.sqrt(t - 0.0) * eq.alpha_sqrt * d.z_x
    return torch.cat([t, x], dim=-1)


# ----------------------------------------------------------------------------------------------
# Estimators (data.py)
# ----------------------------------------------------------------------------------------------
def estimate_terminal_with_gradients(eq, tx: Tensor, zT: Tensor, M_T: int) -> Tensor:
    """data.py:899-926."""
    n = tx.shape[0]
    t = tx[:, :1].repeat_interleave(M_T, 0)
    x = tx[:, 1:].repeat_interleave(M_T, 0)
    XT = x + torch.sqrt(eq.T - t) * eq.alpha_sqrt * zT
    gT = eq.g(XT)
    Y = zT / torch.sqrt(eq.T - t) / eq.alpha_sqrt
    eY = torch.cat([torch.ones_like(gT), Y], dim=-1)
    g_single = eq.g(tx[:, 1:])
    term = (gT - g_single.repeat_interleave(M_T, 0)) * eY
    term = term.view(n, M_T, -1).sum(1) / M_T
    term[:, 0:1] = term[:, 0:1] + g_single
    return term


def get_f(eq, net, s: Tensor, Xs: Tensor, idx: Optional[Tensor], baseline_repeat: Optional[int], mode: str) -> Tensor:
    """data.py:1226-1325.  Returns f [rows(*repeat),1]."""
    if net.kind == "zero":
        # ZeroSolution: u=0, u_x=None->zeros (data.py:1257-1260 for the Hessian case, :1316 otherwise)
        u = torch.zeros(Xs.shape[0], 1, dtype=Xs.dtype)
        if eq.has_hessian_term:
            nrep = baseline_repeat or 1
            f = eq.ffi(s, Xs, u, torch.zeros(Xs.shape[0], eq.nx, dtype=Xs.dtype))  # |0| sums: only source remains
        else:
            f = eq.ff(s, Xs, u, torch.zeros_like(Xs))
        return f.repeat_interleave(baseline_repeat, 0) if baseline_repeat else f

    if mode == "autograd":
        return _get_f_autograd(eq, net, s, Xs, idx, baseline_repeat)

    sX = torch.cat([s, Xs], dim=-1)
    if eq.name == "Cha":
        u, du, _ = net.jets(sX, None, order=1)
        f = eq.ff_from_sum(u, du)
    elif eq.name == "GBM":
        if baseline_repeat is None:
            u, _, d2 = net.jets(sX, idx, order=2)
            f = eq.ffi(s, Xs, u, d2)
        else:
            full = torch.arange(eq.nx).repeat(Xs.shape[0], 1)
            u, _, d2 = net.jets(sX, full, order=2)  # full Hessian diagonal at (t,x) (data.py:1276-1278)
            d2 = d2.repeat_interleave(baseline_repeat, 0).gather(1, idx)
            return eq.ffi(s.repeat_interleave(baseline_repeat, 0), Xs.repeat_interleave(baseline_repeat, 0),
                          u.repeat_interleave(baseline_repeat, 0), d2)
    elif eq.name == "OU":
        if net.kind == "pisgrad":
            u, gx = net.value_and_grad_closed(sX, eq)
        else:  # plain MLP under the OU equation: full gradient by reverse mode
            return _get_f_autograd(eq, net, s, Xs, idx, baseline_repeat)
        f = eq.ff(s, Xs, u, gx)
    else:
        raise NotImplementedError(eq.name)
    return f.repeat_interleave(baseline_repeat, 0) if baseline_repeat else f


def _net_forward(eq, net, sX):
    return net.forward(sX, eq.g) if net.kind == "pisgrad" else net.forward(sX)


def _get_f_autograd(eq, net, s, Xs, idx, baseline_repeat):
    """Op-for-op mirror of data.py:1226-1325 (the arm the CPU baseline times)."""
    with torch.enable_grad():
        _net_forward(eq, net, torch.zeros(1, 1 + eq.nx, dtype=Xs.dtype))  # output-dim probe, data.py:1228
        Xs = Xs.detach().clone().requires_grad_()
        u = _net_forward(eq, net, torch.cat([s, Xs], dim=-1))
        u_x = torch.autograd.grad(u.sum(), Xs, create_graph=True, retain_graph=True, allow_unused=True)[0]
        if eq.has_hessian_term:
            ind_eff = idx if baseline_repeat is None else torch.arange(eq.nx).repeat(Xs.shape[0], 1)
            u_ii_all = torch.zeros(ind_eff.shape, dtype=Xs.dtype)
            for i in range(ind_eff.shape[1]):
                a = ind_eff[:, i].unsqueeze(-1)
                act = torch.gather(u_x, 1, a)
                u_ii = torch.autograd.grad(act, Xs, grad_outputs=torch.ones_like(act), create_graph=True)[0]
                u_ii_all[:, i] = torch.gather(u_ii, 1, a).squeeze(-1).detach()
            if baseline_repeat is None:
                return eq.ffi(s, Xs.detach(), u.detach(), u_ii_all)
            u_ii_all = u_ii_all.repeat_interleave(baseline_repeat, 0).gather(1, idx)
            return eq.ffi(s.repeat_interleave(baseline_repeat, 0), Xs.detach().repeat_interleave(baseline_repeat, 0),
                          u.detach().repeat_interleave(baseline_repeat, 0), u_ii_all)
        f = eq.ff(s, Xs.detach(), u.detach(), u_x.detach())
    return f.repeat_interleave(baseline_repeat, 0) if baseline_repeat else f


def estimate_integral_with_gradients(eq, net, tx: Tensor, u_s: Tensor, zS: Tensor, idx: Optional[Tensor], M_I: int,
                                     mode: str = "closed") -> Tensor:
    """data.py:471-527 (+ generate_sx_for_integral, data.py:350-366)."""
    n = tx.shape[0]
    t = tx[:, :1].repeat_interleave(M_I, 0)
    x = tx[:, 1:].repeat_interleave(M_I, 0)
    s = u_s * (eq.T - t) + t
    Xs = x + torch.sqrt(s - t) * eq.alpha_sqrt * zS
    f = get_f(eq, net, s, Xs, idx, None, mode)
    f_b = get_f(eq, net, tx[:, :1], tx[:, 1:], idx, M_I, mode)
    Ys = zS / torch.sqrt(s - t) / eq.alpha_sqrt
    eYs = torch.cat([torch.ones_like(s), Ys], dim=-1)
    integral = (eq.T - t) * (f - f_b) * eYs
    integral[:, 0:1] = integral[:, 0:1] + f_b * (eq.T - t)
    return integral.view(n, M_I, -1).sum(1) / M_I


def estimate_terminal_with_gradients_td(eq, net, tx: Tensor, zT: Tensor, M_T: int, delta_t: float) -> Tensor:
    """data.py:928-952: the step ends at t_next = min(t + delta, T); u_k(t_next, X) replaces g(X) where t_next < T."""
    n = tx.shape[0]
    t = tx[:, :1].repeat_interleave(M_T, 0)
    x = tx[:, 1:].repeat_interleave(M_T, 0)
    t_next = torch.clip(t + delta_t, max=eq.T)
    XT = x + torch.sqrt(t_next - t) * eq.alpha_sqrt * zT
    if net.kind == "zero":
        u_next = torch.zeros_like(t)
    else:
        u_next = _net_forward(eq, net, torch.cat([t_next, XT], dim=1)).detach()
    gT = torch.where(t_next == eq.T, eq.g(XT), u_next)
    Y = zT / torch.sqrt(t_next - t) / eq.alpha_sqrt
    eY = torch.cat([torch.ones_like(gT), Y], dim=-1)
    g_single = eq.g(tx[:, 1:])
    term = (gT - g_single.repeat_interleave(M_T, 0)) * eY
    term = term.view(n, M_T, -1).sum(1) / M_T
    term[:, 0:1] = term[:, 0:1] + g_single
    return term


def estimate_integral_with_gradients_td(eq, net, tx: Tensor, u_s: Tensor, zS: Tensor, idx: Optional[Tensor], M_I: int,
                                        delta_t: float, mode: str = "closed") -> Tensor:
    """data.py:529-575: the integral estimator over [t, t_next]."""
    assert eq.name in ("Cha", "OU")  # data.py:533
    n = tx.shape[0]
    t = tx[:, :1].repeat_interleave(M_I, 0)
    x = tx[:, 1:].repeat_interleave(M_I, 0)
    t_next = torch.clip(t + delta_t, max=eq.T)
    s = u_s * (t_next - t) + t
    Xs = x + torch.sqrt(s - t) * eq.alpha_sqrt * zS
    f = get_f(eq, net, s, Xs, idx, None, mode)
    f_b = get_f(eq, net, tx[:, :1], tx[:, 1:], idx, M_I, mode)
    Ys = zS / torch.sqrt(s - t) / eq.alpha_sqrt
    eYs = torch.cat([torch.ones_like(s), Ys], dim=-1)
    integral = (t_next - t) * (f - f_b) * eYs
    integral[:, 0:1] = integral[:, 0:1] + f_b * (t_next - t)
    return integral.view(n, M_I, -1).sum(1) / M_I


def generate_with_gradients(eq, net, tx: Tensor, d: Draws, M_T: int, M_I: int, mode: str = "closed",
                            sample_bound: float = float("inf"), delta_t: float = 0.0) -> Tensor:
    """data.py:1208-1218 (TD branch when delta_t > 0) followed by the clip of data.py:222."""
    if delta_t > 0:
        y = (estimate_terminal_with_gradients_td(eq, net, tx, d.zT, M_T, delta_t)
             + estimate_integral_with_gradients_td(eq, net, tx, d.u_s, d.zS, d.idx, M_I, delta_t, mode))
    else:
        y = (estimate_terminal_with_gradients(eq, tx, d.zT, M_T)
             + estimate_integral_with_gradients(eq, net, tx, d.u_s, d.zS, d.idx, M_I, mode))
    return torch.clip(y.detach(), -sample_bound, sample_bound)


def sample_with_gradients(eq, net, n: int, M_T: int, M_I: int, v: Optional[int] = None, mode: str = "autograd",
                          generator: Optional[torch.Generator] = None, dtype=torch.float64):
    """data.py:211-223: draw (t,x) and all MC randomness in reference order, then label."""
    d = draw(eq, n, M_T, M_I, v, generator, dtype)
    tx = sample_tx(eq, d)
    return tx, generate_with_gradients(eq, net, tx, d, M_T, M_I, mode)


# ----------------------------------------------------------------------------------------------
# helpers to build nets with torch's default nn.Linear init (used by tests / bench for synthetic weights)
# ----------------------------------------------------------------------------------------------
def _linear(n_in, n_out, generator, dtype):
    bound = 1.0 / math.sqrt(n_in)  # nn.Linear default: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in))
    W = (torch.rand(n_out, n_in, generator=generator, dtype=dtype) * 2 - 1) * bound
    b = (torch.rand(n_out, generator=generator, dtype=dtype) * 2 - 1) * bound
    return W, b


def make_mlp(n_in: int, hidden: Sequence[int], n_out: int = 1, seed: int = 0, dtype=torch.float64,
             round_fp32: bool = True, gain: float = 1.0) -> MLP:
    g = torch.Generator().manual_seed(seed)
    dims = [n_in] + list(hidden) + [n_out]
    Ws, bs = [], []
    for i in range(len(dims) - 1):
        W, b = _linear(dims[i], dims[i + 1], g, dtype)
        W = W * gain
        if round_fp32:
            W, b = W.float().to(dtype), b.float().to(dtype)
        Ws.append(W)
        bs.append(b)
    return MLP(Ws, bs)


def make_pisgrad(nx: int, hidden: Sequence[int], T: float = 1.0, seed: int = 0, dtype=torch.float64,
                 round_fp32: bool = True, phase_scale: float = 0.0) -> PISGradNet:
    g = torch.Generator().manual_seed(seed)
    C = 64
    rnd = (lambda t: t.float().to(dtype)) if round_fp32 else (lambda t: t)

    def chain(dims):
        out = []
        for i in range(len(dims) - 1):
            W, b = _linear(dims[i], dims[i + 1], g, dtype)
            out.append((rnd(W), rnd(b)))
        return out

    t_enc = chain([2 * C, C, C])
    smooth = chain([2 * C] + [C] * (len(hidden) + 1) + [nx])
    main = chain([nx + C] + list(hidden) + [nx])
    phase = rnd(torch.randn(1, C, generator=g, dtype=dtype) * phase_scale)
    coeff = rnd(torch.linspace(0.1, 100.0, C, dtype=dtype).unsqueeze(0))
    return PISGradNet(phase, coeff, t_enc, smooth, main, T)


def rel_l2(a: Tensor, b: Tensor) -> float:
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))
