"""Host-side mirror of the reference's `picard/equations.py` for the three equations on the label path.

Same class names, constructor kwargs, attributes and flags as the reference (file:line cited per class), so a config
that names `EQUATION.cls: Cha` etc. builds the same problem instance.  These objects are PARAMETER HOLDERS for the
CUDA library (`descriptor()` flattens them into the POD that `dpi_set_problem` takes) plus the closed-form leaves
(g, exact solutions) that the evaluation callback and the tests use on the host.  The Monte Carlo work itself
(sampling, f, g on path points) runs in libdpi_b200.so, never here.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch


class Equation:
    """picard/equations.py:62-133 (general form, flags, T/nx/nu)."""

    has_gradient_term = None
    has_laplacian_term = None
    has_hessian_term = None
    supported_approximate_methods: tuple = tuple()

    def __init__(self, T: float = 1.0, nx: int = 1):
        self.T = float(T)
        self.nx = int(nx)
        self.nu = 1
        self._device = torch.device("cpu")

    def to(self, device):  # ParametersMixin.to, equations.py:21-34
        self._device = torch.device(device)
        for k, v in list(self.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(self._device))
        return self

    # --- sampling of collocation points on the host (the product path uses dpi_sample_tx) ---
    def sample_x0(self, n: int):
        return torch.zeros(n, self.nx, device=self._device)

    def sample_x_ts(self, t, s, x, return_dW: bool = False):  # equations.py:217-230, 560-573 (drift-free)
        dW = torch.randn_like(x)
        x_next = x + torch.sqrt(s - t) * self.alpha_sqrt * dW
        return (x_next, dW) if return_dW else x_next

    def sample_x(self, t):  # equations.py:118-119
        return self.sample_x_ts(torch.zeros_like(t), t, self.sample_x0(len(t)))

    def f(self, t, x, y):  # equations.py:262-263
        raise NotImplementedError("The equation has dependence on z, use fff or ff instead.")

    def descriptor(self) -> dict:
        raise NotImplementedError


class SimpleDiffusionEquation(Equation):
    """Sigma = sqrt(alpha) I, mu = 0 (equations.py:180-230)."""

    def __init__(self, *, alpha: float = 1.0, **kwargs):
        super().__init__(**kwargs)
        self.alpha = torch.scalar_tensor(float(alpha))
        self.alpha_sqrt = torch.sqrt(self.alpha)

    def ff(self, t, x, y, w):  # equations.py:199-200
        return self.fff(t, x, y, self.alpha_sqrt * w)


class Cha(SimpleDiffusionEquation):
    """Burgers-type equation with exact solution sigmoid(t + k sum x) (equations.py:266-338)."""

    has_gradient_term, has_laplacian_term, has_hessian_term = True, False, False

    def __init__(self, nx: int, alpha: float, k=1.0, T: float = 1.0):
        super().__init__(nx=nx, alpha=alpha, T=T)
        self.k_config = float(k)
        self.k = torch.scalar_tensor(float(k) / math.sqrt(self.nx))  # equations.py:285
        self.k_alpha_d = self.k * self.alpha * self.nx
        self.c = (2 + self.k * self.k_alpha_d) / (2 * self.k_alpha_d)

    def fff(self, t, x, y, z):  # equations.py:297-302
        return self.alpha_sqrt * (self.k * y - self.c) * torch.sum(z, dim=-1, keepdim=True)

    def g(self, x):  # equations.py:304-305
        return torch.sigmoid(self.T + self.k * torch.sum(x, dim=-1, keepdim=True))

    def exact_solution(self, t, x):  # equations.py:318-319
        return torch.sigmoid(t + self.k * torch.sum(x, dim=-1, keepdim=True))

    def u_x(self, t, x):  # equations.py:325-327
        u = self.exact_solution(t, x)
        return torch.ones_like(x) * (self.k * u * (1 - u))

    def u_u_x(self, t, x):  # equations.py:329-332
        return self.exact_solution(t, x), self.u_x(t, x)

    def descriptor(self) -> dict:
        return dict(kind="Cha", nx=self.nx, T=self.T, alpha=float(self.alpha), cha_k=self.k_config)


class GBMEquationComplexExact(SimpleDiffusionEquation):
    """Fully nonlinear test equation u_t + 1/2 Lap u + 1/4 sum |u_ii| - f = 0 with the exact solution
    sum_k v_k sin(w_k . [t, x]) (equations.py:388-486).  `w` [nodes, 1+nx] and `v` [nodes, 1] are loaded from the CWD
    like the reference (`gbm_2nodes_{w,v}_{nx}d.pt`) unless passed in."""

    has_gradient_term, has_laplacian_term, has_hessian_term = True, False, True
    supported_approximate_methods = ("SDGD",)

    def __init__(self, nx: int, alpha: float = 1.0, T: float = 1.0, w: Optional[torch.Tensor] = None,
                 v: Optional[torch.Tensor] = None):
        super().__init__(nx=nx, alpha=alpha, T=T)
        self.d = torch.scalar_tensor(float(nx))
        if w is None or v is None:
            wf, vf = f"gbm_2nodes_w_{nx}d.pt", f"gbm_2nodes_v_{nx}d.pt"
            if os.path.exists(wf) and os.path.exists(vf):  # equations.py:408-411
                w, v = torch.load(wf), torch.load(vf)
            else:  # equations.py:412-419 (the reference also writes the files; kept)
                w = torch.randn(2, 1 + nx) / math.sqrt(nx)
                w[:, 0] = 1.0
                v = torch.randn(2, 1)
                torch.save(w, wf)
                torch.save(v, vf)
        self.w, self.v = w.detach().clone(), v.detach().clone().reshape(-1, 1)

    def _arg(self, t, x):
        t = t if torch.is_tensor(t) else torch.scalar_tensor(float(t), dtype=x.dtype, device=x.device)
        tx = torch.cat([t * torch.ones(x.shape[0], 1, dtype=x.dtype, device=x.device), x], dim=-1)
        return tx @ self.w.to(x).t()

    def exact_solution(self, t, x):  # equations.py:421-424
        return torch.sin(self._arg(t, x)) @ self.v.to(x)

    def g(self, x):  # equations.py:415-416
        return self.exact_solution(self.T, x)

    def u_t(self, t, x):  # equations.py:426-428
        return torch.cos(self._arg(t, x)) @ (self.v * self.w[:, 0:1]).to(x)

    def u_x(self, t, x):  # equations.py:430-432
        return torch.cos(self._arg(t, x)) @ (self.v * self.w[:, 1:]).to(x)

    def u_u_x(self, t, x):
        return self.exact_solution(t, x), self.u_x(t, x)

    def hessian_diagonal(self, t, x):  # diagonal of equations.py:445-450
        return -torch.sin(self._arg(t, x)) @ (self.v * self.w[:, 1:] ** 2).to(x)

    def u_hessian(self, t, x):  # equations.py:445-450: -sum_k v_k sin(a_k) w_k w_k^T, [n, nx, nx]
        sin_term = -torch.sin(self._arg(t, x))
        wx = self.w[:, 1:].to(x)
        weights = self.v.to(x).unsqueeze(-1) * (wx.unsqueeze(2) * wx.unsqueeze(1))
        return torch.einsum("ij,jkl->ikl", sin_term, weights)

    def ffi(self, t, x, u, u_ii):  # equations.py:457-466
        hd = self.hessian_diagonal(t, x)
        src = -self.u_t(t, x) - 0.5 * hd.sum(-1, keepdim=True) - 0.25 * hd.abs().sum(-1, keepdim=True)
        return (0.5 * (1.0 - self.alpha) * self.d * u_ii.mean(-1, keepdim=True)
                + 0.25 * self.d * u_ii.abs().mean(-1, keepdim=True) + src)

    def descriptor(self) -> dict:
        return dict(kind="GBM", nx=self.nx, T=self.T, alpha=float(self.alpha),
                    gbm_w=self.w.detach().double().cpu().numpy(), gbm_v=self.v.detach().double().cpu().numpy())


class OUProcessEquation(Equation):
    """HJB / OU-process equation with a diagonal Gaussian-mixture terminal condition
    (equations.py:488-596 ComplexDiffusionEquation, :599-714 OUProcessEquation; utils.py:792-915 GMM).
    `mean` [K,nx], `var` [K,nx,nx] (diagonal) and `pi` [K] are loaded from the CWD like the reference unless passed in;
    a missing `var` file is synthesised as var_scale * I, which is what the reference's generator would produce
     -- unlike the reference, existing mean/pi files are NOT overwritten in that case."""

    has_gradient_term, has_laplacian_term, has_hessian_term = True, False, False

    def __init__(self, nx, T=1.0, theta: float = 1.0, mu: float = 0.0, alpha: float = 1.0, num_components=2,
                 mean_scale=1.0, var_scale=2.0, alpha_scale=4.0, mean=None, var=None, pi=None):
        super().__init__(nx=nx, T=T)
        self.alpha = torch.scalar_tensor(float(alpha))
        self.alpha_sqrt = torch.sqrt(self.alpha)
        self.theta, self.mu = torch.scalar_tensor(float(theta)), torch.scalar_tensor(float(mu))
        self.d = torch.scalar_tensor(float(nx))
        self.num_components = K = int(num_components)
        tag = f"{nx}d_ms={mean_scale}_vs={var_scale}_{K}.pt"

        def load(name, make):
            return torch.load(f"{name}_{tag}") if os.path.exists(f"{name}_{tag}") else make()

        if mean is None:
            mean = load("mean", lambda: torch.stack([mean_scale * (torch.rand(nx) * 2 - 1) for _ in range(K)]))
        if var is None:
            var = load("var", lambda: torch.stack([var_scale * torch.eye(nx) for _ in range(K)]))
        if pi is None:
            pi = load("pi", lambda: (lambda p: p / p.sum())(torch.rand(K)))
        self.mean, self.var, self.pi = mean.detach().clone(), var.detach().clone(), pi.detach().clone()
        self.diagvar = torch.diagonal(self.var, dim1=-2, dim2=-1).clone() if self.var.dim() == 3 else self.var.clone()
        self.alpha_init = float(alpha_scale) * self.alpha  # equations.py:556

    def _log_terms(self, x, mean, diagvar):  # utils.py:852-880
        diff = x.unsqueeze(1) - mean.to(x).unsqueeze(0)
        expo = -0.5 * torch.einsum("bkn,kn->bk", diff ** 2, 1.0 / diagvar.to(x))
        norm = -0.5 * (self.nx * math.log(2.0 * math.pi) + torch.log(diagvar.to(x)).sum(1))
        return torch.log(self.pi.to(x)) + norm.unsqueeze(0) + expo, diff

    def g(self, x):  # equations.py:592-593
        lp, _ = self._log_terms(x, self.mean, self.diagvar)
        return -torch.logsumexp(lp, dim=1, keepdim=True)

    def g_x(self, x):  # equations.py:595-596, utils.py:882-915
        lp, diff = self._log_terms(x, self.mean, self.diagvar)
        resp = torch.softmax(lp, dim=1)
        return (diff / self.diagvar.to(x).unsqueeze(0) * resp.unsqueeze(-1)).sum(1)

    def ff(self, t, x, y, z):  # equations.py:657-666
        F = self.theta * (self.mu - x)
        return (-(F * z).sum(-1, keepdim=True) - self.alpha / 2 * (z ** 2).sum(-1, keepdim=True)
                - self.d * self.theta * torch.ones_like(y))

    def exact_solution(self, t, x):  # equations.py:625-655: the GMM pushed through the OU flow (all rows at once)
        lam = (self.T - t).reshape(-1, 1, 1)                                    # [b,1,1]
        e1, e2 = torch.exp(-self.theta * lam), torch.exp(-2 * self.theta * lam)
        mean_t = self.mu + (self.mean.to(x).unsqueeze(0) - self.mu) * e1         # [b,K,n]
        var_t = self.diagvar.to(x).unsqueeze(0) * e2 + (self.alpha / (2 * self.theta)) * (1 - e2)
        diff = x.unsqueeze(1) - mean_t
        expo = -0.5 * (diff ** 2 / var_t).sum(-1)
        norm = -0.5 * (self.nx * math.log(2.0 * math.pi) + torch.log(var_t).sum(-1))
        lp = torch.log(self.pi.to(x)).unsqueeze(0) + norm + expo
        return -torch.logsumexp(lp, dim=1, keepdim=True)

    def u_x(self, t, x, retain_graph=False):  # equations.py:675-686: by automatic differentiation, as upstream
        with torch.enable_grad():
            xr = x.detach().clone().requires_grad_(True)
            tt = t if t.shape[0] == xr.shape[0] else t.expand(xr.shape[0], -1)
            u = self.exact_solution(tt, xr)
            (ux,) = torch.autograd.grad(u.sum(), xr)
        return ux

    def u_u_x(self, t, x):  # equations.py:697-700
        return self.exact_solution(t, x), self.u_x(t, x)

    def sample_x0(self, n: int):  # equations.py:710-711, utils.py:785-789
        return torch.sqrt(self.alpha_init) * torch.randn(n, self.nx, device=self._device)

    def descriptor(self) -> dict:
        return dict(kind="OU", nx=self.nx, T=self.T, alpha=float(self.alpha),
                    ou_mean=self.mean.detach().double().cpu().numpy(),
                    ou_diagvar=self.diagvar.detach().double().cpu().numpy(),
                    ou_pi=self.pi.detach().double().cpu().numpy(), ou_theta=float(self.theta), ou_mu=float(self.mu),
                    ou_alpha_init=float(self.alpha_init))


def describe_equation(eq) -> dict:
    """Flatten one of OUR equations or one of the REFERENCE's (duck-typed: same attribute names) into the
    keyword arguments of `LabelEngine.set_problem`."""
    if hasattr(eq, "descriptor"):
        return eq.descriptor()
    name = type(eq).__name__
    nx, T, alpha = int(eq.nx), float(eq.T), float(eq.alpha)
    if name == "Cha":
        return dict(kind="Cha", nx=nx, T=T, alpha=alpha, cha_k=float(eq.k) * math.sqrt(nx))
    if name == "GBMEquationComplexExact":
        return dict(kind="GBM", nx=nx, T=T, alpha=alpha, gbm_w=eq.w.detach().double().cpu().numpy(),
                    gbm_v=eq.v.detach().double().cpu().numpy())
    if name == "OUProcessEquation":
        var = eq.var
        diagvar = torch.diagonal(var, dim1=-2, dim2=-1) if var.dim() == 3 else var
        return dict(kind="OU", nx=nx, T=T, alpha=alpha, ou_mean=eq.mean.detach().double().cpu().numpy(),
                    ou_diagvar=diagvar.detach().double().cpu().numpy(), ou_pi=eq.pi.detach().double().cpu().numpy(),
                    ou_theta=float(eq.theta), ou_mu=float(eq.mu), ou_alpha_init=float(eq.alpha_init))
    raise NotImplementedError(f"equation {name} is not on the B200 label path (Cha, GBMEquationComplexExact, "
                              f"OUProcessEquation are)")
