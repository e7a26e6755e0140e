"""Host-side mirror of the u_k network definitions in the reference's `picard/solution.py` (forward only).

Same class names, constructor arguments and state_dict key layout (`model.<idx>.weight/bias`), as plain
`torch.nn.Module`s (the reference derives from LightningModule, which only matters for its trainer).  On the label
path these modules are WEIGHT CONTAINERS: `describe_solution()` flattens them into the layer list that
`dpi_set_network` packs; their `forward` is used on the host by the evaluation callback and the tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def construct_mlp(n_in: int, n_out: int, n_neurons: list, activations: list, bound: Optional[float]):
    """picard/solution.py:123-135."""
    assert len(n_neurons) == len(activations)
    layers = []
    dims = [n_in] + list(n_neurons)
    for i in range(len(activations)):
        layers.append(torch.nn.Linear(dims[i], dims[i + 1]))
        layers.append(getattr(torch.nn, activations[i])())
    layers.append(torch.nn.Linear(dims[-1], n_out))
    if bound is not None:
        assert bound > 0
        layers.append(torch.nn.Hardtanh(-bound, bound))
    return torch.nn.Sequential(*layers)


class PISGradNet(torch.nn.Module):
    """picard/solution.py:138-289: sin/cos time embedding with learnable phase, smoothing net, main net
    `nn_module([t_emb, x]) -> R^dim`, output `smooth <net_out, x> + (1 - smooth) g0(exp(-lambda/2) x)`."""

    def __init__(self, hidden_shapes: list, dim: int, g0, T=1.0):
        super().__init__()
        self.hidden_shapes, self.n_layers, self.dim = list(hidden_shapes), len(hidden_shapes), dim
        self.act = torch.nn.ELU()
        self.channels = C = 64
        self.timestep_phase = torch.nn.Parameter(torch.zeros(1, C))
        self.register_buffer("timestep_coeff", torch.linspace(0.1, 100.0, steps=C).unsqueeze(0))
        self.t_encoder = torch.nn.Sequential(torch.nn.Linear(2 * C, C), self.act, torch.nn.Linear(C, C))
        smooth = [torch.nn.Linear(2 * C, C)]
        for _ in range(self.n_layers):
            smooth += [self.act, torch.nn.Linear(C, C)]
        smooth += [self.act, torch.nn.Linear(C, dim)]
        self.smooth_net = torch.nn.Sequential(*smooth)
        net, n_in = [], dim + C
        for hs in self.hidden_shapes:
            net += [torch.nn.Linear(n_in, hs), self.act]
            n_in = hs
        net.append(torch.nn.Linear(n_in, dim))
        self.nn_module = torch.nn.Sequential(*net)
        self.g0, self.T = g0, T

    def get_pis_timestep_embedding(self, lbd):
        arg = self.timestep_coeff * lbd + self.timestep_phase
        return torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)

    def smoothing_function(self, lbd):
        if lbd.ndim == 1:
            lbd = lbd.unsqueeze(-1)
        out = self.smooth_net(self.get_pis_timestep_embedding(lbd))
        out0 = self.smooth_net(self.get_pis_timestep_embedding(torch.zeros_like(lbd)))
        return out[..., 0:1] - out0[..., 0:1]

    def forward(self, tx):
        lbd, x = self.T - tx[..., 0:1], tx[..., 1:]
        smooth = self.smoothing_function(lbd)
        t_emb = self.t_encoder(self.get_pis_timestep_embedding(lbd))
        net_out = self.nn_module(torch.cat([t_emb, x], dim=-1))
        sp_out = torch.sum(net_out * x, dim=-1, keepdim=True)
        return smooth * sp_out + (1.0 - smooth) * self.g0(torch.exp(-0.5 * lbd) * x)


class PicardSolution(torch.nn.Module):
    """picard/solution.py:292-327.  `network_cfg` needs TYPE, PISGRADNET, NEURONS, ACTIVATIONS, BOUND."""

    def __init__(self, equation, network_cfg, train_cfg=None):
        super().__init__()
        self.equation, self.train_cfg = equation, train_cfg
        typ = network_cfg.TYPE
        if typ == "Value":
            self.output_dim = equation.nu
        elif typ == "ValueGradient":
            self.output_dim = equation.nu + equation.nx
        elif typ == "OnlyGradient":
            self.output_dim = equation.nx
        else:
            raise ValueError(f"Unknown network type {typ}")
        if network_cfg.PISGRADNET:
            self.model = PISGradNet(hidden_shapes=network_cfg.NEURONS, dim=equation.nx, g0=equation.g, T=equation.T)
        else:
            self.model = construct_mlp(1 + equation.nx, self.output_dim, list(network_cfg.NEURONS),
                                       list(network_cfg.ACTIVATIONS), network_cfg.BOUND)

    def forward(self, tx):
        return self.model(tx)


class ZeroSolution(torch.nn.Module):
    """picard/solution.py:330-337: u_0 == 0 (its input gradient is treated as zero, data.py:1316)."""

    def __init__(self, output_dim: int = 1):
        super().__init__()
        self.output_dim = output_dim

    def forward(self, tx):
        return torch.zeros(tx.shape[0], self.output_dim, dtype=tx.dtype, device=tx.device)


# -----------------------------------------------------------------------------------------------------
def _linears(seq, what: str) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """(W, b) of every Linear in a Sequential whose other members are ELU; anything else is not on the B200 path."""
    out = []
    for m in seq:
        if isinstance(m, torch.nn.Linear):
            out.append((m.weight.detach(), m.bias.detach()))
        elif isinstance(m, torch.nn.ELU):
            if m.alpha != 1.0:
                raise NotImplementedError(f"{what}: ELU(alpha={m.alpha}) is not built (alpha=1 only)")
        else:
            raise NotImplementedError(f"{what}: layer {type(m).__name__} is not built on the B200 label path "
                                      f"(Linear + ELU chains are; NETWORK.BOUND must be None)")
    return out


def describe_solution(solution) -> dict:
    """Flatten a frozen solution -- ours or the reference's (duck-typed) -- into `LabelEngine.set_network` kwargs.

    Accepts ZeroSolution, PicardSolution(MLP | PISGradNet), a bare Sequential / PISGradNet, and the reference's
    PicardSolutionGradientWrapper (`.solution`, solution_jac.py:113-216: its forward delegates to the inner net).
    """
    name = type(solution).__name__
    if name == "ZeroSolution":
        if getattr(solution, "output_dim", 1) != 1:
            raise NotImplementedError("ZeroSolution(output_dim != 1) is not built")
        return dict(kind="zero")
    if name == "PicardSolutionGradientWrapper" or (hasattr(solution, "solution") and not hasattr(solution, "model")):
        return describe_solution(solution.solution)
    model = getattr(solution, "model", solution)
    if type(model).__name__ == "PISGradNet":
        return dict(kind="pisgrad", main=_linears(model.nn_module, "PISGradNet.nn_module"),
                    t_encoder=_linears(model.t_encoder, "PISGradNet.t_encoder"),
                    smooth_net=_linears(model.smooth_net, "PISGradNet.smooth_net"),
                    phase=model.timestep_phase.detach(), coeff=model.timestep_coeff.detach())
    if isinstance(model, torch.nn.Sequential):
        lins = _linears(model, name)
        if lins[-1][0].shape[0] != 1:
            raise NotImplementedError("only NETWORK.TYPE=Value (scalar output) is built on the B200 label path")
        return dict(kind="mlp", main=lins)
    raise NotImplementedError(f"solution of type {name} is not on the B200 label path")
