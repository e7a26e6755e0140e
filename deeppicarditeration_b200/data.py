"""Drop-in for the reference's label generator `OnlineDataGenerator` (picard/data.py:88-223, 369-437, 1208-1218).

Same constructor keywords and methods as the reference class; every Monte Carlo operation behind it runs in
libdpi_b200.so through the C ABI (include/dpi_b200.h).  There is no CPU fallback: constructing this class
without a CUDA device, or with a combination the library has not built, raises.

Differences that are deliberate and visible:
  * labels are float32 tensors (the reference's configs run in float64); pass `out_dtype=torch.float64` to get
    them cast for a float64 trainer.  Arithmetic: fp32 CUDA-core path (`precision="fp32"`, within 1e-4 rel of the
    fp64 reference on replayed draws) or the 16-bit tcgen05 tensor-core path with fp32 accumulation (`precision="f16"`:
    fp16 operands, within 1e-2).  `"bf16"` is rejected (not built) rather than aliased; labels that come out inf / NaN
    (a network outside fp16's range) raise FloatingPointError instead of being returned.
  * randomness is a counter-based Philox stream keyed by (`seed`, global point index) instead of torch's global
    generator, so the labels of a point do not depend on the batch size, the call pattern or the GPU count.
    Successive `sample*` calls advance the point index; `rank`/`world_size` give each GPU a disjoint index range.
  * `generate_with_gradients(tx, draws=...)` additionally accepts the reference's own draws (deterministic mode).
"""
from __future__ import annotations

import math
from typing import Optional, Union

import torch

from .dataset import IterableDatasetWithInternalBatch
from .engine import LabelEngine
from .equations import describe_equation
from .solution import describe_solution


class DataGenerator:
    """picard/data.py:53-85."""

    do_internal_batching = True

    def dataset(self, n_total, n_batch_buffer, batch_size):
        raise NotImplementedError

    def dataset_with_gradients(self, n_total, n_batch_buffer, batch_size):
        raise NotImplementedError


class OnlineDataGenerator(DataGenerator):
    def __init__(self, equation, solution: torch.nn.Module, N: int, i: int, *,
                 device: Union[str, torch.device] = "cuda", t_always_uniform=False, n_estimate_terminal=1,
                 n_estimate_integral=1, hessian_approximation=None, sample_bound=None, estimate_terminal="OU_ByGx",
                 estimate_integral="OU_Element", estimate_delta_t=0.0,
                 # ---- additions of this implementation (all optional) ----
                 precision: str = "fp32", seed: int = 0, rank: int = 0, world_size: int = 1,
                 points_per_rank: int = 1 << 40, out_dtype: Optional[torch.dtype] = None):
        self.equation, self.solution, self.N, self.i = equation, solution, N, i
        for p in getattr(solution, "parameters", lambda: [])():  # data.py:409-412
            p.requires_grad = False
        if hasattr(solution, "eval"):
            solution.eval()
        assert getattr(equation, "nu", 1) == 1, "Currently only nu=1 is supported"  # data.py:431
        # data.py:108-111: t_always_uniform selects U[eps, T - eps]; otherwise t = T (1 - prod of N - i + 1 uniforms)
        self.t_always_uniform = bool(t_always_uniform)
        self._t_factors = 0 if self.t_always_uniform else max(1, int(N) - int(i) + 1)
        # data.py:529-575, :928-952: > 0 selects the temporal-difference estimators (both integrate to min(t + delta, T),
        # the terminal one takes u_k there); upstream asserts SimpleDiffusionEquation or OUProcessEquation
        self.estimate_delta_t = float(estimate_delta_t or 0.0)
        if self.estimate_delta_t < 0:
            raise ValueError("estimate_delta_t must be >= 0")
        if self.estimate_delta_t > 0 and getattr(equation, "has_hessian_term", False):
            raise AssertionError("ESTIMATE_DELTA_T > 0 is defined for SimpleDiffusionEquation / OUProcessEquation only")
        method = getattr(hessian_approximation, "method", None) if hessian_approximation is not None else None
        self.sdgd_v = 0
        if method is not None:  # data.py:113-121
            assert method in equation.supported_approximate_methods, \
                f"Current equation does not support the method {method}"
            self.sdgd_v = int(hessian_approximation.kwargs["v"])
        self.n_estimate_terminal, self.n_estimate_integral = int(n_estimate_terminal), int(n_estimate_integral)
        self.sample_bound = float("inf") if sample_bound is None else float(sample_bound)  # data.py:123-128
        self.estimate_terminal_type, self.estimate_integral_type = estimate_terminal, estimate_integral
        # data.py:134-137: epsilon = 0.01 iff "ByGx" in ESTIMATE_TERMINAL or "Joint" in ESTIMATE_INTEGRAL
        eps_on = ("ByGx" in (estimate_terminal or "")) or ("Joint" in (estimate_integral or ""))
        self.generate_t_uniform_epsilon = 0.01 if eps_on else 0.0
        self.precision, self.seed, self.out_dtype = precision, int(seed), out_dtype
        # disjoint blocks of the global point index per rank; the Philox counter holds 48 bits of it
        if not (0 <= int(rank) < int(world_size)) or (int(rank) + 1) * int(points_per_rank) > (1 << 48):
            raise ValueError(f"rank {rank} of {world_size} with {points_per_rank} points per rank leaves the 48-bit "
                             "point index range of the Philox counter")
        self._base = int(rank) * int(points_per_rank)
        self._limit = int(points_per_rank)
        self._drawn = 0
        self._device = torch.device(device)
        self._engine: Optional[LabelEngine] = None
        self.to(self._device)

    # ------------------------------------------------------------------------------------------
    def to(self, device):  # data.py:139-143, 433-436
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("deeppicarditeration_b200.OnlineDataGenerator runs on a CUDA (sm_100a) device only; "
                               "there is no CPU fallback")
        if self._engine is not None:
            self._engine.close()
        self._device = device
        self._engine = LabelEngine(device)
        desc = dict(describe_equation(self.equation))
        kind = desc.pop("kind")
        self._engine.set_problem(kind, desc.pop("nx"), desc.pop("T"), desc.pop("alpha"), sdgd_v=self.sdgd_v, **desc)
        self.set_solution(self.solution)
        return self

    def set_solution(self, solution):
        """Re-pack the frozen network (call after the weights of `solution` changed, e.g. next Picard iteration)."""
        self.solution = solution
        net = dict(describe_solution(solution))
        self._engine.set_network(net.pop("kind"), **net)

    @property
    def device(self):
        return self._device

    @property
    def engine(self) -> LabelEngine:
        return self._engine

    def _out(self, t: torch.Tensor) -> torch.Tensor:
        return t if self.out_dtype is None else t.to(self.out_dtype)

    # ------------------------------------------------------------------------------------------
    def _sample_tx(self, n: int, base: int) -> torch.Tensor:
        """t_sampler + equation.sample_x (data.py:149-167, 211-219) on the device."""
        return self._engine.sample_tx(n, seed=self.seed, point_index_base=base, n_factors=self._t_factors,
                                      eps=self.generate_t_uniform_epsilon)

    def _advance(self, n: int) -> int:
        """Reserve n global point indices for a call: successive calls never share Philox streams."""
        if self._drawn + n > self._limit:
            raise RuntimeError("this rank's block of the global point index range is exhausted")
        base = self._base + self._drawn
        self._drawn += n
        return base

    def _label(self, tx: torch.Tensor, base: int, bound: float, draws: Optional[dict]) -> torch.Tensor:
        """One call of the custom op `picard_b200::label_gen` (engine.py) = one dpi_label_gen through the C ABI."""
        d = draws or {}
        y = torch.ops.picard_b200.label_gen(
            tx.to(self._device, torch.float32).contiguous(), self._engine.id, self.n_estimate_terminal,
            self.n_estimate_integral, self.precision, float(bound) if math.isfinite(bound) else 0.0, self.seed, int(base),
            d.get("zT"), d.get("u_s"), d.get("zS"), d.get("idx"), self.estimate_delta_t)
        self._engine.check_labels()  # raises FloatingPointError instead of handing out inf / NaN labels
        return y

    def generate_with_gradients(self, tx: torch.Tensor, draws: Optional[dict] = None, point_index_base=None):
        """data.py:1208-1218 (terminal + integral estimator, non-TD branch) -> y [n, 1+nx], UNCLIPPED like the
        reference's (the clip belongs to sample_with_gradients, data.py:222)."""
        if draws is not None:
            draws = {k: (None if v is None else v.to(self._device)) for k, v in draws.items()}
            if draws.get("idx") is not None:
                draws["idx"] = draws["idx"].to(torch.int32)
            base = 0 if point_index_base is None else point_index_base
        else:
            base = self._advance(tx.shape[0]) if point_index_base is None else point_index_base
        return self._out(self._label(tx, base, float("inf"), draws))

    def sample_with_gradients(self, n_batch: int):
        """data.py:211-223: draw (t, x), label, clip -> (tx [n,1+nx], u_ux [n,1+nx])."""
        base = self._advance(n_batch)
        tx = self._sample_tx(n_batch, base)
        y = self._label(tx, base, self.sample_bound, None)  # clip fused into the kernels' final write
        return self._out(tx), self._out(y)

    def generate(self, tx):  # data.py:1203-1206 -> estimate_integral -> equation.f, which raises for every shipped
        raise NotImplementedError("The equation has dependence on z, use fff or ff instead.")  # equation (F3)

    def sample(self, n_batch):  # data.py:196-209
        return self.generate(None)

    def generate_with_gradients_and_hessians(self, tx):  # data.py:1220-1223 (SUPERVISE_HESSIAN), SURVEY.md 8f-4
        raise NotImplementedError("Hessian labels (estimate_*_and_hessians_double) are not built yet")

    def sample_exact_with_gradients(self, n_batch):  # data.py:251-262
        tx = self._sample_tx(n_batch, self._advance(n_batch))
        t, x = tx[:, :1], tx[:, 1:]
        eq = self.equation.to(tx.device) if hasattr(self.equation, "to") else self.equation
        u, ux = eq.u_u_x(t, x)
        return self._out(tx), self._out(torch.cat([u, ux], dim=-1))

    def dataset_with_gradients(self, n_total, n_batch_buffer, batch_size):  # data.py:1191-1201 contract
        return IterableDatasetWithInternalBatch(n_total, n_batch_buffer, batch_size, self.sample_with_gradients)

    def dataset_exact_with_gradients(self, n_total, n_batch_buffer, batch_size):
        return IterableDatasetWithInternalBatch(n_total, n_batch_buffer, batch_size, self.sample_exact_with_gradients)
