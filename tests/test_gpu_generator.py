"""GPU tests of the drop-in `OnlineDataGenerator` (same constructor keywords / methods as picard/data.py:369-437)."""
import pytest
import torch

from oracle import dpi_oracle as orc
from tests.golden_util import Golden

pytestmark = pytest.mark.gpu


class _HA:
    def __init__(self, method=None, v=None):
        self.method, self.kwargs = method, {"v": v}


class _NetCfg:
    TYPE, PISGRADNET, NEURONS, ACTIVATIONS, BOUND = "Value", False, [128] * 4, ["ELU"] * 4, None


def _load_mlp(net, g):
    lins = [m for m in net.model if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for m, W, b in zip(lins, g.net.weights, g.net.biases):
            m.weight.copy_(W)
            m.bias.copy_(b)


def _generator(g, **kw):
    from deeppicarditeration_b200 import equations as eqs, solution as sol
    from deeppicarditeration_b200.data import OnlineDataGenerator
    eq = eqs.Cha(nx=g.nx, alpha=g.alpha, k=g.k, T=g.T)
    if g.net.kind == "zero":
        net = sol.ZeroSolution()
    else:
        net = sol.PicardSolution(eq, _NetCfg)
        _load_mlp(net, g)
    # the keyword set PicardDataModule.get_data_generator passes (data.py:1465-1496)
    return OnlineDataGenerator(equation=eq, solution=net, N=80, i=2, device="cuda", t_always_uniform=True,
                               n_estimate_terminal=g.M_T, n_estimate_integral=g.M_I, hessian_approximation=_HA(),
                               sample_bound=None, estimate_terminal="OU_ByGx", estimate_integral="OU_Element",
                               estimate_delta_t=0.0, **kw)


@pytest.mark.parametrize("name,precision,tol", [("burgers_100d_mlp128", "fp32", 1e-4), ("burgers_100d_mlp128", "f16", 1e-2),
                                                ("burgers_100d_zero", "fp32", 1e-5)])
def test_generate_with_gradients_on_reference_draws(name, precision, tol):
    from tests.gpu_util import draws_dict
    g = Golden(name)
    gen = _generator(g, precision=precision)
    y = gen.generate_with_gradients(g.tx.float().cuda(), draws=draws_dict(g.draws))
    assert y.dtype == torch.float32 and y.is_cuda and tuple(y.shape) == tuple(g.y.shape)
    assert orc.rel_l2(y.double().cpu(), g.y) < tol
    assert all(not p.requires_grad for p in gen.solution.parameters())  # frozen like data.py:409-412


def test_sample_with_gradients_contract_and_streams():
    g = Golden("burgers_100d_mlp128")
    gen = _generator(g, seed=3, out_dtype=torch.float64)
    tx, y = gen.sample_with_gradients(32)
    assert tx.shape == (32, 101) and y.shape == (32, 101) and tx.dtype == torch.float64 and tx.is_cuda
    t = tx[:, 0]
    assert float(t.min()) >= 0.01 - 1e-6 and float(t.max()) <= 0.99 + 1e-6 and torch.isfinite(y).all()
    tx2, _ = gen.sample_with_gradients(32)  # the stream advances: fresh points
    assert not torch.equal(tx, tx2)
    # a second generator with the same seed replays the same stream; two calls of 32 == one call of 64
    gen_b = _generator(g, seed=3, out_dtype=torch.float64)
    tx_b, y_b = gen_b.sample_with_gradients(64)
    assert torch.equal(tx_b[:32], tx) and torch.equal(tx_b[32:], tx2) and torch.equal(y_b[:32], y)
    # ranks own disjoint index ranges
    gen_r1 = _generator(g, seed=3, rank=1, world_size=2, points_per_rank=64)
    tx_r1, _ = gen_r1.sample_with_gradients(8)
    gen_all = _generator(g, seed=3)  # one rank owning the whole index range: its points 64.. are rank 1's block
    tx_all, _ = gen_all.sample_with_gradients(72)
    assert torch.equal(tx_all[64:], tx_r1)
    gen_r0 = _generator(g, seed=3, rank=0, world_size=2, points_per_rank=64)
    gen_r0.sample_with_gradients(64)
    with pytest.raises(RuntimeError):
        gen_r0.sample_with_gradients(1)  # a rank never spills into its neighbour's block


def test_dataset_with_gradients_iterates_batches():
    g = Golden("burgers_100d_mlp128")
    gen = _generator(g)
    gen.n_estimate_terminal = gen.n_estimate_integral = 64
    data = gen.dataset_with_gradients(n_total=64, n_batch_buffer=2, batch_size=16)
    batches = list(data)
    assert len(batches) == 4 and all(x.shape == (16, 101) and y.shape == (16, 101) for x, y in batches)


def test_unreachable_paths_raise_like_the_reference():
    g = Golden("burgers_100d_zero")
    gen = _generator(g)
    with pytest.raises(NotImplementedError):
        gen.sample(4)  # value-only generator: equation.f raises for every shipped equation
    from deeppicarditeration_b200 import equations as eqs, solution as sol
    from deeppicarditeration_b200.data import OnlineDataGenerator
    eq = eqs.Cha(nx=10, alpha=1.0, k=5.0)
    with pytest.raises(ValueError):
        OnlineDataGenerator(eq, sol.ZeroSolution(), 10, 1, device="cuda", t_always_uniform=True,
                            hessian_approximation=_HA(), estimate_delta_t=-0.1)
    td = OnlineDataGenerator(eq, sol.ZeroSolution(), 10, 1, device="cuda", t_always_uniform=True, n_estimate_terminal=32,
                             n_estimate_integral=32, hessian_approximation=_HA(), estimate_delta_t=0.1)
    assert torch.isfinite(td.sample_with_gradients(8)[1]).all()  # the TD estimators are built (tests/test_gpu_parity.py)
    td.engine.close()
    with pytest.raises(AssertionError):
        OnlineDataGenerator(eq, sol.ZeroSolution(), 10, 1, device="cuda", t_always_uniform=True,
                            hessian_approximation=_HA("SDGD", 10))  # Cha does not support SDGD (data.py:119-121)
