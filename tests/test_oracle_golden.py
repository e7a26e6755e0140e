"""Pin the CPU restatement (oracle/dpi_oracle.py) against outputs of the unmodified reference.

The reference has no tests / golden vectors of its own; the fixtures are its own outputs on
replayed draws (oracle/make_golden.py).  fp64 vs fp64, so the tolerance is round-off only.
"""
import pytest
import torch

from oracle import dpi_oracle as orc
from oracle import ref_shim
from tests.golden_util import Golden, golden_names

TOL = 1e-12


@pytest.mark.parametrize("name", golden_names())
def test_sample_tx_matches_reference(name):
    g = Golden(name)
    assert orc.rel_l2(orc.sample_tx(g.eq, g.draws), g.tx) < TOL


@pytest.mark.parametrize("name", golden_names())
def test_terminal_matches_reference(name):
    g = Golden(name)
    if g.delta_t > 0:  # TD estimators (data.py:928-952)
        yT = orc.estimate_terminal_with_gradients_td(g.eq, g.net, g.tx, g.draws.zT, g.M_T, g.delta_t)
    else:
        yT = orc.estimate_terminal_with_gradients(g.eq, g.tx, g.draws.zT, g.M_T)
    assert orc.rel_l2(yT, g.y_terminal) < TOL


@pytest.mark.parametrize("mode", ["closed", "autograd"])
@pytest.mark.parametrize("name", golden_names())
def test_integral_and_labels_match_reference(name, mode):
    g = Golden(name)
    if g.delta_t > 0:  # data.py:529-575
        yI = orc.estimate_integral_with_gradients_td(g.eq, g.net, g.tx, g.draws.u_s, g.draws.zS, g.draws.idx, g.M_I,
                                                     g.delta_t, mode)
    else:
        yI = orc.estimate_integral_with_gradients(g.eq, g.net, g.tx, g.draws.u_s, g.draws.zS, g.draws.idx, g.M_I, mode)
    if float(g.y_integral.norm()) > 0:
        assert orc.rel_l2(yI, g.y_integral) < 1e-10  # cancellation in (f-f_b)/sqrt(s-t) costs a few digits
    else:
        assert float(yI.abs().max()) == 0.0
    y = orc.generate_with_gradients(g.eq, g.net, g.tx, g.draws, g.M_T, g.M_I, mode, delta_t=g.delta_t)
    assert orc.rel_l2(y, g.y) < TOL * 10


def test_fixture_count_and_coverage():
    names = golden_names()
    assert {"burgers_100d_mlp128", "burgers_100d_zero", "burgers_10d_ragged", "gbm_100d_mlp64_sdgd",
            "gbm_100d_zero_sdgd", "hjb_100d_pisgrad", "hjb_100d_zero", "burgers_100d_mlp128_td", "burgers_10d_zero_td",
            "hjb_100d_pisgrad_td"} <= set(names)
    # the TD fixtures exercise both branches of `where(t_next == T, g, u_k)`
    for n in ("burgers_100d_mlp128_td", "hjb_100d_pisgrad_td"):
        g = Golden(n)
        at_T = (g.tx[:, 0] + g.delta_t >= g.T)
        assert bool(at_T.any()) and bool((~at_T).any()), n


def test_exact_solution_is_picard_fixed_point_statistically():
    """Known-answer check: with u_k = exact solution the labels converge to (u*, grad u*)."""
    eq = orc.Cha(nx=10, alpha=1.0, k=5.0, T=1.0)

    class Exact:
        kind = "exact"

    gen = torch.Generator().manual_seed(3)
    n, M = 16, 4096
    d = orc.draw(eq, n, M, M, None, gen)
    tx = orc.sample_tx(eq, d)
    t, x = tx[:, :1], tx[:, 1:]
    # integral part with the exact u: f = sigma^2 (k u - c) sum_j u_xj evaluated in closed form
    tt, xx = t.repeat_interleave(M, 0), x.repeat_interleave(M, 0)
    s = d.u_s * (eq.T - tt) + tt
    Xs = xx + torch.sqrt(s - tt) * eq.alpha_sqrt * d.zS
    f = eq.ff(s, Xs, eq.exact_solution(s, Xs), eq.u_x(s, Xs))
    yI0 = ((eq.T - tt) * f).view(n, M).mean(1, keepdim=True)
    yT = orc.estimate_terminal_with_gradients(eq, tx, d.zT, M)
    u_hat = yT[:, :1] + yI0
    assert orc.rel_l2(u_hat, eq.exact_solution(t, x)) < 3e-2


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference not mounted (GPU box)")
def test_live_reference_agrees_with_fixture():
    """In the build container, re-run the live reference for one case to detect shim / fixture drift."""
    from oracle import make_golden as mg
    import numpy as np
    import tempfile, os

    old = mg.GOLDEN_DIR
    with tempfile.TemporaryDirectory() as tmp:
        mg.GOLDEN_DIR = tmp
        try:
            case = [c for c in mg.CASES if c["name"] == "burgers_10d_ragged"][0]
            mg.make_case(**case)
            new = np.load(os.path.join(tmp, "burgers_10d_ragged.npz"))
        finally:
            mg.GOLDEN_DIR = old
    g = Golden("burgers_10d_ragged")
    assert np.allclose(new["y"], g.y.numpy(), rtol=1e-13, atol=1e-15)
