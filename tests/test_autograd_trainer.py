"""AutogradRegressionTrainer (the library path for PISGradNet and the non-fixed loss scalers) against the live
reference's `PicardSolutionGradientWrapper.training_step` + Adam (solution_jac.py:40-78,167-213), on CPU in fp64."""
import copy

import pytest
import torch

from deeppicarditeration_b200 import equations as eqs
from deeppicarditeration_b200 import solution as sol
from deeppicarditeration_b200.trainer import AutogradRegressionTrainer
from oracle import ref_shim


def _cfgs(scaler, pis):
    C = ref_shim.cfgnode
    ncfg = C(TYPE="Value", PISGRADNET=pis, NEURONS=[32, 32], ACTIVATIONS=["ELU", "ELU"], BOUND=None)
    kw = C(fixed_weight=0.1) if scaler == "FixedLossScaler" else C()
    tcfg = C(NUM_HESS_SAMPLES=0,
             LOSS=C(beta=0.3, SCALER=C(cls=scaler, kwargs=kw), FN=C(cls=None, kwargs=C()), use_aux_loss=False, weight_aux_loss=0.1),
             OPTIMIZER=C(cls="Adam", kwargs=C(), SCHEDULER=C(cls=None, kwargs=C(), config=C())))
    return ncfg, tcfg


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference not mounted (GPU box)")
@pytest.mark.parametrize("scaler", ["FixedLossScaler", "SimpleLossScaler", "DimensionLossScaler"])
@pytest.mark.parametrize("pis", [False, True])
def test_loss_and_adam_step_match_reference(scaler, pis):
    ref_shim.install()
    import picard.equations as req
    import picard.solution as rsol
    import picard.solution_jac as rjac

    torch.set_default_dtype(torch.float64)
    try:
        torch.manual_seed(3)
        nx = 6
        if pis:
            req_eq = req.Cha(nx=nx, alpha=1.0, k=2.0, T=1.0)  # any g0 does: PISGradNet only calls equation.g
            my_eq = eqs.Cha(nx=nx, alpha=1.0, k=2.0, T=1.0)
        else:
            req_eq, my_eq = req.Cha(nx=nx, alpha=1.0, k=2.0, T=1.0), eqs.Cha(nx=nx, alpha=1.0, k=2.0, T=1.0)
        ncfg, tcfg = _cfgs(scaler, pis)
        rsolution = rsol.PicardSolution(req_eq, ncfg, tcfg)
        wrap = rjac.PicardSolutionGradientWrapper(rsolution)
        mine = sol.PicardSolution(my_eq, ncfg).double()
        mine.load_state_dict(copy.deepcopy(rsolution.state_dict()), strict=False)
        tx = torch.rand(16, 1 + nx)
        y = torch.randn(16, 1 + nx)
        tr = AutogradRegressionTrainer(mine, lr=1e-3, loss_beta=0.3, scaler=scaler, fixed_weight=0.1, device="cpu")
        total, v, g = tr.loss_terms(tx, y)
        ref_loss = wrap.training_step((tx.clone(), y.clone()), 0)
        assert torch.allclose(total.sum(), ref_loss.sum(), rtol=1e-10, atol=1e-12)
        # one optimiser step each
        opt = torch.optim.Adam(rsolution.parameters(), lr=1e-3)
        opt.zero_grad()
        wrap.training_step((tx.clone(), y.clone()), 0).sum().backward()
        opt.step()
        losses = tr.step(tx, y)
        assert abs(float(losses[0]) - float(ref_loss.detach().sum())) <= 1e-6 * max(1.0, abs(float(ref_loss.detach().sum())))
        ref_sd = rsolution.state_dict()  # (carries a `beta` buffer besides the network)
        for n1, p1 in mine.state_dict().items():
            assert torch.allclose(p1, ref_sd[n1], rtol=1e-9, atol=1e-12), n1
    finally:
        torch.set_default_dtype(torch.float32)
