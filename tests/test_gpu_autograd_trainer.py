"""AutogradRegressionTrainer on the GPU: the graph-captured step follows the eager step."""
import copy
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_captured_step_matches_eager_step():
    from deeppicarditeration_b200 import equations as eqs
    from deeppicarditeration_b200 import solution as sol
    from deeppicarditeration_b200.trainer import AutogradRegressionTrainer

    class Cfg:
        TYPE, PISGRADNET, NEURONS, ACTIVATIONS, BOUND = "Value", True, [64, 64], ["ELU", "ELU"], None

    torch.manual_seed(1)
    ou = eqs.OUProcessEquation(nx=12, num_components=2, mean=torch.randn(2, 12), var=torch.ones(2, 12) * 2.0,
                               pi=torch.ones(2) / 2)
    base = sol.PicardSolution(ou, Cfg).float().cuda()
    tx = torch.rand(256, 13, device="cuda")
    y = torch.randn(256, 13, device="cuda")
    results = []
    for no_graph in (True, False):
        net = copy.deepcopy(base)
        if no_graph:
            os.environ["DPI_AUTOGRAD_NO_GRAPH"] = "1"
        else:
            os.environ.pop("DPI_AUTOGRAD_NO_GRAPH", None)
        try:
            tr = AutogradRegressionTrainer(net, lr=1e-3, scaler="FixedLossScaler", fixed_weight=0.1, device="cuda")
            tr.fit(tx, y, 64, 3)        # 12 steps: 3 eager, then 9 replays on the graph path
            tr.reset_optimizer()        # state cleared in place under the graph
            losses = tr.fit(tx, y, 64, 2)
            assert tr.step_count == 8 and (no_graph or tr._graph is not None)
            results.append(([p.detach().clone() for p in net.parameters()], losses.clone()))
            tr.close()
        finally:
            os.environ.pop("DPI_AUTOGRAD_NO_GRAPH", None)
    (pe, le), (pg, lg) = results
    for a, b in zip(pe, pg):
        assert torch.allclose(a, b, rtol=2e-4, atol=2e-6)
    assert torch.allclose(le, lg, rtol=1e-3)
