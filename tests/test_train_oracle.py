"""The regression-step oracle (oracle/train_oracle.py) against fixtures generated from the UNMODIFIED reference
(`PicardSolutionGradientWrapper.training_step` + torch.optim.Adam; oracle/make_train_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import train_oracle as tro

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["train_burgers_20d", "train_burgers_100d"]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    n = 2 * int(z["n_lin"])
    t = lambda a: torch.from_numpy(np.asarray(a)).double()
    return dict(z=z, p0=[t(z[f"p0_{i}"]) for i in range(n)], g0=[t(z[f"g0_{i}"]) for i in range(n)],
                pK=[t(z[f"pK_{i}"]) for i in range(n)], tx=t(z["tx"]), y=t(z["y"]), B=int(z["B"]), steps=int(z["steps"]),
                beta=float(z["beta"]), fw=float(z["fixed_weight"]), losses=t(z["losses"]))


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("mode", ["autograd", "closed"])
def test_first_step_gradients_match_reference(name, mode):
    g = load(name)
    fn = tro.loss_and_grads_autograd if mode == "autograd" else tro.loss_and_grads_closed
    B = g["B"]
    loss, v_loss, g_loss, grads = fn(g["p0"], g["tx"][:B], g["y"][:B], g["beta"], g["fw"])
    assert abs(float(loss) - float(g["losses"][0])) < 1e-12 * max(1.0, abs(float(loss)))
    for a, b in zip(grads, g["g0"]):
        assert rel(a.reshape(b.shape), b) < 1e-10


@pytest.mark.parametrize("name", NAMES)
def test_adam_trajectory_matches_reference(name):
    g = load(name)
    params = [p.clone() for p in g["p0"]]
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    B = g["B"]
    for s in range(g["steps"]):
        loss, _, _, grads = tro.loss_and_grads_closed(params, g["tx"][s * B:(s + 1) * B], g["y"][s * B:(s + 1) * B],
                                                      g["beta"], g["fw"])
        assert abs(float(loss) - float(g["losses"][s])) < 1e-9 * max(1.0, abs(float(loss)))
        tro.adam_step(params, [gr.reshape(p.shape) for gr, p in zip(grads, params)], m, v, s + 1)
    for a, b in zip(params, g["pK"]):
        assert rel(a, b) < 1e-9
