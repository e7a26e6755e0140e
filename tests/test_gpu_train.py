"""GPU parity of the regression step (dpi_train_grad / dpi_train_adam through the C ABI) against fixtures generated
from the UNMODIFIED reference (`PicardSolutionGradientWrapper.training_step` + torch.optim.Adam, fp64)."""
import pytest
import torch

from tests.test_train_oracle import NAMES, load, rel

pytestmark = pytest.mark.gpu


def _model(g, nx):
    n_lin = len(g["p0"]) // 2
    layers = []
    for i in range(n_lin):
        W, b = g["p0"][2 * i], g["p0"][2 * i + 1]
        lin = torch.nn.Linear(W.shape[1], W.shape[0])
        with torch.no_grad():
            lin.weight.copy_(W)
            lin.bias.copy_(b)
        layers.append(lin)
        if i < n_lin - 1:
            layers.append(torch.nn.ELU())
    return torch.nn.Sequential(*layers).float().cuda()


@pytest.mark.parametrize("name", NAMES)
def test_first_step_loss_and_gradient_match_reference(name):
    from deeppicarditeration_b200.trainer import RegressionTrainer
    g = load(name)
    model = _model(g, int(g["z"]["nx"]))
    tr = RegressionTrainer(model, loss_beta=g["beta"], fixed_weight=g["fw"], max_batch=g["B"])
    B = g["B"]
    losses = tr.backward(g["tx"][:B], g["y"][:B]).double().cpu()
    assert abs(float(losses[0]) - float(g["losses"][0])) < 1e-5 * abs(float(g["losses"][0]))
    flat_ref = torch.cat([x.reshape(-1) for x in g["g0"]])
    assert rel(tr.grad.double().cpu(), flat_ref) < 1e-5
    # per-tensor check (a small tensor must not hide behind a large one)
    off = 0
    for x in g["g0"]:
        n = x.numel()
        assert rel(tr.grad[off:off + n].double().cpu(), x.reshape(-1)) < 1e-4
        off += n
    tr.close()


@pytest.mark.parametrize("name", NAMES)
def test_adam_trajectory_matches_reference(name):
    from deeppicarditeration_b200.trainer import RegressionTrainer
    g = load(name)
    model = _model(g, int(g["z"]["nx"]))
    tr = RegressionTrainer(model, loss_beta=g["beta"], fixed_weight=g["fw"], max_batch=g["B"])
    B = g["B"]
    for s in range(g["steps"]):
        losses = tr.step(g["tx"][s * B:(s + 1) * B], g["y"][s * B:(s + 1) * B]).double().cpu()
        assert abs(float(losses[0]) - float(g["losses"][s])) < 1e-4 * abs(float(g["losses"][s]))
    params = [p for m in model if isinstance(m, torch.nn.Linear) for p in (m.weight, m.bias)]
    for a, b in zip(params, g["pK"]):
        # fp32 Adam vs fp64 Adam after a few steps: the update is lr * sign-like, so compare the displacement
        assert rel(a.detach().double().cpu(), b) < 1e-5
    tr.close()


def test_fit_reduces_the_loss_on_a_learnable_target():
    """Inner loop of one Picard iteration on labels produced by a teacher network: the loss must go down."""
    from deeppicarditeration_b200.trainer import RegressionTrainer
    torch.manual_seed(0)
    nx = 30
    teacher = torch.nn.Sequential(torch.nn.Linear(1 + nx, 64), torch.nn.ELU(), torch.nn.Linear(64, 1)).double()
    tx = torch.cat([torch.rand(2048, 1), torch.randn(2048, nx)], 1).double().requires_grad_()
    u = teacher(tx)
    (du,) = torch.autograd.grad(u.sum(), tx)
    y = torch.cat([u.detach(), du[:, 1:]], 1)
    student = torch.nn.Sequential(torch.nn.Linear(1 + nx, 64), torch.nn.ELU(), torch.nn.Linear(64, 64), torch.nn.ELU(),
                                  torch.nn.Linear(64, 1)).float().cuda()
    tr = RegressionTrainer(student, fixed_weight=1.0, max_batch=256)
    first = float(tr.backward(tx.detach()[:256], y[:256])[0])
    last = float(tr.fit(tx.detach(), y, batch_size=256, n_epochs=30)[0])
    assert last < 0.2 * first
    tr.close()
