"""Generate tests/golden/train_*.npz from the UNMODIFIED reference (run in the build container only):
`PicardSolutionGradientWrapper.training_step` (picard/solution_jac.py:172-216) + torch.optim.Adam steps on fixed batches."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def make(name, nx, hidden, B, steps, beta, fixed_weight, seed):
    ref_shim.install()
    import picard.equations as req
    import picard.solution as rsol
    import picard.solution_jac as rjac

    torch.set_default_dtype(torch.float64)
    torch.manual_seed(seed)
    eq = req.Cha(nx=nx, alpha=1.0, k=5.0, T=1.0)
    ncfg = ref_shim.cfgnode(TYPE="Value", PISGRADNET=False, NEURONS=list(hidden), ACTIVATIONS=["ELU"] * len(hidden), BOUND=None)
    tcfg = ref_shim.cfgnode(
        NUM_HESS_SAMPLES=0,
        LOSS=ref_shim.cfgnode(beta=beta, SCALER=ref_shim.cfgnode(cls="FixedLossScaler", kwargs=ref_shim.cfgnode(fixed_weight=fixed_weight)),
                              FN=ref_shim.cfgnode(cls=None, kwargs=ref_shim.cfgnode()), use_aux_loss=False, weight_aux_loss=0.1),
        OPTIMIZER=ref_shim.cfgnode(cls="Adam", kwargs=ref_shim.cfgnode(), SCHEDULER=ref_shim.cfgnode(cls=None, kwargs=ref_shim.cfgnode(), config=ref_shim.cfgnode())))
    sol = rsol.PicardSolution(eq, ncfg, tcfg)
    # fp32-representable initial weights (the CUDA path holds fp32 master weights)
    with torch.no_grad():
        for p in sol.parameters():
            p.copy_(p.float().double())
    wrap = rjac.PicardSolutionGradientWrapper(sol)
    params0 = [p.detach().clone() for p in sol.model.parameters()]
    opt = torch.optim.Adam(sol.model.parameters())
    g = torch.Generator().manual_seed(seed + 1)
    tx = torch.cat([torch.rand(steps * B, 1, generator=g) * 0.98 + 0.01, torch.randn(steps * B, nx, generator=g)], 1).float().double()
    y = torch.randn(steps * B, 1 + nx, generator=g).float().double() * 0.3
    out = dict(nx=nx, B=B, steps=steps, beta=beta, fixed_weight=fixed_weight, n_lin=len(hidden) + 1, tx=tx.numpy(), y=y.numpy())
    for i, p in enumerate(params0):
        out[f"p0_{i}"] = p.numpy()
    losses = []
    for s in range(steps):
        opt.zero_grad()
        loss = wrap.training_step((tx[s * B:(s + 1) * B], y[s * B:(s + 1) * B]), s)
        loss.backward()
        if s == 0:
            for i, p in enumerate(sol.model.parameters()):
                out[f"g0_{i}"] = p.grad.detach().clone().numpy()
        opt.step()
        losses.append(float(loss))
    out["losses"] = np.array(losses)
    for i, p in enumerate(sol.model.parameters()):
        out[f"pK_{i}"] = p.detach().numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "losses", losses[:3], "...", losses[-1])


if __name__ == "__main__":
    make("train_burgers_20d", nx=20, hidden=[32, 48, 32], B=24, steps=6, beta=0.0, fixed_weight=1.0, seed=3)
    make("train_burgers_100d", nx=100, hidden=[128] * 4, B=64, steps=4, beta=0.5, fixed_weight=0.3, seed=5)
