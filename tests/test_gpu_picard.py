"""End-to-end check (north star, second criterion): a short Deep Picard Iteration run on the GPU stages of this
library (label kernels + regression step) must drive the error against the exact solution of the Burgers-type
equation down, iteration over iteration."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_picard_iterations_converge_on_burgers_20d():
    from deeppicarditeration_b200 import equations as eqs
    from deeppicarditeration_b200.picard import run_picard
    eq = eqs.Cha(nx=20, alpha=1.0, k=5.0, T=1.0)
    out = run_picard(eq, [128] * 4, n_iterations=6, data_size=2048, n_mc=1024, batch_size=256, n_epochs=24,
                     precision="f16", eval_points=4000, log=None)
    h = out["history"]
    assert all(torch.isfinite(torch.tensor([m["rRMSE"], m["rRMSEg"]])).all() for m in h)
    assert h[-1]["rRMSE"] < 0.12, h
    assert h[-1]["rRMSE"] < 0.6 * h[0]["rRMSE"], h
    assert h[-1]["rRMSEg"] < 0.6 * h[0]["rRMSEg"], h
