"""Measure the reference's run-to-run spread of the end-to-end error (nothing is seeded upstream, so the
"reference's reported spread" of north_star has to be measured).  TEST INFRASTRUCTURE, build container only.

Runs the UNMODIFIED reference (behind oracle/ref_shim.py, CPU, fp64) as `PicardRunner.run_one` does
(picard/picard_iteration.py:238-299), minus Lightning: for i = 1..N

    labels  <- OnlineDataGenerator(equation, u_{i-1}).sample_with_gradients(DATA_SIZE)        picard/data.py:211-223
    u_i     <- N_EPOCHS passes of Adam over unshuffled mini-batches, loss =
               PicardSolutionGradientWrapper.training_step                                     picard/solution_jac.py:167-213
    metrics <- rRMSE / rRMSEg exactly as EvalCallback computes them                            picard/utils.py:362-452

on a reduced config that a CPU finishes in minutes (Cha nx=100, MLP [101,128,128,128,128,1], M_T = M_I = 1024,
DATA_SIZE 1024, batch 128, 16 epochs, 8 Picard iterations), for several seeds.  `tools/spread_repo.py` runs this
repository's `run_picard` on the same config on a GPU; `profiles/r02_spread.md` puts the two side by side.

    python oracle/measure_spread.py [--seeds 0 1 2] [--iterations 8] [--out profiles/r02_spread_reference.json]
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

CONFIG = dict(nx=100, k=5.0, alpha=1.0, T=1.0, neurons=[128] * 4, M=1024, data_size=1024, batch_size=128, n_epochs=16,
              fixed_weight=1.0, loss_beta=0.0, eval_points=10000)


def eval_metrics(eq, net, n_points):
    """EvalCallback.on_validation_epoch_start with TEST_GRAD (picard/utils.py:362-452): t = linspace(0, T, n),
    x = equation.sample_x(t), rRMSE over all points, rRMSEg per dimension then averaged."""
    t = torch.linspace(0.0, eq.T, n_points).reshape(n_points, 1)
    x = eq.sample_x(t)
    tx = torch.cat([t, x], -1).requires_grad_()
    u = net(tx)
    (du,) = torch.autograd.grad(u.sum(), tx)
    u_ex, ux_ex = eq.u_u_x(t, x)
    err = (u.detach() - u_ex).numpy()
    errx = (du[:, 1:] - ux_ex).numpy()
    rr = float(np.sqrt((err ** 2).sum()) / np.sqrt((u_ex.numpy() ** 2).sum()))
    rg = float((np.sqrt((errx ** 2).sum(0)) / np.sqrt((ux_ex.numpy() ** 2).sum(0))).mean())
    return rr, rg


def run(seed: int, n_iterations: int, cfg=CONFIG, log=print):
    ref_shim.install()
    import picard.data as rdata
    import picard.equations as req
    import picard.solution as rsol
    import picard.solution_jac as rjac

    torch.set_default_dtype(torch.float64)  # DATA.FLOAT: double in every BASELINE config
    torch.manual_seed(seed)
    np.random.seed(seed)
    eq = req.Cha(nx=cfg["nx"], alpha=cfg["alpha"], k=cfg["k"], T=cfg["T"])
    C = ref_shim.cfgnode
    ncfg = C(TYPE="Value", PISGRADNET=False, NEURONS=list(cfg["neurons"]), ACTIVATIONS=["ELU"] * len(cfg["neurons"]), BOUND=None)
    tcfg = C(NUM_HESS_SAMPLES=0,
             LOSS=C(beta=cfg["loss_beta"], SCALER=C(cls="FixedLossScaler", kwargs=C(fixed_weight=cfg["fixed_weight"])),
                    FN=C(cls=None, kwargs=C()), use_aux_loss=False, weight_aux_loss=0.1),
             OPTIMIZER=C(cls="Adam", kwargs=C(), SCHEDULER=C(cls=None, kwargs=C(), config=C())))
    sol = rsol.PicardSolution(eq, ncfg, tcfg)
    wrap = rjac.PicardSolutionGradientWrapper(sol)
    frozen = rsol.ZeroSolution()
    hist = []
    t0 = time.time()
    for i in range(1, n_iterations + 1):
        gen = rdata.OnlineDataGenerator(eq, frozen, n_iterations, i, device="cpu", t_always_uniform=True,
                                        n_estimate_terminal=cfg["M"], n_estimate_integral=cfg["M"],
                                        hessian_approximation=C(method=None, kwargs=C()), sample_bound=None,
                                        estimate_terminal="OU_ByGx", estimate_integral="OU_Element", estimate_delta_t=0.0)
        xs, ys = [], []
        for _ in range(cfg["data_size"] // 256):  # the reference's buffer chunking: keeps [n*M, 1+nx] tensors small
            tx, y = gen.sample_with_gradients(256)
            xs.append(tx.detach())
            ys.append(y.detach())
        tx, y = torch.cat(xs), torch.cat(ys)
        for p in sol.parameters():
            p.requires_grad_(True)
        sol.train()
        opt = torch.optim.Adam(sol.parameters())  # fresh optimiser per Picard iteration (new Trainer), weights carried over
        step = 0
        for _ in range(cfg["n_epochs"]):
            for b in range(0, cfg["data_size"], cfg["batch_size"]):
                opt.zero_grad()
                loss = wrap.training_step((tx[b:b + cfg["batch_size"]], y[b:b + cfg["batch_size"]]), step)
                loss.backward()
                opt.step()
                step += 1
        frozen = copy.deepcopy(wrap)  # u_i, frozen by the next generator (data.py:407-412)
        rr, rg = eval_metrics(eq, wrap, cfg["eval_points"])
        hist.append(dict(iteration=i, loss=float(loss), rRMSE=rr, rRMSEg=rg))
        log(f"[reference seed {seed}] iter {i}: loss {float(loss):.3e} rRMSE {rr:.3e} rRMSEg {rg:.3e} ({time.time() - t0:.0f} s)")
    return hist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs="+", default=[0, 1, 2])
    ap.add_argument("--iterations", type=int, default=8)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_spread_reference.json"))
    ap.add_argument("--data-size", type=int, default=CONFIG["data_size"])
    ap.add_argument("--mc", type=int, default=CONFIG["M"])
    ap.add_argument("--batch-size", type=int, default=CONFIG["batch_size"])
    ap.add_argument("--threads", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    a = ap.parse_args()
    CONFIG.update(data_size=a.data_size, M=a.mc, batch_size=a.batch_size)
    torch.set_num_threads(a.threads)
    out = dict(config=CONFIG, iterations=a.iterations, impl="reference (unmodified picard/ behind oracle/ref_shim.py, CPU fp64)", runs={})
    for s in a.seeds:
        out["runs"][str(s)] = run(s, a.iterations)
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)
    finals = np.array([[r[-1]["rRMSE"], r[-1]["rRMSEg"]] for r in out["runs"].values()])
    out["final"] = dict(rRMSE_mean=float(finals[:, 0].mean()), rRMSE_min=float(finals[:, 0].min()), rRMSE_max=float(finals[:, 0].max()),
                        rRMSEg_mean=float(finals[:, 1].mean()), rRMSEg_min=float(finals[:, 1].min()), rRMSEg_max=float(finals[:, 1].max()))
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["final"]))


if __name__ == "__main__":
    main()
