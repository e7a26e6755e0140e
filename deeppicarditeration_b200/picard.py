"""Host-side mirror of the reference's outer loop for the MLP configs (`PicardRunner.run` / `run_one`,
picard/picard_iteration.py:238-311), on this library's two GPU stages:

    for i = 1..N:   labels (tx, y) <- OnlineDataGenerator(u_{i-1}).sample_with_gradients(DATA_SIZE)     [label kernels]
                    u_i            <- N_EPOCHS passes of Adam over unshuffled mini-batches of (tx, y)   [regression step]

`u_0` is `ZeroSolution` (picard_iteration.py:182), the network is re-loaded from the previous iterate
(`NETWORK.RELOAD`, picard_iteration.py:249-251) and every iteration gets a fresh optimiser (a new `pl.Trainer` per
iteration, picard_iteration.py:279-289).  The evaluation follows `EvalCallback` (picard/utils.py:362-458):
rRMSE = ||u - u*||_2 / ||u*||_2 on points drawn like the training points, and the same for the gradient.
Lightning, loggers, checkpoints and the CLI are out of scope.
"""
from __future__ import annotations

import time
from typing import Callable, Dict, List, Optional, Sequence

import torch

from .data import OnlineDataGenerator
from .solution import ZeroSolution, construct_mlp
from .trainer import RegressionTrainer


class _HA:
    def __init__(self, method=None, v=0):
        self.method, self.kwargs = method, {"v": v}


def evaluate(generator: OnlineDataGenerator, equation, n_points: int, seed: int = 12345) -> Dict[str, float]:
    """rRMSE of u and of grad_x u against the exact solution, the reference's EvalCallback metrics on its points
    (t = linspace(0, T, n), x = sample_x(t); picard/utils.py:362-452) -- see evaluation.py."""
    from .evaluation import EvalCallback
    cb = EvalCallback({"L2_N_POINTS": n_points, "TEST_GRAD": True}, equation, seed=seed)
    return cb(generator.engine)


def run_picard(equation, neurons: Sequence[int], n_iterations: int, *, data_size: int = 4096, n_mc: int = 4096,
               batch_size: int = 512, n_epochs: int = 16, fixed_weight: float = 1.0, loss_beta: float = 0.0,
               lr: float = 1e-3, precision: str = "f16", eval_points: int = 10000, eval_freq: int = 1, seed: int = 0,
               device="cuda", log: Optional[Callable[[str], None]] = print) -> Dict[str, object]:
    """Solve the PDE by Deep Picard Iteration with a plain MLP; returns the trained network and the error history."""
    torch.manual_seed(seed)
    dev = torch.device(device)
    nx = equation.nx
    model = construct_mlp(1 + nx, 1, list(neurons), ["ELU"] * len(neurons), None).float().to(dev)
    frozen = ZeroSolution(1)
    history: List[Dict[str, float]] = []
    gen = None
    trainer = RegressionTrainer(model, lr=lr, loss_beta=loss_beta, fixed_weight=fixed_weight, max_batch=batch_size,
                                device=dev)
    t_label = t_train = 0.0
    for i in range(1, n_iterations + 1):
        if gen is None:
            gen = OnlineDataGenerator(equation, frozen, n_iterations, i, device=dev, t_always_uniform=True,
                                      n_estimate_terminal=n_mc, n_estimate_integral=n_mc, hessian_approximation=_HA(),
                                      precision=precision, seed=seed + 1000 * i)
        else:
            gen.i, gen.seed = i, seed + 1000 * i
            gen.set_solution(frozen)  # packs a private copy of the weights: u_{i-1} stays frozen while u_i trains
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        tx, y = gen.sample_with_gradients(data_size)
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        trainer.reset_optimizer()  # fresh Adam state per iteration, weights carried over
        losses = trainer.fit(tx, y, batch_size, n_epochs)
        torch.cuda.synchronize(dev)
        t2 = time.perf_counter()
        t_label += t1 - t0
        t_train += t2 - t1
        frozen = model
        if eval_freq and (i % eval_freq == 0 or i == n_iterations):
            gen.set_solution(model)
            m = evaluate(gen, equation, eval_points)
            m.update(iteration=i, loss=float(losses[0]), label_ms=1e3 * (t1 - t0), train_ms=1e3 * (t2 - t1))
            history.append(m)
            if log:
                log(f"picard iter {i:3d}: loss {m['loss']:.3e}  rRMSE {m['rRMSE']:.3e}  rRMSEg {m['rRMSEg']:.3e}  "
                    f"labels {m['label_ms']:.1f} ms  regression {m['train_ms']:.1f} ms")
    trainer.close()
    if gen is not None:
        gen.engine.close()
    return {"model": model, "history": history, "label_s": t_label, "train_s": t_train}
