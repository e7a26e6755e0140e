"""`PicardRunner`: the outer Deep-Picard-Iteration loop driven by the reference's config tree
(picard/picard_iteration.py:76-307), on this library's GPU stages.

    for i = 1..PICARD.N:
        u_i <- solution of NETWORK.* (weights reloaded from model_{i-1}.ckpt when NETWORK.RELOAD)
        labels  <- OnlineDataGenerator(equation, u_{i-1}, **DATA.kwargs, ...)     one fused call per DATA_SIZE/world points
        cache   <- LabelStore on the device (CacheToMemoryWrapper; DATA.SAVE writes split_<rank>.h5 / .npz with tx, u_ux)
        fit     <- TRAIN.N_EPOCHS passes of the regression step (RegressionTrainer: value + gradient matching loss, Adam)
        eval    <- EvalCallback (rRMSE, rRMSEg, rRMSEh) -> metrics.jsonl
        ckpt    <- exp_dir/model_{i}.ckpt  = {"state_dict": ...} with the reference's key layout

What is NOT here (out of scope): Lightning, wandb / tensorboard loggers, the GPU-memory prober that sizes
the reference's generator calls (`memory.py` -- the fused kernels have no path tensors to size), the competitor
solvers (METHOD.cls PINN / Diffusion / FullyNonlinearSolver).  Under `torchrun` every rank labels its contiguous block
of the points, gathers nothing, and the regression step all-reduces gradients (data-parallel, NCCL).
"""
from __future__ import annotations

import json
import os
import pathlib
import shutil
import time
from typing import Dict, List, Optional

import torch

from . import equations as eqs
from . import solution as sols
from .data import OnlineDataGenerator
from .dataset import CacheToMemoryWrapper
from .evaluation import EvalCallback
from .sharding import shard_range
from .trainer import AutogradRegressionTrainer, RegressionTrainer


def _dist():
    d = torch.distributed
    if d.is_available() and d.is_initialized():
        return d.get_rank(), d.get_world_size()
    return 0, 1


class PicardRunner:
    @staticmethod
    def get_model_checkpoint_path(exp_dir: pathlib.Path, i: int) -> pathlib.Path:
        return pathlib.Path(exp_dir) / f"model_{i}.ckpt"  # picard_iteration.py:78-79

    @staticmethod
    def get_config_path(exp_dir: pathlib.Path) -> pathlib.Path:
        return pathlib.Path(exp_dir) / "config.yaml"

    @staticmethod
    def get_equation(cfg):
        if not hasattr(eqs, cfg.EQUATION.cls):
            raise NotImplementedError(f"equation {cfg.EQUATION.cls} is not on the B200 label path")
        return getattr(eqs, cfg.EQUATION.cls)(**dict(cfg.EQUATION.kwargs))

    def __init__(self, cfg, precision: Optional[str] = None, seed: int = 0, log=print):
        self.cfg, self.log = cfg, log or (lambda *_: None)
        self.rank, self.world = _dist()
        if cfg.METHOD.cls != "Picard":
            raise NotImplementedError(f"METHOD.cls={cfg.METHOD.cls}: only the Picard method is built (DESIGN.md section 6)")
        if cfg.PICARD.FORMULA is not None:
            raise NotImplementedError("PICARD.FORMULA (two-layer formula, data.py:1328-1379) is not built")
        if cfg.TRAIN.SUPERVISE_HESSIAN:
            raise NotImplementedError("TRAIN.SUPERVISE_HESSIAN (Hessian labels, data.py:680-897) is not built")
        if not cfg.DATA.ONLINE or cfg.DATA.EXACT:
            raise NotImplementedError("offline / exact-solution datasets are not built (DATA.ONLINE true, DATA.EXACT false)")
        self.exp_dir = pathlib.Path(cfg.NAME)
        if self.rank == 0:
            self._prepare_exp_dir()
        if self.world > 1:
            torch.distributed.barrier()
        self.equation = self.get_equation(cfg)
        # picard_iteration.py:149-160: every shipped equation has a gradient term, so gradients are always supervised
        self.supervise_gradient = bool(cfg.TRAIN.SUPERVISE_GRADIENT or self.equation.has_gradient_term)
        if not self.supervise_gradient:
            raise NotImplementedError("value-only labels are unreachable with the shipped equations (SURVEY.md F3)")
        if cfg.NETWORK.TYPE != "Value":
            raise NotImplementedError("NETWORK.TYPE != Value is not built")
        self.N, self.i = int(cfg.PICARD.N), 0
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.precision = precision or dict(cfg.DATA.kwargs).get("precision", "f16")
        self.seed = int(seed)
        self.u_current = sols.ZeroSolution(1)  # picard_iteration.py:182
        self.u_history = [self.u_current]
        self.model = None
        self.trainer: Optional[RegressionTrainer] = None
        self.generator: Optional[OnlineDataGenerator] = None
        self.history: List[Dict[str, float]] = []
        self.label_s = self.train_s = 0.0

    def _prepare_exp_dir(self) -> None:
        """picard_iteration.py:120-145: refuse to overwrite unless FORCE; dump the full config."""
        d, cfg_file = self.exp_dir, self.get_config_path(self.exp_dir)
        if d.exists():
            content = [p for p in d.iterdir()]
            if content and not (len(content) == 1 and content[0].name == cfg_file.name):
                if not self.cfg.FORCE:
                    raise FileExistsError(f"Experiment directory {d} already exists. Please change the name in the "
                                          f"config file to avoid overwriting.")
                shutil.rmtree(d)
        d.mkdir(parents=True, exist_ok=True)
        cfg_file.write_text(self.cfg.dump())
        self.log(f"Experiment directory: {d.absolute()}")

    # ------------------------------------------------------------------------------------------------
    def _new_solution(self):
        torch.manual_seed(self.seed)  # identical initial weights on every rank
        net = sols.PicardSolution(self.equation, self.cfg.NETWORK).float().to(self.device)
        return net  # (PISGradNet trains through AutogradRegressionTrainer, the library path: trainer.py)

    def state_dict_for_ckpt(self) -> Dict[str, torch.Tensor]:
        """Key layout of the reference's checkpoints: `model.<idx>.weight` under PicardSolution, prefixed `solution.`
        under the gradient wrapper that training uses when gradients are supervised (solution_jac.py:113-146)."""
        prefix = "solution." if self.supervise_gradient else ""
        return {prefix + k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()}

    def _generator_kwargs(self) -> dict:
        """The keyword set of PicardDataModule.get_data_generator (data.py:1465-1496)."""
        D = self.cfg.DATA
        kw = dict(D.kwargs)
        kw.pop("precision", None)
        return dict(device=self.device, hessian_approximation=D.HESSIAN_APPROXIMATION, sample_bound=D.SAMPLE_BOUND,
                    estimate_terminal=D.ESTIMATE_TERMINAL, estimate_integral=D.ESTIMATE_INTEGRAL,
                    estimate_delta_t=D.ESTIMATE_DELTA_T, precision=self.precision, **kw)

    def run_one(self) -> bool:
        """picard_iteration.py:238-299."""
        cfg = self.cfg
        self.i += 1
        if self.model is None:
            self.model = self._new_solution()
            if cfg.NETWORK.PRETRAIN_PATH is not None:
                self.load_checkpoint(cfg.NETWORK.PRETRAIN_PATH)
            lr = dict(cfg.TRAIN.OPTIMIZER.kwargs).get("lr", 1e-3)
            if cfg.TRAIN.OPTIMIZER.cls != "Adam" or cfg.TRAIN.OPTIMIZER.SCHEDULER.cls is not None:
                raise NotImplementedError("only Adam without a scheduler is built (what every BASELINE config uses)")
            scaler = cfg.TRAIN.LOSS.SCALER
            if scaler.cls not in (None, "FixedLossScaler", "SimpleLossScaler", "DimensionLossScaler"):
                raise NotImplementedError(f"loss scaler {scaler.cls} is not built")
            fw = float(dict(scaler.kwargs).get("fixed_weight", 1.0)) if scaler.cls else 1.0
            if cfg.NETWORK.PISGRADNET or scaler.cls not in (None, "FixedLossScaler"):
                # what the native step does not cover runs on torch autograd (same loss, same optimiser)
                self.trainer = AutogradRegressionTrainer(self.model, lr=lr, loss_beta=float(cfg.TRAIN.LOSS.beta),
                                                         scaler=scaler.cls, fixed_weight=fw, device=self.device)
            else:
                self.trainer = RegressionTrainer(self.model, lr=lr, loss_beta=float(cfg.TRAIN.LOSS.beta), fixed_weight=fw,
                                                 max_batch=int(cfg.TRAIN.BATCH_SIZE), device=self.device)
        elif not cfg.NETWORK.RELOAD:
            # a fresh network per iteration unless RELOAD (picard_iteration.py:242,249-251): re-initialise in place
            fresh = self._new_solution()
            with torch.no_grad():
                for p, q in zip(self.model.parameters(), fresh.parameters()):
                    p.copy_(q)
        n_total = int(cfg.DATA.DATA_SIZE)
        start, count = shard_range(n_total, self.rank, self.world)
        seed_i = self.seed + 1000 * self.i
        if self.generator is None:
            self.generator = OnlineDataGenerator(self.equation, self.u_current, self.N, self.i, seed=seed_i,
                                                 **self._generator_kwargs())
        else:
            self.generator.i, self.generator.seed = self.i, seed_i
            self.generator._t_factors = 0 if self.generator.t_always_uniform else max(1, self.N - self.i + 1)
            self.generator.set_solution(self.u_current)  # packs a private frozen copy of u_{i-1}
        self.generator._base, self.generator._drawn, self.generator._limit = start, 0, max(count, 1)
        torch.cuda.synchronize(self.device)
        t0 = time.perf_counter()
        data = self.generator.dataset_with_gradients(count, 1, count)  # one fused call for the rank's block
        cache = CacheToMemoryWrapper(data, batch_size=int(cfg.TRAIN.BATCH_SIZE))
        if cfg.DATA.SAVE:  # data.py:1513-1521: split_XX files with datasets tx, u_ux
            cache.enable_save_to_file(str(self.exp_dir / f"iter_{self.i}_split_{self.rank:02d}.h5"))
        nx = self.equation.nx
        cache.init(count, [1 + nx, 1 + nx], preload=True)
        tx, y = cache.saver.data
        torch.cuda.synchronize(self.device)
        t1 = time.perf_counter()
        self.trainer.reset_optimizer()  # a new Trainer / optimiser per iteration (picard_iteration.py:279-289)
        local_batch = max(1, int(cfg.TRAIN.BATCH_SIZE) // self.world)
        losses = self.trainer.fit(tx, y, local_batch, int(cfg.TRAIN.N_EPOCHS))
        torch.cuda.synchronize(self.device)
        t2 = time.perf_counter()
        self.label_s += t1 - t0
        self.train_s += t2 - t1
        self.u_current = self.model
        metrics = {"iteration": self.i, "loss": float(losses[0]), "label_ms": 1e3 * (t1 - t0), "train_ms": 1e3 * (t2 - t1)}
        if cfg.EVAL.FREQ is not None:
            self.generator.set_solution(self.model)
            cb = EvalCallback(cfg.EVAL, self.equation, iter=self.i, seed=12345)
            metrics.update(cb(self.generator.engine))
        self.history.append(metrics)
        if self.rank == 0:
            torch.save({"state_dict": self.state_dict_for_ckpt(), "picard_iteration": self.i, "global_step": self.trainer.step_count},
                       self.get_model_checkpoint_path(self.exp_dir, self.i))
            with open(self.exp_dir / "metrics.jsonl", "a") as f:
                f.write(json.dumps(metrics) + "\n")
            self.log("picard iter %3d: " % self.i + "  ".join(
                f"{k} {v:.3e}" for k, v in metrics.items() if k in ("loss", "rRMSE", "rRMSEg", "rRMSEh"))
                + f"  labels {metrics['label_ms']:.1f} ms  regression {metrics['train_ms']:.1f} ms")
        return True

    def load_checkpoint(self, path) -> None:
        sd = torch.load(path, map_location="cpu")["state_dict"]
        sd = {(k[len("solution."):] if k.startswith("solution.") else k): v for k, v in sd.items()}
        with torch.no_grad():
            for k, p in self.model.state_dict().items():
                p.copy_(sd[k].to(p))

    def run(self):
        for _ in range(self.N):
            if not self.run_one():
                break
            self.u_history.append(self.u_current)
        return self.history

    def close(self) -> None:
        if self.trainer is not None:
            self.trainer.close()
        if self.generator is not None:
            self.generator.engine.close()
