"""Developer helper: the three BASELINE configs solved end to end through `picard train` on this repository's kernels
(labels -> regression -> evaluation callback), with the shipped problem instances taken from tests/golden.

    python tools/solve_configs.py [burgers hjb gbm] [--iterations N]     -> gpurun_out/r02_solve_<name>.json

The YAML texts below restate the values of scripts/burgers/base_100d_T1.0_w1.0_0.yaml, scripts/hjb/base_100d_T1.0_w0.1_0.yaml
and scripts/fully_nonlinear/case_1/base_100d_T1.0_w0.0_nov_0.yaml (the reference tree does not exist on the GPU box)."""
import json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

COMMON = """
METHOD: {cls: Picard}
FORCE: true
"""
CFG = {
    "burgers": """
NAME: burgers100d
EQUATION: {cls: Cha, kwargs: {nx: 100, alpha: 1.0, k: 5.0, T: 1.0}}
PICARD: {N: 100}
DATA:
  FLOAT: double
  DATA_SIZE: 4096
  PREFETCH_FACTOR: None
  kwargs: {t_always_uniform: true, n_estimate_terminal: 4096, n_estimate_integral: 4096}
TRAIN:
  N_EPOCHS: 16
  BATCH_SIZE: 512
  SUPERVISE_GRADIENT: true
  LOSS: {beta: 0.0, SCALER: {cls: FixedLossScaler, kwargs: {fixed_weight: 1.0}}}
NETWORK: {NEURONS: [128, 128, 128, 128], ACTIVATIONS: [ELU, ELU, ELU, ELU], BOUND: None, RELOAD: true}
EVAL: {L2_N_POINTS: 10000, FREQ: 8, TEST_GRAD: true}
""",
    "hjb": """
NAME: hjb100d
EQUATION:
  cls: OUProcessEquation
  kwargs: {nx: 100, alpha: 1.0, T: 1.0, num_components: 5, mean_scale: 1.0, var_scale: 2.0, alpha_scale: 4.0}
PICARD: {N: 40}
DATA:
  FLOAT: double
  DATA_SIZE: 4096
  PREFETCH_FACTOR: None
  kwargs: {t_always_uniform: true, n_estimate_terminal: 4096, n_estimate_integral: 4096}
TRAIN:
  N_EPOCHS: 16
  BATCH_SIZE: 512
  SUPERVISE_GRADIENT: true
  LOSS: {beta: 0.0, SCALER: {cls: FixedLossScaler, kwargs: {fixed_weight: 0.1}}}
  OPTIMIZER: {kwargs: {lr: 0.001}}
NETWORK: {NEURONS: [512, 512, 512, 512], ACTIVATIONS: [ELU, ELU, ELU, ELU], BOUND: None, RELOAD: true, PISGRADNET: true}
EVAL: {L2_N_POINTS: 100, FREQ: 8, TEST_GRAD: true}
""",
    "gbm": """
NAME: gbm100d
EQUATION: {cls: GBMEquationComplexExact, kwargs: {nx: 100, alpha: 1.0, T: 1.0}}
PICARD: {N: 40}
DATA:
  FLOAT: double
  DATA_SIZE: 2048
  PREFETCH_FACTOR: None
  kwargs: {t_always_uniform: true, n_estimate_terminal: 1024, n_estimate_integral: 1024}
  HESSIAN_APPROXIMATION: {method: SDGD, kwargs: {v: 100}}
TRAIN:
  N_EPOCHS: 16
  BATCH_SIZE: 512
  SUPERVISE_GRADIENT: true
  SUPERVISE_HESSIAN: false
  LOSS: {beta: 0.0, SCALER: {cls: FixedLossScaler, kwargs: {fixed_weight: 0.0}}}
NETWORK: {NEURONS: [64, 64, 64], ACTIVATIONS: [ELU, ELU, ELU], BOUND: None, RELOAD: true}
EVAL: {L2_N_POINTS: 1000, FREQ: 4, TEST_GRAD: true, TEST_HESSIAN: true}
""",
}


def instance_files(name, d):
    """The shipped problem instances (scripts/**.pt), carried by the reference fixtures under tests/golden."""
    g = os.path.join(ROOT, "tests", "golden")
    if name == "gbm":
        z = np.load(os.path.join(g, "gbm_100d_mlp64_sdgd.npz"))
        torch.save(torch.from_numpy(z["gbm_w"]), os.path.join(d, "gbm_2nodes_w_100d.pt"))
        torch.save(torch.from_numpy(z["gbm_v"]), os.path.join(d, "gbm_2nodes_v_100d.pt"))
    if name == "hjb":
        z = np.load(os.path.join(g, "hjb_100d_pisgrad.npz"))
        tag = "100d_ms=1.0_vs=2.0_5.pt"
        torch.save(torch.from_numpy(z["ou_mean"]), os.path.join(d, "mean_" + tag))
        torch.save(torch.from_numpy(z["ou_pi"]), os.path.join(d, "pi_" + tag))
        torch.save(torch.stack([torch.diag(torch.from_numpy(z["ou_diagvar"][k])) for k in range(5)]), os.path.join(d, "var_" + tag))


def main():
    n_it = int(sys.argv[sys.argv.index("--iterations") + 1]) if "--iterations" in sys.argv else None
    names = [a for a in sys.argv[1:] if a in CFG] or ["burgers", "hjb", "gbm"]
    from deeppicarditeration_b200.main import train
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    for name in names:
        with tempfile.TemporaryDirectory() as d:
            instance_files(name, d)
            with open(os.path.join(d, "cfg.yaml"), "w") as f:
                f.write(CFG[name] + COMMON)
            cwd = os.getcwd()
            os.chdir(d)
            t0 = time.time()
            try:
                hist = train("cfg.yaml", ["PICARD.N", str(n_it)] if n_it else [])
            finally:
                os.chdir(cwd)
                torch.set_default_dtype(torch.float32)
            wall = time.time() - t0
        keys = ("iteration", "loss", "rRMSE", "rRMSEg", "rRMSEh", "label_ms", "train_ms")
        rows = [{k: h[k] for k in keys if k in h} for h in hist]
        res = dict(config=name, iterations=len(hist), wall_s=wall, label_s=sum(h["label_ms"] for h in hist) / 1e3,
                   train_s=sum(h["train_ms"] for h in hist) / 1e3, final=rows[-1],
                   evaluated=[r for r in rows if "rRMSE" in r])
        json.dump(res, open(os.path.join(out_dir, f"r02_solve_{name}.json"), "w"), indent=1)
        print(name, json.dumps({k: res[k] for k in ("iterations", "wall_s", "label_s", "train_s", "final")}))


if __name__ == "__main__":
    main()
