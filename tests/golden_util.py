"""Load `tests/golden/*.npz` (written by oracle/make_golden.py from the unmodified reference) into oracle objects."""
import glob
import os

import numpy as np
import torch

from oracle import dpi_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    """Label-generator fixtures (the `train_*` files belong to the regression step, tests/test_train_oracle.py)."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if not n.startswith("train_")]


def _t(a, dtype=torch.float64):
    return torch.from_numpy(np.asarray(a)).to(dtype)


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.z = z
        self.eq_name = str(z["eq_name"])
        self.nx, self.M_T, self.M_I, self.v, self.n = (int(z[k]) for k in ("nx", "M_T", "M_I", "v", "n"))
        self.alpha, self.T, self.k = float(z["alpha"]), float(z["T"]), float(z["k"])
        self.delta_t = float(z["delta_t"]) if "delta_t" in z else 0.0  # DATA.ESTIMATE_DELTA_T of the generating run
        if self.eq_name == "Cha":
            self.eq = orc.Cha(self.nx, self.alpha, self.k, self.T)
        elif self.eq_name == "GBMEquationComplexExact":
            self.eq = orc.GBM(self.nx, _t(z["gbm_w"]), _t(z["gbm_v"]), self.alpha, self.T)
        else:
            self.eq = orc.OU(self.nx, _t(z["ou_mean"]), _t(z["ou_diagvar"]), _t(z["ou_pi"]), T=self.T, alpha=self.alpha)
        kind = str(z["net_kind"])
        if kind == "zero":
            self.net = orc.ZeroNet()
        elif kind == "mlp":
            L = int(z["n_layers"])
            self.net = orc.MLP([_t(z[f"W{i}"]) for i in range(L)], [_t(z[f"b{i}"]) for i in range(L)])
        else:
            def lins(nm):
                return [(_t(z[f"{nm}_W{i}"]), _t(z[f"{nm}_b{i}"])) for i in range(int(z[f"{nm}_n"]))]
            self.net = orc.PISGradNet(_t(z["phase"]), _t(z["coeff"]), lins("t_encoder"), lins("smooth_net"),
                                      lins("nn_module"), self.T)
        self.draws = orc.Draws(
            _t(z["u_t"]), _t(z["z_x0"]) if "z_x0" in z else None, _t(z["z_x"]), _t(z["zT"]), _t(z["u_s"]), _t(z["zS"]),
            torch.from_numpy(z["idx"].astype(np.int64)) if "idx" in z else None)
        self.tx, self.y = _t(z["tx"]), _t(z["y"])
        self.y_terminal, self.y_integral = _t(z["y_terminal"]), _t(z["y_integral"])
