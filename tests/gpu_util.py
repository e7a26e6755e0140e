"""Helpers for the `-m gpu` parity tests: build a LabelEngine from a golden fixture or oracle objects."""
import torch

from deeppicarditeration_b200.engine import LabelEngine


def engine_for(eq, net, sdgd_v=0, device="cuda"):
    """eq / net are oracle objects (oracle.dpi_oracle) used here only to carry parameters."""
    e = LabelEngine(device)
    if eq.name == "Cha":
        import math
        e.set_problem("Cha", eq.nx, eq.T, eq.alpha, cha_k=eq.k * math.sqrt(eq.nx))
    elif eq.name == "GBM":
        e.set_problem("GBM", eq.nx, eq.T, eq.alpha, gbm_w=eq.w.numpy(), gbm_v=eq.v.numpy(), sdgd_v=sdgd_v)
    else:
        e.set_problem("OU", eq.nx, eq.T, eq.alpha, ou_mean=eq.mean.numpy(), ou_diagvar=eq.diagvar.numpy(),
                      ou_pi=eq.pi.numpy(), ou_theta=eq.theta, ou_mu=eq.mu, ou_alpha_init=eq.alpha_init)
    if net.kind == "zero":
        e.set_network("zero")
    elif net.kind == "mlp":
        e.set_network("mlp", main=list(zip(net.weights, net.biases)))
    else:
        e.set_network("pisgrad", main=net.nn_module, t_encoder=net.t_encoder, smooth_net=net.smooth_net,
                      phase=net.phase, coeff=net.coeff)
    return e


def draws_dict(d):
    return {"zT": d.zT, "u_s": d.u_s, "zS": d.zS, "idx": d.idx}
