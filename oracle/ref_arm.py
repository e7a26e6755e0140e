"""CPU arm of `bench.py` on the reference's OWN code: `OnlineDataGenerator.sample_with_gradients` (picard/data.py:211-223)
of the UNMODIFIED reference, imported behind `oracle/ref_shim.py` from `/root/reference` (build container) or from the
byte-identical staged copy `oracle/_ref/` (GPU box; `oracle/build_ref.py`).

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- only `bench.py`'s CPU legs and `tests/` import this file.

The generator is configured the way `PicardRunner` configures it from the BASELINE YAMLs (picard/picard_iteration.py:
160-215; `DATA.FLOAT: double`, `ESTIMATE_TERMINAL: OU_ByGx`, `ESTIMATE_INTEGRAL: OU_Simple`, `t_always_uniform`), on the
same network weights and the same shipped problem instance as the GPU arm (the `.pt` instance files the reference loads
CWD-relative, equations.py:408-411,527-532, are rewritten from the arrays carried by `tests/golden/*.npz`).
"""
from __future__ import annotations

import contextlib
import os
import shutil
import sys
import tempfile
import time

import torch

from oracle import ref_shim


def available() -> bool:
    return ref_shim.reference_available()


def _set_linears(seq, lins):
    mods = [m for m in seq if isinstance(m, torch.nn.Linear)]
    assert len(mods) == len(lins), (len(mods), len(lins))
    with torch.no_grad():
        for m, (W, b) in zip(mods, lins):
            assert tuple(m.weight.shape) == tuple(W.shape), (tuple(m.weight.shape), tuple(W.shape))
            m.weight.copy_(W)
            m.bias.copy_(b)


def build_generator(wl: dict, net: dict, inst: dict, M: int):
    """Reference equation + frozen network + OnlineDataGenerator for one bench workload (fp64, CPU)."""
    ref_shim.install()
    import picard.data as rdata
    import picard.equations as reqs
    import picard.solution as rsol

    torch.set_default_dtype(torch.float64)
    nx = wl["nx"]
    scratch = tempfile.mkdtemp(prefix="dpi_ref_arm_")
    try:
        if wl["eq"] == "Cha":
            eq = reqs.Cha(nx=nx, alpha=1.0, k=wl["cha_k"], T=1.0)
        elif wl["eq"] == "GBM":
            torch.save(torch.from_numpy(inst["gbm_w"]).double(), os.path.join(scratch, f"gbm_2nodes_w_{nx}d.pt"))
            torch.save(torch.from_numpy(inst["gbm_v"]).double(), os.path.join(scratch, f"gbm_2nodes_v_{nx}d.pt"))
            with ref_shim.chdir(scratch):
                eq = reqs.GBMEquationComplexExact(nx=nx, alpha=1.0, T=1.0)
        else:
            K = int(inst["ou_pi"].shape[0])
            tag = f"{nx}d_ms=1.0_vs=2.0_{K}.pt"
            torch.save(torch.from_numpy(inst["ou_mean"]).double(), os.path.join(scratch, "mean_" + tag))
            torch.save(torch.from_numpy(inst["ou_pi"]).double(), os.path.join(scratch, "pi_" + tag))
            torch.save(torch.stack([torch.diag(torch.from_numpy(inst["ou_diagvar"][k]).double()) for k in range(K)]),
                       os.path.join(scratch, "var_" + tag))
            with ref_shim.chdir(scratch):
                eq = reqs.OUProcessEquation(nx=nx, alpha=1.0, T=1.0, num_components=K, mean_scale=1.0, var_scale=2.0,
                                            alpha_scale=4.0)
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    eq = eq.to(device="cpu")
    hid = list(wl["hidden"])
    dd = lambda lins: [(W.double(), b.double()) for W, b in lins]
    if net["kind"] == "mlp":
        rnet = rsol.construct_mlp(1 + nx, 1, hid, ["ELU"] * len(hid), None)
        _set_linears(rnet, dd(net["main"]))
    else:
        rnet = rsol.PISGradNet(hidden_shapes=hid, dim=nx, g0=eq.g, T=1.0)
        with torch.no_grad():
            rnet.timestep_phase.copy_(net["phase"].double())
            rnet.timestep_coeff.copy_(net["coeff"].double())
        _set_linears(rnet.t_encoder, dd(net["t_encoder"]))
        _set_linears(rnet.smooth_net, dd(net["smooth_net"]))
        _set_linears(rnet.nn_module, dd(net["main"]))
    v = wl.get("v") or 0
    hess = ref_shim.cfgnode({"method": "SDGD" if v else None, "kwargs": {"v": v} if v else {}})
    return rdata.OnlineDataGenerator(eq, rnet, N=40, i=2, device="cpu", t_always_uniform=True, n_estimate_terminal=M,
                                     n_estimate_integral=M, hessian_approximation=hess, sample_bound=None,
                                     estimate_terminal="OU_ByGx", estimate_integral="OU_Simple", estimate_delta_t=0.0)


def time_reference(wl: dict, net: dict, inst: dict, M: int, n_sample: int, steps: int, warmup: int, threads: int) -> dict:
    """`steps` timed calls of the reference's sample_with_gradients(n_sample) after `warmup` untimed ones."""
    torch.set_num_threads(threads)
    prev = torch.get_default_dtype()
    try:
        with contextlib.redirect_stdout(sys.stderr):  # the reference prints while constructing; stdout is bench.py's JSON line
            gen = build_generator(wl, net, inst, M)
        torch.manual_seed(0)
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            tx, y = gen.sample_with_gradients(n_sample)
            dt = time.perf_counter() - t0
            assert tuple(y.shape) == (n_sample, 1 + wl["nx"]) and bool(torch.isfinite(y).all())
            if i >= warmup:
                times.append(dt)
    finally:
        torch.set_default_dtype(prev)
    tot = sum(times)
    return dict(value=n_sample * len(times) / tot, ms_per_step=1e3 * tot / len(times), cores=torch.get_num_threads(),
                steps=len(times))
