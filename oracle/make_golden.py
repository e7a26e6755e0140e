"""Generate `tests/golden/*.npz` by running the UNMODIFIED reference (via `oracle/ref_shim.py`).

TEST INFRASTRUCTURE ONLY.  Run in the build container, where `/root/reference` is mounted:

    python -m oracle.make_golden            # writes tests/golden/*.npz

The reference has no tests, no seeds and no golden vectors, so its own outputs on
*replayed draws* are the pin: every `torch.rand/randn/randn_like/rand_like/randint` call the reference
makes inside `OnlineDataGenerator.sample_with_gradients` (picard/data.py:211-223) is intercepted and
served from a pre-drawn queue (values rounded to fp32-representable numbers so that the fp32 CUDA path
can consume bit-identical inputs).  The queue order is asserted shape by shape, which also re-verifies
the consumption order documented in SURVEY.md 3.2.

Each fixture stores: equation kwargs, network weights, the draws, and the reference's `tx`, `y`,
`y_terminal` (estimate_terminal_with_gradients, data.py:899-926) and `y_integral`
(estimate_integral_with_gradients, data.py:471-527), all computed by the reference in fp64.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import dpi_oracle as orc  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


class Replay:
    """Serve torch RNG calls from a queue of pre-drawn tensors, asserting kind and shape."""

    def __init__(self, items):
        self.items = list(items)
        self.pos = 0

    def _pop(self, kind, shape):
        k, t = self.items[self.pos]
        self.pos += 1
        assert k == kind, f"draw #{self.pos}: reference asked {kind}{tuple(shape)}, queue has {k}{tuple(t.shape)}"
        assert tuple(t.shape) == tuple(shape), f"draw #{self.pos}: shape {tuple(shape)} vs queued {tuple(t.shape)}"
        return t.clone()

    def __enter__(self):
        self._saved = {n: getattr(torch, n) for n in ("rand", "randn", "rand_like", "randn_like", "randint")}

        def _shape(args):
            return tuple(args[0]) if len(args) == 1 and isinstance(args[0], (tuple, list, torch.Size)) else tuple(args)

        torch.rand = lambda *a, **k: self._pop("rand", _shape(a))
        torch.randn = lambda *a, **k: self._pop("randn", _shape(a))
        torch.rand_like = lambda x, **k: self._pop("rand", x.shape)
        torch.randn_like = lambda x, **k: self._pop("randn", x.shape)
        torch.randint = lambda lo, hi, size, **k: self._pop("randint", size)
        return self

    def __exit__(self, *exc):
        for n, f in self._saved.items():
            setattr(torch, n, f)
        if exc[0] is None:
            assert self.pos == len(self.items), f"reference consumed {self.pos} of {len(self.items)} draws"


def _queue(d: orc.Draws, part="all"):
    q = []
    if part == "all":
        q.append(("rand", d.u_t))
        if d.z_x0 is not None:
            q.append(("randn", d.z_x0))
        q.append(("randn", d.z_x))
    if part in ("all", "terminal"):
        q.append(("randn", d.zT))
    if part in ("all", "integral"):
        q.append(("rand", d.u_s))
        q.append(("randn", d.zS))
        if d.idx is not None:
            q.append(("randint", d.idx))
    return q


def _ref_mlp(net: orc.MLP, sol):
    dims = [net.weights[0].shape[1]] + [w.shape[0] for w in net.weights]
    m = sol.construct_mlp(dims[0], dims[-1], dims[1:-1], ["ELU"] * (len(dims) - 2), None)
    lin = [l for l in m if isinstance(l, torch.nn.Linear)]
    with torch.no_grad():
        for l, W, b in zip(lin, net.weights, net.biases):
            l.weight.copy_(W)
            l.bias.copy_(b)
    return m


def _ref_pisgrad(net: orc.PISGradNet, sol, hidden, nx, g0):
    m = sol.PISGradNet(hidden_shapes=list(hidden), dim=nx, g0=g0, T=net.T)
    with torch.no_grad():
        m.timestep_phase.copy_(net.phase)
        m.timestep_coeff.copy_(net.coeff)  # fp32-representable copy of linspace(0.1,100,64)
        for seq, lins in ((m.t_encoder, net.t_encoder), (m.smooth_net, net.smooth_net), (m.nn_module, net.nn_module)):
            mods = [l for l in seq if isinstance(l, torch.nn.Linear)]
            assert len(mods) == len(lins)
            for l, (W, b) in zip(mods, lins):
                l.weight.copy_(W)
                l.bias.copy_(b)
    return m


def _net_arrays(net):
    out = {"net_kind": np.array(net.kind)}
    if net.kind == "mlp":
        for i, (W, b) in enumerate(zip(net.weights, net.biases)):
            out[f"W{i}"] = W.numpy().astype(np.float32)
            out[f"b{i}"] = b.numpy().astype(np.float32)
        out["n_layers"] = np.array(len(net.weights))
    elif net.kind == "pisgrad":
        out["phase"] = net.phase.numpy().astype(np.float32)
        out["coeff"] = net.coeff.numpy().astype(np.float32)
        for name in ("t_encoder", "smooth_net", "nn_module"):
            lins = getattr(net, name)
            out[f"{name}_n"] = np.array(len(lins))
            for i, (W, b) in enumerate(lins):
                out[f"{name}_W{i}"] = W.numpy().astype(np.float32)
                out[f"{name}_b{i}"] = b.numpy().astype(np.float32)
    return out


def make_case(name, eq_name, eq_kwargs, net_spec, n, M_T, M_I, v=None, seed=0, delta_t=0.0):
    ref_shim.install()
    import picard.data as rdata
    import picard.equations as reqs
    import picard.solution as rsol

    torch.set_default_dtype(torch.float64)
    nx = eq_kwargs["nx"]
    scratch = tempfile.mkdtemp(prefix="dpi_golden_")
    extra = {}
    try:
        # ---- problem instance (the reference loads .pt files CWD-relative) ------------------
        if eq_name == "GBMEquationComplexExact":
            src = os.path.join(ref_shim.REFERENCE_ROOT, "scripts", "fully_nonlinear", "case_1")
            for f in ("gbm_2nodes_w_100d.pt", "gbm_2nodes_v_100d.pt"):
                shutil.copy(os.path.join(src, f), scratch)
        if eq_name == "OUProcessEquation":
            src = os.path.join(ref_shim.REFERENCE_ROOT, "scripts", "hjb")
            for f in ("mean_100d_ms=1.0_vs=2.0_5.pt", "pi_100d_ms=1.0_vs=2.0_5.pt"):
                shutil.copy(os.path.join(src, f), scratch)
            # `var_...pt` is not shipped; it is deterministic by construction (var_scale * I)
            torch.save(torch.stack([2.0 * torch.eye(nx) for _ in range(5)]),
                       os.path.join(scratch, "var_100d_ms=1.0_vs=2.0_5.pt"))
        with ref_shim.chdir(scratch):
            req = getattr(reqs, eq_name)(**eq_kwargs)
        req = req.to(device="cpu")

        # ---- oracle-side equation (for draw shapes) and the shared network weights -----------
        if eq_name == "Cha":
            oeq = orc.Cha(**eq_kwargs)
        elif eq_name == "GBMEquationComplexExact":
            oeq = orc.GBM(nx, req.w.double(), req.v.double(), eq_kwargs["alpha"], eq_kwargs["T"])
            extra.update(gbm_w=req.w.double().numpy(), gbm_v=req.v.double().numpy())
        else:
            diagvar = torch.diagonal(req.var, dim1=1, dim2=2).double()
            oeq = orc.OU(nx, req.mean.double(), diagvar, req.pi.double(), T=eq_kwargs["T"], alpha=eq_kwargs["alpha"],
                         alpha_scale=eq_kwargs.get("alpha_scale", 4.0))
            extra.update(ou_mean=req.mean.double().numpy(), ou_diagvar=diagvar.numpy(), ou_pi=req.pi.double().numpy())

        kind = net_spec["kind"]
        if kind == "zero":
            onet, rnet = orc.ZeroNet(), rsol.ZeroSolution()
        elif kind == "mlp":
            onet = orc.make_mlp(1 + nx, net_spec["hidden"], 1, seed=seed + 1, gain=net_spec.get("gain", 1.0))
            rnet = _ref_mlp(onet, rsol)
        else:
            onet = orc.make_pisgrad(nx, net_spec["hidden"], T=eq_kwargs["T"], seed=seed + 1, phase_scale=0.3)
            rnet = _ref_pisgrad(onet, rsol, net_spec["hidden"], nx, req.g)

        hess = ref_shim.cfgnode({"method": "SDGD" if v else None, "kwargs": {"v": v} if v else {}})
        gen = rdata.OnlineDataGenerator(
            req, rnet, N=40, i=2, device="cpu", t_always_uniform=True, n_estimate_terminal=M_T,
            n_estimate_integral=M_I, hessian_approximation=hess, sample_bound=None,
            estimate_terminal="OU_ByGx", estimate_integral="OU_Simple", estimate_delta_t=delta_t)

        g = torch.Generator().manual_seed(seed)
        d = orc.draw(oeq, n, M_T, M_I, v, g, round_fp32=True)
        with Replay(_queue(d, "all")):
            tx, y = gen.sample_with_gradients(n)
        tx = tx.detach()
        with Replay(_queue(d, "terminal")):
            yT = (gen.estimate_terminal_with_gradients_td(tx.clone()) if delta_t > 0
                  else gen.estimate_terminal_with_gradients(tx.clone())).detach()
        with Replay(_queue(d, "integral")):
            yI = (gen.estimate_integral_with_gradients_td(tx.clone()) if delta_t > 0
                  else gen.estimate_integral_with_gradients(tx.clone())).detach()
        assert torch.allclose(yT + yI, y, rtol=0, atol=1e-12)

        arrays = dict(
            eq_name=np.array(eq_name), nx=np.array(nx), alpha=np.array(eq_kwargs["alpha"]), T=np.array(eq_kwargs["T"]),
            k=np.array(eq_kwargs.get("k", 0.0)), M_T=np.array(M_T), M_I=np.array(M_I), v=np.array(v or 0), n=np.array(n),
            delta_t=np.array(float(delta_t)),
            u_t=d.u_t.numpy().astype(np.float32), z_x=d.z_x.numpy().astype(np.float32),
            zT=d.zT.numpy().astype(np.float32), u_s=d.u_s.numpy().astype(np.float32),
            zS=d.zS.numpy().astype(np.float32),
            tx=tx.numpy(), y=y.detach().numpy(), y_terminal=yT.numpy(), y_integral=yI.numpy(),
            **extra, **_net_arrays(onet))
        if d.z_x0 is not None:
            arrays["z_x0"] = d.z_x0.numpy().astype(np.float32)
        if d.idx is not None:
            arrays["idx"] = d.idx.numpy().astype(np.int16)
        # the rounding to fp32 must have been lossless for everything stored as float32
        for key in ("u_t", "z_x", "zT", "u_s", "zS"):
            assert np.array_equal(arrays[key].astype(np.float64), getattr(d, key).numpy())
        os.makedirs(GOLDEN_DIR, exist_ok=True)
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: tx{tuple(tx.shape)} |y|={float(y.norm()):.4f} |yI|/|y|={float(yI.norm() / y.norm()):.3e}"
              f" -> {os.path.getsize(path) / 1024:.0f} KiB")
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


CASES = [
    dict(name="burgers_100d_mlp128", eq_name="Cha", eq_kwargs=dict(nx=100, alpha=1.0, k=5.0, T=1.0),
         net_spec=dict(kind="mlp", hidden=[128] * 4), n=6, M_T=24, M_I=24, seed=10),
    dict(name="burgers_100d_zero", eq_name="Cha", eq_kwargs=dict(nx=100, alpha=1.0, k=5.0, T=1.0),
         net_spec=dict(kind="zero"), n=6, M_T=24, M_I=24, seed=11),
    dict(name="burgers_10d_ragged", eq_name="Cha", eq_kwargs=dict(nx=10, alpha=0.7, k=2.0, T=1.5),
         net_spec=dict(kind="mlp", hidden=[32, 16], gain=2.0), n=7, M_T=17, M_I=5, seed=12),
    dict(name="gbm_100d_mlp64_sdgd", eq_name="GBMEquationComplexExact", eq_kwargs=dict(nx=100, alpha=1.0, T=1.0),
         net_spec=dict(kind="mlp", hidden=[64] * 3, gain=1.5), n=4, M_T=12, M_I=12, v=100, seed=13),
    dict(name="gbm_100d_zero_sdgd", eq_name="GBMEquationComplexExact", eq_kwargs=dict(nx=100, alpha=1.0, T=1.0),
         net_spec=dict(kind="zero"), n=4, M_T=12, M_I=12, v=100, seed=14),
    dict(name="hjb_100d_pisgrad", eq_name="OUProcessEquation",
         eq_kwargs=dict(nx=100, alpha=1.0, T=1.0, num_components=5, mean_scale=1.0, var_scale=2.0, alpha_scale=4.0),
         net_spec=dict(kind="pisgrad", hidden=[96, 64, 80]), n=4, M_T=12, M_I=12, seed=15),
    dict(name="hjb_100d_zero", eq_name="OUProcessEquation",
         eq_kwargs=dict(nx=100, alpha=1.0, T=1.0, num_components=5, mean_scale=1.0, var_scale=2.0, alpha_scale=4.0),
         net_spec=dict(kind="zero"), n=4, M_T=12, M_I=12, seed=16),
    # temporal-difference estimators (DATA.ESTIMATE_DELTA_T = 0.3: about a third of the points end their step at T)
    dict(name="burgers_100d_mlp128_td", eq_name="Cha", eq_kwargs=dict(nx=100, alpha=1.0, k=5.0, T=1.0),
         net_spec=dict(kind="mlp", hidden=[128] * 4), n=8, M_T=24, M_I=24, seed=17, delta_t=0.3),
    dict(name="burgers_10d_zero_td", eq_name="Cha", eq_kwargs=dict(nx=10, alpha=0.7, k=2.0, T=1.5),
         net_spec=dict(kind="zero"), n=8, M_T=17, M_I=5, seed=18, delta_t=0.4),
    dict(name="hjb_100d_pisgrad_td", eq_name="OUProcessEquation",
         eq_kwargs=dict(nx=100, alpha=1.0, T=1.0, num_components=5, mean_scale=1.0, var_scale=2.0, alpha_scale=4.0),
         net_spec=dict(kind="pisgrad", hidden=[96, 64, 80]), n=6, M_T=12, M_I=12, seed=19, delta_t=0.3),
]


def main():
    if not ref_shim.reference_available():
        raise SystemExit("reference not mounted; fixtures can only be generated in the build container")
    only = set(sys.argv[1:])
    for c in CASES:
        if not only or c["name"] in only:
            make_case(**c)


if __name__ == "__main__":
    main()
