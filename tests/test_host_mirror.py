"""CPU tests of the host-side mirror (equations / solution / dataset / sharding), of the C-ABI export table and of
the N>1 sharding logic (world_size 2, gloo).  No compute call reaches the CUDA library here."""
import os
import re
import sys

import pytest
import torch

from deeppicarditeration_b200 import _native as nat
from deeppicarditeration_b200 import dataset as ds
from deeppicarditeration_b200 import equations as eqs
from deeppicarditeration_b200 import sharding
from deeppicarditeration_b200 import solution as sol
from oracle import dpi_oracle as orc
from oracle import ref_shim
from tests.golden_util import Golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _NetCfg:
    TYPE, PISGRADNET, NEURONS, ACTIVATIONS, BOUND = "Value", False, [128] * 4, ["ELU"] * 4, None


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dpi_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int64_t|int|const char\*)\s+(dpi_\w+)\s*\(", hdr, flags=re.M))
    assert declared >= {"dpi_create", "dpi_label_gen", "dpi_label_gen_host", "dpi_sample_tx", "dpi_set_network"}
    assert declared == set(nat.EXPORTED_SYMBOLS)
    lib = nat.load()  # dlopen; raises AttributeError if a declared symbol is missing
    assert lib.dpi_version() == 101
    for name in declared:
        assert hasattr(lib, name)


def test_no_cpu_fallback_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from deeppicarditeration_b200.data import OnlineDataGenerator
    from deeppicarditeration_b200.engine import LabelEngine
    with pytest.raises(RuntimeError):
        LabelEngine("cuda")
    eq = eqs.Cha(nx=10, alpha=1.0, k=5.0)
    with pytest.raises(RuntimeError):
        OnlineDataGenerator(eq, sol.ZeroSolution(), 10, 1, device="cpu", t_always_uniform=True)
    # dpi_create itself must fail (status, message) rather than fall back
    import ctypes as C
    h = C.c_void_p()
    rc = nat.load().dpi_create(0, C.byref(h))
    assert rc == -2 and b"no CPU fallback" in nat.load().dpi_last_error()


def test_cha_mirror_matches_oracle_and_descriptor():
    g = Golden("burgers_100d_mlp128")
    e = eqs.Cha(nx=100, alpha=1.0, k=5.0, T=1.0)
    x, t = g.tx[:, 1:], g.tx[:, :1]
    assert torch.allclose(e.g(x), g.eq.g(x), rtol=1e-6, atol=1e-7)
    assert torch.allclose(e.exact_solution(t, x), g.eq.exact_solution(t, x), rtol=1e-6, atol=1e-7)
    z = torch.randn(5, 100, dtype=torch.float64)
    assert torch.allclose(e.ff(t[:5], x[:5], 0.3 * torch.ones(5, 1, dtype=torch.float64), z),
                          g.eq.ff(t[:5], x[:5], 0.3 * torch.ones(5, 1, dtype=torch.float64), z), rtol=1e-6)
    d = e.descriptor()
    assert d["kind"] == "Cha" and d["cha_k"] == 5.0 and d["nx"] == 100
    assert e.has_gradient_term and not e.has_hessian_term
    with pytest.raises(NotImplementedError):
        e.f(None, None, None)


def test_gbm_and_ou_mirror_match_oracle():
    g = Golden("gbm_100d_mlp64_sdgd")
    e = eqs.GBMEquationComplexExact(nx=100, w=g.eq.w, v=g.eq.v)
    x, t = g.tx[:, 1:], g.tx[:, :1]
    assert torch.allclose(e.g(x), g.eq.g(x), rtol=1e-10)
    uii = torch.randn(x.shape[0], 7, dtype=torch.float64)
    assert torch.allclose(e.ffi(t, x, None, uii), g.eq.ffi(t, x, None, uii), rtol=1e-9, atol=1e-12)
    assert e.supported_approximate_methods == ("SDGD",) and e.has_hessian_term
    # exact Hessian (equations.py:445-450) against autograd of the exact solution; its diagonal is hessian_diagonal
    Hx = e.u_hessian(t[:3], x[:3])
    for i in range(3):
        ref = torch.autograd.functional.hessian(lambda v: e.exact_solution(t[i:i + 1], v.unsqueeze(0)).sum(), x[i])
        assert torch.allclose(Hx[i], ref, rtol=1e-9, atol=1e-12)
    assert torch.allclose(torch.diagonal(Hx, dim1=1, dim2=2), e.hessian_diagonal(t[:3], x[:3]), rtol=1e-12)
    h = Golden("hjb_100d_zero")
    var = torch.diag_embed(h.eq.diagvar)
    o = eqs.OUProcessEquation(nx=100, T=1.0, num_components=5, mean=h.eq.mean, var=var, pi=h.eq.pi)
    xx = h.tx[:, 1:]
    assert torch.allclose(o.g(xx), h.eq.g(xx), rtol=1e-10)
    assert torch.allclose(o.g_x(xx), h.eq.g_x(xx), rtol=1e-9, atol=1e-12)
    zz = torch.randn_like(xx)
    assert torch.allclose(o.ff(None, xx, torch.zeros(xx.shape[0], 1, dtype=xx.dtype), zz),
                          h.eq.ff(None, xx, torch.zeros(xx.shape[0], 1, dtype=xx.dtype), zz), rtol=1e-10)
    d = o.descriptor()
    assert d["kind"] == "OU" and d["ou_mean"].shape == (5, 100) and abs(d["ou_alpha_init"] - 4.0) < 1e-12
    # exact solution (all rows at once) and its x-gradient by autograd against the oracle's row-by-row restatement
    tt = torch.rand(xx.shape[0], 1, dtype=xx.dtype)
    assert torch.allclose(o.exact_solution(tt, xx), h.eq.exact_solution(tt, xx), rtol=1e-10, atol=1e-12)
    xr = xx.clone().requires_grad_(True)
    (gx,) = torch.autograd.grad(h.eq.exact_solution(tt, xr).sum(), xr)
    assert torch.allclose(o.u_x(tt, xx), gx, rtol=1e-9, atol=1e-12)
    u2, ux2 = o.u_u_x(tt, xx)
    assert torch.allclose(ux2, gx, rtol=1e-9, atol=1e-12) and u2.shape == (xx.shape[0], 1)


def test_describe_solution_layouts():
    eq = eqs.Cha(nx=100, alpha=1.0, k=5.0)
    net = sol.PicardSolution(eq, _NetCfg)
    assert list(net.state_dict().keys())[:2] == ["model.0.weight", "model.0.bias"]  # reference ckpt key layout
    d = sol.describe_solution(net)
    assert d["kind"] == "mlp" and [tuple(W.shape) for W, _ in d["main"]] == [(128, 101), (128, 128), (128, 128),
                                                                              (128, 128), (1, 128)]
    assert sol.describe_solution(sol.ZeroSolution())["kind"] == "zero"

    class Wrapper(torch.nn.Module):  # stands for PicardSolutionGradientWrapper (solution_jac.py:113-216)
        def __init__(self, s):
            super().__init__()
            self.solution = s

    assert sol.describe_solution(Wrapper(net))["kind"] == "mlp"

    class Cfg2(_NetCfg):
        BOUND = 3.0

    with pytest.raises(NotImplementedError):
        sol.describe_solution(sol.PicardSolution(eq, Cfg2))

    class CfgP(_NetCfg):
        PISGRADNET, NEURONS = True, [32, 32]

    ou = eqs.OUProcessEquation(nx=6, num_components=2, mean=torch.zeros(2, 6), var=torch.ones(2, 6), pi=torch.ones(2) / 2)
    p = sol.describe_solution(sol.PicardSolution(ou, CfgP))
    assert p["kind"] == "pisgrad" and len(p["main"]) == 3 and len(p["t_encoder"]) == 2 and len(p["smooth_net"]) == 4
    assert tuple(p["main"][0][0].shape) == (32, 6 + 64)


def test_pisgradnet_mirror_matches_oracle_forward():
    g = Golden("hjb_100d_pisgrad")
    var = torch.diag_embed(g.eq.diagvar)
    ou = eqs.OUProcessEquation(nx=100, T=1.0, num_components=5, mean=g.eq.mean, var=var, pi=g.eq.pi)
    net = sol.PISGradNet([96, 64, 80], 100, ou.g, 1.0).double()
    with torch.no_grad():
        net.timestep_phase.copy_(g.net.phase)
        net.timestep_coeff.copy_(g.net.coeff)  # built under the fixture's default dtype (float64)
        for seq, lins in ((net.t_encoder, g.net.t_encoder), (net.smooth_net, g.net.smooth_net),
                          (net.nn_module, g.net.nn_module)):
            for m, (W, b) in zip([m for m in seq if isinstance(m, torch.nn.Linear)], lins):
                m.weight.copy_(W)
                m.bias.copy_(b)
        out = net(g.tx)
    assert torch.allclose(out, g.net.forward(g.tx, g.eq.g), rtol=1e-9, atol=1e-11)


def test_dataset_contract():
    """Boundary caller (picard/dataset.py:20-137): one fused generator call per refresh, batch views out."""
    calls = []

    def gen(n):
        calls.append(n)
        base = sum(calls[:-1])
        x = torch.arange(base, base + n, dtype=torch.float32).view(-1, 1).repeat(1, 3)
        return x, 2 * x

    d = ds.IterableDatasetWithInternalBatch(64, 2, 16, gen)
    batches = list(d)
    assert len(d) == 4 and len(batches) == 4 and calls == [32, 32]
    assert all(x.shape == (16, 3) and torch.equal(y, 2 * x) for x, y in batches)
    assert torch.equal(torch.cat([x for x, _ in batches])[:, 0], torch.arange(64.))
    calls.clear()
    # fractional buffer (the reference splits one batch into 1/f generator calls to bound its path tensors): the
    # buffer is one batch, generated by ONE call
    d2 = ds.IterableDatasetWithInternalBatch(32, 0.25, 16, gen)
    assert len(d2) == 2 and len(list(d2)) == 2 and calls == [16, 16]
    for bad in ((40, 1, 16), (32, 0.3, 16), (32, 1.5, 16)):
        with pytest.raises(ValueError):
            ds.IterableDatasetWithInternalBatch(*bad, gen)


def test_label_cache_replays_epochs_without_regenerating(tmp_path):
    """Label caching (picard/dataset.py:203-255, data_saver.py:69-83): labels are generated once per iteration,
    every epoch replays the stored set, other batch sizes re-slice it, and the on-disk copy keeps the reference's
    dataset names `tx` / `u_ux` (picard/data.py:1513-1521)."""
    calls = []

    def gen(n):
        calls.append(n)
        base = sum(calls[:-1])
        x = torch.arange(base, base + n, dtype=torch.float32).view(-1, 1).repeat(1, 3)
        return x, torch.cat([2 * x, x[:, :1]], 1)

    inner = ds.IterableDatasetWithInternalBatch(16, 1, 16, gen)
    cache = ds.CacheToMemoryWrapper(inner)
    with pytest.raises(RuntimeError):
        iter(cache)
    cache.init(64, [3, 4])  # the data module sets the real size here (picard/data.py:1735-1760)
    first = list(cache)  # generation pass: streams batches while filling the store
    assert calls == [16] * 4 and len(first) == 4 and cache.saver.complete
    for _ in range(3):  # later epochs: no generator call
        again = list(cache)
        assert len(again) == len(cache) == 4
        assert all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(first, again))
    assert calls == [16] * 4
    # preload + a different trainer batch size + file
    calls.clear()
    c2 = ds.CacheToMemoryWrapper(ds.IterableDatasetWithInternalBatch(32, 2, 16, gen), batch_size=8, shuffle=False)
    c2.enable_save_to_file(str(tmp_path / "split_00"))
    c2.init(64, [3, 4], preload=True)
    assert calls == [32, 32] and len(c2) == 8
    b = list(c2)
    assert b[0][0].shape == (8, 3) and torch.equal(torch.cat([x for x, _ in b])[:, 0], torch.arange(64.))
    import glob

    import numpy as np
    (path,) = glob.glob(str(tmp_path / "split_00*"))
    if path.endswith(".npz"):
        z = np.load(path)
        assert z["tx"].shape == (64, 3) and z["u_ux"].shape == (64, 4)
    shuffled = ds.LabelStore(8, [1])
    shuffled.save((torch.arange(8.).view(-1, 1),), 8)
    got = torch.cat([x for (x,) in shuffled.batches(4, shuffle=True, generator=torch.Generator().manual_seed(0))])
    assert sorted(got.view(-1).tolist()) == list(range(8))
    with pytest.raises(ValueError):
        ds.LabelStore(8, [1]).batches(4)


def test_shard_range_partitions():
    for n, w in ((4096, 8), (1 << 20, 8), (10, 4), (3, 8), (0, 2)):
        blocks = [sharding.shard_range(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
        for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
            assert s0 + c0 == s1
        assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _gloo_worker(rank, world, port, n_total, q):
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    start, count = sharding.shard_range(n_total, rank, world)
    # rank-local "labels": a deterministic function of the GLOBAL point index, as the Philox keying guarantees
    idx = torch.arange(start, start + count, dtype=torch.float32).view(-1, 1)
    tx, y = idx.repeat(1, 4), (idx * 0.5).repeat(1, 4)
    gtx, gy = sharding.all_gather_labels(tx, y, n_total)
    ok = (gtx.shape == (n_total, 4) and torch.equal(gtx[:, 0], torch.arange(n_total, dtype=torch.float32))
          and torch.equal(gy, 0.5 * gtx))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo_sharding_and_gather():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 37, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
    assert res == [(0, True), (1, True)]


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference not mounted (GPU box)")
def test_describe_accepts_live_reference_objects():
    """The drop-in boundary takes the REFERENCE's own equation / solution objects (duck-typed)."""
    ref_shim.install()
    import picard.equations as req
    import picard.solution as rsol

    eq = req.Cha(nx=100, alpha=1.0, k=5.0, T=1.0)
    d = eqs.describe_equation(eq)
    assert d["kind"] == "Cha" and abs(d["cha_k"] - 5.0) < 1e-5 and d["nx"] == 100
    mlp = rsol.construct_mlp(101, 1, [128] * 4, ["ELU"] * 4, None)
    assert sol.describe_solution(mlp)["kind"] == "mlp"
    assert sol.describe_solution(rsol.ZeroSolution())["kind"] == "zero"


def test_ctypes_structures_match_the_c_header_layout(tmp_path):
    """sizeof / offsetof of every structure that crosses the C ABI, as gcc lays out include/dpi_b200.h, against the
    ctypes mirrors in _native.py (a silent mismatch would shift every field after it)."""
    import ctypes as C
    import subprocess
    structs = {"dpi_linear": nat.dpi_linear, "dpi_network": nat.dpi_network, "dpi_problem": nat.dpi_problem,
               "dpi_draws": nat.dpi_draws, "dpi_label_config": nat.dpi_label_config, "dpi_train_config": nat.dpi_train_config}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "dpi_b200.h"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{name}.{fname} %zu\\n", offsetof({name}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for name, cls in structs.items():
        assert int(got[name]) == C.sizeof(cls), name
        for fname, _ in cls._fields_:
            assert int(got[f"{name}.{fname}"]) == getattr(cls, fname).offset, f"{name}.{fname}"


@pytest.mark.skipif(not ref_shim.reference_available(), reason="neither /root/reference nor oracle/_ref present")
def test_reference_cpu_arm_builds_the_same_network_and_instance():
    """bench.py's CPU arm (oracle/ref_arm.py) drives the UNMODIFIED reference: its network must carry the bench's
    weights (same function as the oracle's network on random inputs) and its generator must run for every equation."""
    import bench
    from oracle import build_ref, ref_arm

    if os.path.isdir("/root/reference/picard"):
        assert build_ref.build_ref() and build_ref.verify_ref()
    prev = torch.get_default_dtype()
    try:
        for name, M in (("burgers100d", 8), ("gbm100d", 6), ("hjb100d", 6)):
            wl = dict(bench.WORKLOADS[name])
            if name == "hjb100d":
                wl["hidden"] = [48, 40]  # same architecture family, CPU-sized
            net, inst = bench.make_network(wl), bench.problem_instance(wl)
            gen = ref_arm.build_generator(wl, net, inst, M)
            tx, y = gen.sample_with_gradients(3)
            assert tuple(y.shape) == (3, 1 + wl["nx"]) and torch.isfinite(y).all()
            dd = lambda lins: [(W.double(), b.double()) for W, b in lins]
            if net["kind"] == "mlp":
                onet = orc.MLP([W.double() for W, _ in net["main"]], [b.double() for _, b in net["main"]])
            else:
                onet = orc.PISGradNet(net["phase"].double(), net["coeff"].double(), dd(net["t_encoder"]),
                                      dd(net["smooth_net"]), dd(net["main"]), 1.0)
            txr = torch.rand(5, 1 + wl["nx"], dtype=torch.float64)
            u_ref = gen.solution(txr).detach()
            if net["kind"] == "mlp":
                u_orc = onet.forward(txr).reshape_as(u_ref)
                assert torch.allclose(u_ref, u_orc, rtol=1e-12, atol=1e-12)
    finally:
        torch.set_default_dtype(prev)
