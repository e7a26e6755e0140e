"""GPU tests of what deterministic-mode parity cannot see: the on-device generator (Philox4x32-10 + the 16+16-bit
Box-Muller of csrc/dpi_common.cuh), the (t, x) sampler of every equation, the SDGD coordinate stream, and the guards
around the reduced-precision paths (range check, frozen network copies, rejected bf16)."""
import math

import numpy as np
import pytest
import torch

from oracle import dpi_oracle as orc
from tests.golden_util import Golden

pytestmark = pytest.mark.gpu


def _eng(eq, net, v=0):
    from tests.gpu_util import engine_for
    return engine_for(eq, net, v)


def test_ou_sample_tx_moments():
    """OUProcessEquation.sample_x0 ~ N(0, alpha_scale * alpha I) (equations.py:710-711, utils.py:785-789) then
    x = x0 + sqrt(t alpha) Z (equations.py:560-573): per-coordinate variance 4 alpha + t alpha, mean 0, Gaussian."""
    g = Golden("hjb_100d_zero")
    e = _eng(g.eq, orc.ZeroNet())
    n = 40000
    tx = e.sample_tx(n, seed=5).double().cpu()
    t, x = tx[:, :1], tx[:, 1:]
    assert float(t.min()) >= 0.01 - 1e-6 and float(t.max()) <= 0.99 + 1e-6 and abs(float(t.mean()) - 0.5) < 0.01
    zs = x / torch.sqrt(g.eq.alpha_init + t * g.eq.alpha)  # standardised: N(0, 1) if the variance is 4 alpha + t alpha
    N = zs.numel()
    assert abs(float(zs.mean())) < 5 / math.sqrt(N)
    assert abs(float(zs.var()) - 1.0) < 5 * math.sqrt(2.0 / N)
    assert abs(float((zs ** 4).mean()) - 3.0) < 5 * math.sqrt(96.0 / N)
    # the x0 and the diffusion parts are independent streams: the variance must GROW with t at slope alpha
    lo, hi = (t[:, 0] < 0.25), (t[:, 0] > 0.75)
    dv = float(x[hi].var()) - float(x[lo].var())
    t_gap = float(t[hi].mean() - t[lo].mean())
    assert abs(dv - g.eq.alpha * t_gap) < 0.03
    # coordinates are independent of each other
    c = np.corrcoef(zs[:, :8].numpy().T)
    assert np.abs(c - np.eye(8)).max() < 5 / math.sqrt(n)
    e.close()


def test_philox_sdgd_indices_are_uniform_and_independent():
    """SDGD coordinates (data.py:497-502: randint(0, nx) with replacement): chi-square of the exported index stream
    against the uniform law on {0..d-1}, of consecutive pairs on a 10 x 10 grid, and of the per-column marginals."""
    g = Golden("gbm_100d_zero_sdgd")
    e = _eng(g.eq, orc.ZeroNet(), 100)
    n, M, v, d = 16, 1024, 100, 100
    idx = e.export_draws(n, 16, M, seed=9, want=("idx",))["idx"].cpu().numpy()
    assert idx.shape == (n * M, v) and idx.min() == 0 and idx.max() == d - 1
    N = idx.size
    counts = np.bincount(idx.ravel(), minlength=d).astype(np.float64)
    chi2 = float(((counts - N / d) ** 2 / (N / d)).sum())
    assert chi2 < 99 + 5 * math.sqrt(2 * 99), chi2  # 99 degrees of freedom, 5 sigma
    a, b = idx[:, :-1].ravel() // 10, idx[:, 1:].ravel() // 10  # consecutive draws (same and neighbouring Philox blocks)
    joint = np.bincount(a * 10 + b, minlength=100).astype(np.float64)
    chi2p = float(((joint - a.size / 100) ** 2 / (a.size / 100)).sum())
    assert chi2p < 99 + 5 * math.sqrt(2 * 99), chi2p
    # same row, different samples / points: column 0 against column 0 of the next row
    a, b = idx[:-1, 0] // 10, idx[1:, 0] // 10
    joint = np.bincount(a * 10 + b, minlength=100).astype(np.float64)
    chi2r = float(((joint - a.size / 100) ** 2 / (a.size / 100)).sum())
    assert chi2r < 99 + 5 * math.sqrt(2 * 99), chi2r
    # what the kernels consume is what export writes: Philox-mode labels == deterministic replay (see baseline-size tests)
    e.close()


def test_exported_normals_have_normal_moments():
    eq = orc.Cha(100, 1.0, 5.0, 1.0)
    e = _eng(eq, orc.ZeroNet())
    dr = e.export_draws(64, 2048, 2048, seed=3, want=("zT", "zS", "u_s"))
    for k in ("zT", "zS"):
        z = dr[k].double().cpu()
        N = z.numel()
        assert abs(float(z.mean())) < 5 / math.sqrt(N)
        assert abs(float((z ** 2).mean()) - 1.0) < 5 * math.sqrt(2.0 / N)
        assert abs(float((z ** 4).mean()) - 3.0) < 5 * math.sqrt(96.0 / N) + 2e-4  # E z^4 = 2.99982 by construction
        assert float(z.abs().max()) < 4.87  # 16-bit radius: |z| <= sqrt(-2 ln 2^-17)
    u = dr["u_s"].double().cpu()
    assert 0.0 < float(u.min()) and float(u.max()) < 1.0 and abs(float(u.mean()) - 0.5) < 5 / math.sqrt(12 * u.numel())
    assert not torch.equal(dr["zT"][: 2048 * 100 // 100], dr["zS"][: 2048 * 100 // 100])  # distinct streams
    e.close()


def test_box_muller_bias_is_below_1e4_on_a_label():
    """The 16+16-bit Box-Muller is a quadrature rule, not an exact sampler; its systematic effect on a label is bounded
    here where the answer is known: the Cha terminal value E[sigmoid(T + k (sum x + a S))], S = sum of d increments
    ~ N(0, d), against Gauss-Hermite quadrature.  6.2e8 samples x 100 dims bring the MC noise to ~2e-5 relative."""
    eq = orc.Cha(100, 1.0, 5.0, 1.0)
    e = _eng(eq, orc.ZeroNet())
    n, M_T = 2368, 1 << 18
    t, xj = 0.3, 0.02
    tx = torch.full((n, 101), xj, dtype=torch.float32)
    tx[:, 0] = t
    y = e.label_gen(tx, M_T, 1, seed=123).double().cpu()  # u_k == 0: labels are the terminal estimator
    mc = float(y[:, 0].mean())
    sem = float(y[:, 0].std()) / math.sqrt(n)
    t32, x32 = float(np.float32(t)), float(np.float32(xj))
    a = math.sqrt(eq.T - t32) * eq.alpha_sqrt
    xi, w = np.polynomial.hermite_e.hermegauss(200)
    arg = eq.T + eq.k * (100 * x32 + a * math.sqrt(100.0) * xi)
    exact = float((w / math.sqrt(2 * math.pi) * (1.0 / (1.0 + np.exp(-arg)))).sum())
    rel = abs(mc - exact) / exact
    print(f"terminal value: MC {mc:.8f} +- {sem:.1e}  Gauss-Hermite {exact:.8f}  rel diff {rel:.2e}")
    assert rel < 1e-4
    assert abs(mc - exact) < 5 * sem + 2e-5  # and consistent with pure MC noise plus fp32 accumulation
    # gradient label, same construction: E[dg Z_j]/a = k E[sigma'(.)] for every j
    gexact = float((w / math.sqrt(2 * math.pi) * eq.k * np.exp(-arg) / (1.0 + np.exp(-arg)) ** 2).sum())
    gmc = float(y[:, 1:].mean())
    assert abs(gmc - gexact) / gexact < 1e-3
    e.close()


def test_fp16_range_guard_reports_instead_of_returning_nan():
    """A network whose pre-activations leave fp16's range on the tensor-core path: labels would be inf / NaN; the
    library reports DPI_ERR_NONFINITE (FloatingPointError) and the fp32 path still works."""
    eq = orc.Cha(100, 1.0, 5.0, 1.0)
    net = orc.make_mlp(101, [128] * 4, seed=3, gain=60.0)
    e = _eng(eq, net)
    tx = e.sample_tx(16, seed=1)
    with pytest.raises(FloatingPointError):
        e.label_gen(tx, 64, 256, precision="f16", seed=2, check=True)
    e.check_labels()  # the status word was cleared by the failed check
    y = e.label_gen(tx, 64, 256, precision="f16", seed=2)  # unchecked call: the flag is raised again ...
    assert not torch.isfinite(y).all()
    with pytest.raises(FloatingPointError):
        e.check_labels()  # ... and reported by the explicit check
    with pytest.raises(FloatingPointError):  # host-buffer entry point checks on its own
        e.label_gen_host(16, 64, 256, tx_host=tx.cpu().pin_memory(), precision="f16", seed=2)
    y32 = e.label_gen(tx, 64, 256, precision="fp32", seed=2, check=True)
    assert torch.isfinite(y32).all()
    # a well-scaled network passes the check on every path
    e2 = _eng(eq, orc.make_mlp(101, [128] * 4, seed=3, gain=2.0))
    e2.label_gen(tx, 64, 256, precision="f16", seed=2, check=True)
    e.close()
    e2.close()


def test_bf16_is_rejected_not_aliased():
    eq = orc.Cha(100, 1.0, 5.0, 1.0)
    e = _eng(eq, orc.make_mlp(101, [128] * 4, seed=5))
    tx = e.sample_tx(8, seed=1)
    with pytest.raises(NotImplementedError):
        e.label_gen(tx, 32, 256, precision="bf16", seed=2)
    e.close()


def test_point_index_range_is_validated():
    eq = orc.Cha(10, 1.0, 5.0, 1.0)
    e = _eng(eq, orc.ZeroNet())
    with pytest.raises(RuntimeError):
        e.sample_tx(4, seed=1, point_index_base=(1 << 48) - 2)
    tx = e.sample_tx(4, seed=1, point_index_base=(1 << 48) - 4)
    with pytest.raises(RuntimeError):
        e.label_gen(tx, 8, 8, seed=1, point_index_base=-1)
    from deeppicarditeration_b200 import equations as eqs, solution as sol
    from deeppicarditeration_b200.data import OnlineDataGenerator

    class _HA:
        method, kwargs = None, {}
    with pytest.raises(ValueError):
        OnlineDataGenerator(eqs.Cha(nx=10, alpha=1.0, k=5.0), sol.ZeroSolution(), 10, 1, device="cuda",
                            t_always_uniform=True, hessian_approximation=_HA(), rank=300, world_size=512)
    e.close()


def test_set_network_freezes_a_private_copy():
    """data.py:407-412: the generator labels from a FROZEN u_k.  Training (or freeing) the caller's tensors after
    dpi_set_network must not change the labels or dpi_net_eval -- plain MLP and PISGradNet."""
    from deeppicarditeration_b200.engine import LabelEngine
    eq = orc.Cha(100, 1.0, 5.0, 1.0)
    net = orc.make_mlp(101, [128] * 4, seed=6, gain=1.5)
    Ws = [(W.float().cuda().contiguous(), b.float().cuda().contiguous()) for W, b in zip(net.weights, net.biases)]
    e = LabelEngine("cuda")
    e.set_problem("Cha", 100, 1.0, 1.0, cha_k=5.0)
    e.set_network("mlp", main=Ws)
    tx = e.sample_tx(16, seed=1)
    y0 = e.label_gen(tx, 64, 256, seed=2)
    u0, du0 = e.net_eval(tx, want_grad=True)
    for W, b in Ws:  # "train in place"
        W.mul_(0.5)
        b.add_(1.0)
    torch.cuda.synchronize()
    assert torch.equal(e.label_gen(tx, 64, 256, seed=2), y0)
    u1, du1 = e.net_eval(tx, want_grad=True)
    assert torch.equal(u1, u0) and torch.equal(du1, du0)
    e.close()

    g = Golden("hjb_100d_pisgrad")
    pn = orc.make_pisgrad(g.nx, [64, 64], T=g.T, seed=4, phase_scale=0.3)
    cu = lambda lins: [(W.float().cuda().contiguous(), b.float().cuda().contiguous()) for W, b in lins]
    main, tenc, smooth = cu(pn.nn_module), cu(pn.t_encoder), cu(pn.smooth_net)
    phase, coeff = pn.phase.float().cuda(), pn.coeff.float().cuda()
    e = LabelEngine("cuda")
    e.set_problem("OU", g.nx, g.T, g.eq.alpha, ou_mean=g.eq.mean.numpy(), ou_diagvar=g.eq.diagvar.numpy(),
                  ou_pi=g.eq.pi.numpy(), ou_theta=g.eq.theta, ou_mu=g.eq.mu, ou_alpha_init=g.eq.alpha_init)
    e.set_network("pisgrad", main=main, t_encoder=tenc, smooth_net=smooth, phase=phase, coeff=coeff)
    tx = e.sample_tx(6, seed=1)
    y0 = e.label_gen(tx, 32, 48, seed=2)
    u0, du0 = e.net_eval(tx, want_grad=True)
    for W, b in main + tenc + smooth:
        W.mul_(0.5)
        b.add_(1.0)
    phase.add_(1.0)
    coeff.mul_(2.0)
    torch.cuda.synchronize()
    assert torch.equal(e.label_gen(tx, 32, 48, seed=2), y0)
    u1, du1 = e.net_eval(tx, want_grad=True)
    assert torch.equal(u1, u0) and torch.equal(du1, du0)
    e.close()


def test_hjb_narrow_network_with_one_sample_per_point():
    """The MC-reduction partials of the PISGradNet pipeline need 8 (1 + d) floats per point, more than the activation
    buffer they used to alias when M_I is tiny and the network narrow (n_estimate_integral defaults to 1,
    data.py:380).  Deterministic mode against the oracle, M_I = 1 and 2, hidden [64, 64]."""
    from tests.gpu_util import draws_dict, engine_for
    g = Golden("hjb_100d_pisgrad")
    net = orc.make_pisgrad(g.nx, [64, 64], T=g.T, seed=8, phase_scale=0.3)
    e = engine_for(g.eq, net)
    for M_I in (1, 2):
        n, M_T = 37, 4
        d = orc.draw(g.eq, n, M_T, M_I, None, torch.Generator().manual_seed(M_I), round_fp32=True)
        tx = orc.sample_tx(g.eq, d).float().double()
        y_ref = orc.generate_with_gradients(g.eq, net, tx, d, M_T, M_I)
        y = e.label_gen(tx, M_T, M_I, draws=draws_dict(d), check=True).double().cpu()
        assert orc.rel_l2(y, y_ref) < 1e-4, (M_I, orc.rel_l2(y, y_ref))
        yp = e.label_gen(tx, M_T, M_I, seed=3)
        yq = torch.cat([e.label_gen(tx[:9], M_T, M_I, seed=3), e.label_gen(tx[9:], M_T, M_I, seed=3, point_index_base=9)])
        assert torch.equal(yp, yq)
    e.close()


def test_custom_op_is_the_generator_path_and_streams_advance():
    """`torch.ops.picard_b200.label_gen` is what OnlineDataGenerator calls; generate_with_gradients advances the Philox
    stream between calls and does not clip (data.py:1208-1218 vs :222)."""
    from deeppicarditeration_b200 import equations as eqs, solution as sol
    from deeppicarditeration_b200.data import OnlineDataGenerator

    class _HA:
        method, kwargs = None, {}

    class _NetCfg:
        TYPE, PISGRADNET, NEURONS, ACTIVATIONS, BOUND = "Value", False, [128] * 4, ["ELU"] * 4, None

    eq = eqs.Cha(nx=100, alpha=1.0, k=5.0, T=1.0)
    torch.manual_seed(0)
    net = sol.PicardSolution(eq, _NetCfg)
    gen = OnlineDataGenerator(eq, net, 10, 2, device="cuda", t_always_uniform=True, n_estimate_terminal=64,
                              n_estimate_integral=128, hessian_approximation=_HA(), sample_bound=0.3, seed=5)
    eng = gen.engine
    tx = eng.sample_tx(8, seed=5)
    y_op = torch.ops.picard_b200.label_gen(tx, eng.id, 64, 128, "fp32", 0.0, 5, 0)
    assert torch.equal(y_op, eng.label_gen(tx, 64, 128, seed=5, point_index_base=0))
    calls = []
    import deeppicarditeration_b200.engine as engine_mod
    orig = engine_mod.LabelEngine.label_gen

    def spy(self, *a, **k):
        calls.append(k.get("point_index_base"))
        return orig(self, *a, **k)

    engine_mod.LabelEngine.label_gen = spy
    try:
        y1 = gen.generate_with_gradients(tx)
        y2 = gen.generate_with_gradients(tx)
    finally:
        engine_mod.LabelEngine.label_gen = orig
    assert calls == [0, 8]  # through the op, bases advance by the batch
    assert torch.equal(y1, y_op) and not torch.equal(y1, y2)  # independent MC noise on the second call
    assert float(y1.abs().max()) > 0.3  # generate_with_gradients does not clip ...
    _, ys = gen.sample_with_gradients(8)
    assert float(ys.abs().max()) <= 0.3 + 1e-6  # ... sample_with_gradients does (0.3 rounded to fp32)
    gen.engine.close()
