"""GPU parity: CUDA label generator (through the C ABI) vs the reference's own outputs (golden fixtures) and vs
the CPU oracle on seeded inputs.  Tolerances are the north-star ones: 1e-4 rel-L2 in fp32 mode, 1e-2 with bf16
GEMM inputs, measured against the fp64 reference on replayed draws."""
import pytest
import torch

from oracle import dpi_oracle as orc
from tests.golden_util import Golden, golden_names

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-4
TOL_BF16 = 1e-2  # north-star tolerance of the 16-bit tensor-core path (fp16 operands; "bf16" is an alias)

BUILT = golden_names()  # every fixture generated from the reference: Burgers / GBM+SDGD / HJB, MLP / PISGradNet / Zero


def _engine(g):
    from tests.gpu_util import engine_for
    return engine_for(g.eq, g.net, g.v)


@pytest.mark.parametrize("name", BUILT)
def test_labels_match_reference_fp32(name):
    from tests.gpu_util import draws_dict
    g = Golden(name)
    e = _engine(g)
    y = e.label_gen(g.tx, g.M_T, g.M_I, precision="fp32", draws=draws_dict(g.draws),
                    estimate_delta_t=g.delta_t).double().cpu()  # delta_t > 0: the TD fixtures (data.py:529-575, :928-952)
    assert torch.isfinite(y).all()
    assert orc.rel_l2(y, g.y) < TOL_FP32
    # gradient columns on their own (the value column dominates the norm otherwise)
    assert orc.rel_l2(y[:, 1:], g.y[:, 1:]) < TOL_FP32
    if float(g.y_integral.norm()) > 0:
        # integral estimator alone: subtract the reference's terminal part
        assert orc.rel_l2(y - g.y_terminal, g.y_integral) < 20 * TOL_FP32
    e.close()


@pytest.mark.parametrize("name", [n for n in BUILT if "zero" in n])
def test_terminal_only_matches_reference(name):
    from tests.gpu_util import draws_dict
    g = Golden(name)
    e = _engine(g)
    y = e.label_gen(g.tx, g.M_T, g.M_I, draws=draws_dict(g.draws), estimate_delta_t=g.delta_t).double().cpu()
    assert orc.rel_l2(y, g.y) < 1e-5
    e.close()


def test_shard_invariance_philox():
    """Philox counters are keyed by the global point index: labels do not depend on how points are sharded."""
    from tests.gpu_util import engine_for
    eq = orc.Cha(100, 1.0, 5.0, 1.0)
    net = orc.make_mlp(101, [128] * 4, seed=5)
    e = engine_for(eq, net)
    tx = e.sample_tx(24, seed=7)
    tx_b = torch.cat([e.sample_tx(10, seed=7), e.sample_tx(14, seed=7, point_index_base=10)])
    assert torch.equal(tx, tx_b)
    y = e.label_gen(tx, 96, 80, seed=7)
    y_b = torch.cat([e.label_gen(tx[:10], 96, 80, seed=7), e.label_gen(tx[10:], 96, 80, seed=7, point_index_base=10)])
    assert torch.equal(y, y_b)
    y2 = e.label_gen(tx, 96, 80, seed=8)
    assert not torch.equal(y, y2)
    e.close()


def test_sample_tx_distribution():
    from tests.gpu_util import engine_for
    eq = orc.Cha(100, 1.0, 5.0, 1.0)
    e = engine_for(eq, orc.ZeroNet())
    tx = e.sample_tx(20000, seed=1).double().cpu()
    t, x = tx[:, :1], tx[:, 1:]
    assert float(t.min()) >= 0.01 - 1e-6 and float(t.max()) <= 0.99 + 1e-6
    assert abs(float(t.mean()) - 0.5) < 0.01
    zs = x / torch.sqrt(t)  # should be N(0,1)
    assert abs(float(zs.mean())) < 5e-3 and abs(float(zs.std()) - 1.0) < 5e-3
    assert abs(float((zs ** 4).mean()) - 3.0) < 0.05
    e.close()


def test_philox_labels_agree_with_oracle_statistically():
    """Where draws cannot be replayed (on-device Philox) the labels must agree with the fp64 oracle on its own
    draws within Monte Carlo error."""
    from tests.gpu_util import engine_for
    eq = orc.Cha(20, 1.0, 5.0, 1.0)
    net = orc.make_mlp(21, [64, 64], seed=3, gain=1.5)
    e = engine_for(eq, net)
    n, M = 64, 4096
    gen = torch.Generator().manual_seed(0)
    d = orc.draw(eq, n, M, M, None, gen)
    tx = orc.sample_tx(eq, d)
    y_ref = orc.generate_with_gradients(eq, net, tx, d, M, M)
    y = e.label_gen(tx, M, M, seed=11).double().cpu()
    # value column: MC std ~ O(0.2/sqrt(M)); compare in rel-L2 over points
    assert orc.rel_l2(y[:, :1], y_ref[:, :1]) < 2e-2
    # two independent MC estimates of the same gradient: their difference is pure noise with variance 2 sigma^2/M
    err = (y[:, 1:] - y_ref[:, 1:])
    both = torch.cat([y[:, 1:], y_ref[:, 1:]])
    assert float(err.mean().abs()) < 5 * float(both.std()) / (n * eq.nx) ** 0.5 + 1e-3
    e.close()


@pytest.mark.parametrize("nx,hidden", [(100, [128] * 4), (10, [64, 64]), (37, [96, 40, 72])])
def test_mlp_value_and_full_gradient_match_autograd(nx, hidden):
    """dpi_net_eval on a plain MLP: u and the full input gradient (one VJP through the batched-GEMM pipeline) vs
    fp64 autograd of the oracle network (picard/data.py:1245-1251 computes exactly this with torch.autograd.grad)."""
    from tests.gpu_util import engine_for
    eq = orc.Cha(nx, 1.0, 5.0, 1.0)
    net = orc.make_mlp(1 + nx, hidden, seed=11, gain=1.4)
    e = engine_for(eq, net)
    gen = torch.Generator().manual_seed(5)
    tx = torch.randn(53, 1 + nx, generator=gen, dtype=torch.float64)
    txg = tx.clone().requires_grad_()
    u_ref = net.forward(txg)
    (g_ref,) = torch.autograd.grad(u_ref.sum(), txg)
    u, du = e.net_eval(tx, want_grad=True)
    assert orc.rel_l2(u.double().cpu().view(-1, 1), u_ref.detach()) < 1e-5
    assert orc.rel_l2(du.double().cpu(), g_ref[:, 1:]) < 1e-5
    e.close()


def test_pisgradnet_value_and_gradient_match_oracle():
    """dpi_net_eval on PISGradNet: u and the closed-form grad_x u (one VJP through nn_module) vs the fp64 oracle,
    whose closed form is itself checked against autograd of the reference network in tests/test_oracle_golden.py."""
    g = Golden("hjb_100d_pisgrad")
    e = _engine(g)
    gen = torch.Generator().manual_seed(3)
    tx = torch.cat([torch.rand(37, 1, generator=gen, dtype=torch.float64) * 0.98 + 0.01,
                    torch.randn(37, g.nx, generator=gen, dtype=torch.float64) * 1.5], dim=1)
    u_ref, du_ref = g.net.value_and_grad_closed(tx, g.eq)
    u, du = e.net_eval(tx, want_grad=True)
    assert orc.rel_l2(u.double().cpu().view(-1, 1), u_ref) < 1e-5
    assert orc.rel_l2(du.double().cpu(), du_ref) < 1e-5
    e.close()


def test_hjb_full_width_network_runs_chunked_and_is_shard_invariant():
    """The shipped HJB shape (hidden 4 x 512, solution.py:195-203) through the batched-GEMM pipeline with Philox
    draws: finite labels, independent of how the points are split across calls."""
    from tests.gpu_util import engine_for
    g = Golden("hjb_100d_pisgrad")
    net = orc.make_pisgrad(g.nx, [512] * 4, T=g.T, seed=4, phase_scale=0.3)
    e = engine_for(g.eq, net)
    tx = e.sample_tx(12, seed=2)
    y = e.label_gen(tx, 64, 200, seed=9)
    assert torch.isfinite(y).all()
    y_b = torch.cat([e.label_gen(tx[:5], 64, 200, seed=9), e.label_gen(tx[5:], 64, 200, seed=9, point_index_base=5)])
    assert torch.equal(y, y_b)
    # against the fp64 oracle on its own draws: the value column agrees within Monte Carlo error
    gen = torch.Generator().manual_seed(0)
    d = orc.draw(g.eq, 12, 64, 200, None, gen)
    y_ref = orc.generate_with_gradients(g.eq, net, tx.double().cpu(), d, 64, 200)
    assert orc.rel_l2(y[:, :1].double().cpu(), y_ref[:, :1]) < 5e-2
    e.close()


def test_hjb_tensor_core_gemms_track_the_fp32_pipeline():
    """PISGradNet with the wide nn_module GEMMs on tcgen05 (TF32 operands read from the fp32 buffers, fp32
    accumulation): replayed reference draws against the fp64 oracle, the fp32 pipeline on the same Philox draws
    (forward and [K][N] backward weights, partial K chunks 164 and 100, ragged row and column tiles), shard invariance."""
    from tests.gpu_util import engine_for, draws_dict
    g = Golden("hjb_100d_pisgrad")
    net = orc.make_pisgrad(g.nx, [512] * 4, T=g.T, seed=4, phase_scale=0.3)
    e = engine_for(g.eq, net)
    n, M = 5, 40  # 205 rows: one full and one ragged 128-row tile
    d = orc.draw(g.eq, n, M, M, None, torch.Generator().manual_seed(3), round_fp32=True)
    tx = orc.sample_tx(g.eq, d)
    y_ref = orc.generate_with_gradients(g.eq, net, tx, d, M, M)
    y19 = e.label_gen(tx, M, M, precision="tf32", draws=draws_dict(d)).double().cpu()
    assert orc.rel_l2(y19, y_ref) < TOL_BF16
    assert orc.rel_l2(y19[:, 1:], y_ref[:, 1:]) < TOL_BF16
    txp = e.sample_tx(20, seed=2)
    y32 = e.label_gen(txp, 64, 300, precision="fp32", seed=9).double().cpu()
    y19 = e.label_gen(txp, 64, 300, precision="tf32", seed=9).double().cpu()
    assert torch.isfinite(y19).all()
    assert orc.rel_l2(y19, y32) < 2e-3 and orc.rel_l2(y19[:, 1:], y32[:, 1:]) < 2e-3
    assert not torch.equal(y19, y32)  # the tensor-core path really ran
    y_b = torch.cat([e.label_gen(txp[:7], 64, 300, precision="tf32", seed=9),
                     e.label_gen(txp[7:], 64, 300, precision="tf32", seed=9, point_index_base=7)]).double().cpu()
    assert torch.equal(y19, y_b)
    e.close()


# ---------------------------------------------------------------------------------------------------------
# bf16 tensor-core (tcgen05) path
# ---------------------------------------------------------------------------------------------------------
def test_umma_selftest_matches_bf16_matmul():
    """tcgen05 building blocks in isolation: descriptors, SWIZZLE_128B operand images, TMEM readback."""
    from tests.gpu_util import engine_for
    e = engine_for(orc.Cha(100, 1.0, 5.0, 1.0), orc.ZeroNet())
    g = torch.Generator().manual_seed(0)
    A = torch.randn(128, 128, generator=g)
    B = torch.randn(128, 128, generator=g)
    D = e.selftest_umma(A, B).cpu().double()
    ref = A.half().double() @ B.half().double().t()
    assert orc.rel_l2(D, ref) < 1e-5
    # structured operands catch transposition / swizzle mix-ups that random data could hide statistically
    A2 = torch.arange(128.).view(-1, 1).repeat(1, 128) / 128 + torch.eye(128)
    B2 = torch.arange(128.).view(1, -1).repeat(128, 1) / 64 - torch.eye(128)
    D2 = e.selftest_umma(A2, B2).cpu().double()
    ref2 = A2.half().double() @ B2.half().double().t()
    assert orc.rel_l2(D2, ref2) < 1e-5
    e.close()


TC_GOLDEN = ["burgers_100d_mlp128", "gbm_100d_mlp64_sdgd", "burgers_100d_mlp128_td"]


@pytest.mark.parametrize("name", TC_GOLDEN)
def test_labels_match_reference_bf16(name):
    from tests.gpu_util import draws_dict
    g = Golden(name)
    e = _engine(g)
    y = e.label_gen(g.tx, g.M_T, g.M_I, precision="f16", draws=draws_dict(g.draws), estimate_delta_t=g.delta_t).double().cpu()
    assert torch.isfinite(y).all()
    assert orc.rel_l2(y, g.y) < TOL_BF16
    assert orc.rel_l2(y[:, 1:], g.y[:, 1:]) < TOL_BF16
    assert orc.rel_l2(y[:, :1], g.y[:, :1]) < TOL_BF16
    assert orc.rel_l2(y, g.y) < 1e-3  # measured 6.2e-5: far inside the north-star tolerance
    # the integral estimator on its own (terminal part subtracted): for this random-init net it is ~1 % of the label
    # norm and its gradient columns are a low signal-to-noise MC sum, so it is the sharpest view on the tensor-core path
    assert orc.rel_l2(y - g.y_terminal, g.y_integral) < TOL_BF16
    e.close()


@pytest.mark.parametrize("nx,M_T,M_I,n", [(100, 64, 128, 7), (100, 256, 4096, 5), (100, 32, 1100, 9), (64, 16, 300, 4),
                                          (112, 16, 200, 3), (10, 16, 131, 6), (77, 16, 257, 3)])
def test_bf16_path_agrees_with_fp32_path_on_the_same_philox_draws(nx, M_T, M_I, n):
    """Both integral kernels consume identical Philox counters, so they differ only by bf16 rounding of the GEMM
    inputs; covers ragged M_I (not a multiple of 128), several work units per point and d not a multiple of 4/16."""
    from tests.gpu_util import engine_for
    eq = orc.Cha(nx, 1.0, 5.0, 1.0)
    net = orc.make_mlp(1 + nx, [128] * 4, seed=2, gain=1.3)
    e = engine_for(eq, net)
    tx = e.sample_tx(n, seed=3)
    y32 = e.label_gen(tx, M_T, M_I, precision="fp32", seed=5).double().cpu()
    y16 = e.label_gen(tx, M_T, M_I, precision="f16", seed=5).double().cpu()
    assert torch.isfinite(y16).all()
    assert orc.rel_l2(y16, y32) < TOL_BF16
    assert orc.rel_l2(y16[:, :1], y32[:, :1]) < TOL_BF16
    e.close()


@pytest.mark.parametrize("gain", [1.0, 2.0, 3.0])
def test_tensor_core_integral_part_tracks_fp32_for_large_weights(gain):
    """Networks with large weights (what training produces) make the integral estimator a large part of the label; the
    16-bit path must follow the fp32 path on the same Philox draws there too.  (This is the case that exposed a
    doubly-scaled time column: invisible on default-init weights, 4-7 % off at gain >= 2.)"""
    from tests.gpu_util import engine_for
    eq = orc.Cha(100, 1.0, 5.0, 1.0)
    net = orc.make_mlp(101, [128] * 4, seed=2, gain=gain)
    e, ez = engine_for(eq, net), engine_for(eq, orc.ZeroNet())
    tx = e.sample_tx(128, seed=3)
    yz = ez.label_gen(tx, 512, 1024, seed=5).double()  # terminal part only (u_k == 0)
    i32 = e.label_gen(tx, 512, 1024, precision="fp32", seed=5).double() - yz
    i16 = e.label_gen(tx, 512, 1024, precision="f16", seed=5).double() - yz
    assert orc.rel_l2(i16[:, :1], i32[:, :1]) < 3e-3
    assert orc.rel_l2(i16[:, 1:], i32[:, 1:]) < TOL_BF16
    e.close()
    ez.close()


def test_gbm_tensor_core_path_tracks_fp32_on_the_same_philox_draws():
    """SDGD 2-jets through tcgen05 (fp16 operands) vs the fp32 kernel: identical Philox counters for increments and
    sampled coordinates, so the two differ only by operand rounding; also shard invariance of the tensor-core kernel."""
    g = Golden("gbm_100d_mlp64_sdgd")
    net = orc.make_mlp(101, [64, 64, 64], seed=4, gain=1.5)
    from tests.gpu_util import engine_for
    e = engine_for(g.eq, net, 100)
    tx = e.sample_tx(24, seed=2)
    y32 = e.label_gen(tx, 128, 160, precision="fp32", seed=6).double().cpu()
    y16 = e.label_gen(tx, 128, 160, precision="f16", seed=6).double().cpu()
    assert torch.isfinite(y16).all()
    assert orc.rel_l2(y16, y32) < TOL_BF16
    assert orc.rel_l2(y16[:, :1], y32[:, :1]) < TOL_BF16
    y_b = torch.cat([e.label_gen(tx[:9], 128, 160, precision="f16", seed=6),
                     e.label_gen(tx[9:], 128, 160, precision="f16", seed=6, point_index_base=9)]).double().cpu()
    assert torch.equal(y16, y_b)
    e.close()


def test_bf16_shard_invariance_and_determinism():
    from tests.gpu_util import engine_for
    eq = orc.Cha(100, 1.0, 5.0, 1.0)
    net = orc.make_mlp(101, [128] * 4, seed=5)
    e = engine_for(eq, net)
    tx = e.sample_tx(300, seed=7)
    y = e.label_gen(tx, 64, 1500, precision="f16", seed=7)
    assert torch.equal(y, e.label_gen(tx, 64, 1500, precision="f16", seed=7))
    y_b = torch.cat([e.label_gen(tx[:11], 64, 1500, precision="f16", seed=7),
                     e.label_gen(tx[11:], 64, 1500, precision="f16", seed=7, point_index_base=11)])
    assert torch.equal(y, y_b)
    e.close()


@pytest.mark.parametrize("nx,hidden", [(100, [96, 64, 128, 32]), (37, [64, 64, 64, 64]), (200, [120, 128, 72, 8])])
def test_tensor_core_path_takes_narrower_hidden_layers(nx, hidden):
    """The tcgen05 kernel is written for 128-wide operand images; hidden layers narrower than that are zero-padded at
    dpi_set_network (a padded unit contributes exactly zero), so four-hidden-layer networks of any widths <= 128 take
    the 16-bit path.  Checked against the pinned oracle in deterministic mode and against the fp32 kernel on the same
    Philox draws."""
    from tests.gpu_util import draws_dict, engine_for
    eq = orc.Cha(nx, 1.0, 5.0, 1.0)
    net = orc.make_mlp(1 + nx, hidden, seed=11, gain=1.5)
    n, M_T, M_I = 5, 64, 300
    d = orc.draw(eq, n, M_T, M_I, None, torch.Generator().manual_seed(7), round_fp32=True)
    tx = orc.sample_tx(eq, d).float().double()
    yT = orc.estimate_terminal_with_gradients(eq, tx, d.zT, M_T)
    y_ref = (yT + orc.estimate_integral_with_gradients(eq, net, tx, d.u_s, d.zS, d.idx, M_I, mode="closed")).detach()
    e = engine_for(eq, net)
    y16 = e.label_gen(tx, M_T, M_I, precision="f16", draws=draws_dict(d), check=True).double().cpu()
    assert orc.rel_l2(y16, y_ref) < TOL_BF16
    assert orc.rel_l2(y16 - yT, y_ref - yT) < TOL_BF16
    y32 = e.label_gen(tx, M_T, M_I, precision="fp32", draws=draws_dict(d), check=True).double().cpu()
    assert orc.rel_l2(y32, y_ref) < 1e-4
    txp = e.sample_tx(9, seed=3)
    p32 = e.label_gen(txp, 64, 1100, precision="fp32", seed=5).double().cpu()
    p16 = e.label_gen(txp, 64, 1100, precision="f16", seed=5).double().cpu()
    assert torch.isfinite(p16).all()
    assert orc.rel_l2(p16, p32) < TOL_BF16
    e.close()


def test_bf16_rejects_unsupported_network_loudly():
    from tests.gpu_util import engine_for
    eq = orc.Cha(100, 1.0, 5.0, 1.0)
    e = engine_for(eq, orc.make_mlp(101, [64, 64], seed=1))
    tx = e.sample_tx(4, seed=0)
    with pytest.raises(NotImplementedError):
        e.label_gen(tx, 16, 16, precision="f16")
    e.close()


def test_td_estimators_philox_mode_and_generator_keyword():
    """ESTIMATE_DELTA_T through the host mirror in Philox mode: statistically the same labels as the oracle's TD
    estimators on its own draws (M = 4096), shard invariant, and rejected for the Hessian equation as upstream asserts."""
    from deeppicarditeration_b200 import equations as eqs
    from deeppicarditeration_b200 import solution as sol
    from deeppicarditeration_b200.data import OnlineDataGenerator
    from tests.gpu_util import engine_for

    class HA:
        method, kwargs = None, {}

    class Cfg:
        TYPE, PISGRADNET, NEURONS, ACTIVATIONS, BOUND = "Value", False, [64, 64], ["ELU", "ELU"], None

    torch.manual_seed(0)
    eq = eqs.Cha(nx=20, alpha=1.0, k=5.0, T=1.0)
    net = sol.PicardSolution(eq, Cfg).float().cuda()
    gen = OnlineDataGenerator(eq, net, 40, 2, device="cuda", t_always_uniform=True, n_estimate_terminal=4096,
                              n_estimate_integral=4096, hessian_approximation=HA(), estimate_delta_t=0.25, seed=3)
    tx, y = gen.sample_with_gradients(24)
    assert torch.isfinite(y).all() and bool((tx[:, 0] + 0.25 < 1.0).any()) and bool((tx[:, 0] + 0.25 >= 1.0).any())
    # oracle on its own draws at the same points
    oeq = orc.Cha(20, 1.0, 5.0, 1.0)
    lin = [m for m in net.model if isinstance(m, torch.nn.Linear)]
    onet = orc.MLP([m.weight.detach().double().cpu() for m in lin], [m.bias.detach().double().cpu() for m in lin])
    g = torch.Generator().manual_seed(5)
    d = orc.draw(oeq, 24, 4096, 4096, None, g)
    yo = orc.generate_with_gradients(oeq, onet, tx.double().cpu(), d, 4096, 4096, delta_t=0.25)
    assert orc.rel_l2(y[:, :1].double().cpu(), yo[:, :1]) < 2e-2  # value labels: MC error ~ 1/sqrt(4096)
    # bitwise shard invariance of the TD path
    e = gen.engine
    ya = e.label_gen(tx, 64, 96, seed=9, estimate_delta_t=0.25)
    yb = torch.cat([e.label_gen(tx[:10], 64, 96, seed=9, estimate_delta_t=0.25),
                    e.label_gen(tx[10:], 64, 96, seed=9, point_index_base=10, estimate_delta_t=0.25)])
    assert torch.equal(ya, yb)
    e.close()
    gg = Golden("gbm_100d_mlp64_sdgd")
    eg = engine_for(gg.eq, gg.net, gg.v)
    with pytest.raises(Exception):
        eg.label_gen(gg.tx, 8, 8, estimate_delta_t=0.2)
    eg.close()


def test_td_estimators_on_the_fused_pisgradnet_kernel():
    """ESTIMATE_DELTA_T on the fused fp16 PISGradNet kernel (its horizon comes from the same helper as every other
    estimator kernel): same Philox draws, fp32 pipeline as the reference; both branches of t_next present."""
    from tests.gpu_util import engine_for
    g = Golden("hjb_100d_pisgrad")
    net = orc.make_pisgrad(100, [128, 128], T=1.0, seed=9, phase_scale=0.3)
    e = engine_for(g.eq, net)
    tx = e.sample_tx(40, seed=2)
    assert bool((tx[:, 0] + 0.3 < 1.0).any()) and bool((tx[:, 0] + 0.3 >= 1.0).any())
    y32 = e.label_gen(tx, 256, 512, precision="fp32", seed=4, estimate_delta_t=0.3).double().cpu()
    y16 = e.label_gen(tx, 256, 512, precision="f16", seed=4, estimate_delta_t=0.3).double().cpu()
    y0 = e.label_gen(tx, 256, 512, precision="f16", seed=4).double().cpu()
    assert torch.isfinite(y16).all() and orc.rel_l2(y16, y32) < 1e-2
    assert orc.rel_l2(y0, y32) > 1e-2  # (the TD labels differ from the plain ones: the keyword reaches the kernel)
    e.close()
