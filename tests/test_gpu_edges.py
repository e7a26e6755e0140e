"""Edge cases and full-size properties of the label path on the GPU (through the C ABI): empty and single-point calls,
clipping, the 1000-dimensional sweep shape against the oracle, chunked large calls, and size-independent properties at
the BASELINE sizes (linearity of the terminal estimator in g, label statistics against the exact solution)."""
import pytest
import torch

from oracle import dpi_oracle as orc

pytestmark = pytest.mark.gpu


def _eng(eq, net, v=0):
    from tests.gpu_util import engine_for
    return engine_for(eq, net, v)


def test_empty_and_single_point_calls():
    eq = orc.Cha(100, 1.0, 5.0, 1.0)
    e = _eng(eq, orc.make_mlp(101, [128] * 4, seed=1))
    tx0 = e.sample_tx(0, seed=1)
    assert tx0.shape == (0, 101)
    for prec in ("fp32", "f16"):
        assert e.label_gen(tx0, 16, 16, precision=prec).shape == (0, 101)
    tx = e.sample_tx(3, seed=1)
    for prec in ("fp32", "f16"):
        y1 = e.label_gen(tx[1:2], 64, 64, precision=prec, seed=4, point_index_base=1)
        y3 = e.label_gen(tx, 64, 64, precision=prec, seed=4)
        assert torch.equal(y1, y3[1:2])  # a single point gets exactly the label it gets inside a batch
    e.close()


def test_sample_bound_clips_like_the_reference():
    """data.py:222: y = clip(y, -bound, +bound) after terminal + integral."""
    eq = orc.Cha(50, 1.0, 5.0, 1.0)
    e = _eng(eq, orc.make_mlp(51, [64, 64], seed=2, gain=2.0))
    tx = e.sample_tx(32, seed=2)
    y = e.label_gen(tx, 128, 128, seed=3)
    b = float(y.abs().median())
    yc = e.label_gen(tx, 128, 128, seed=3, sample_bound=b)
    assert torch.equal(yc, y.clamp(-b, b))
    assert float(yc.abs().max()) <= b
    e.close()


def test_burgers_1000d_sweep_shape_matches_oracle_on_replayed_draws():
    """BASELINE config 5 (Cha nx=1000, MLP [1001,128,128,128,128,1]) on the fp32 path: K chunks of 128 input dims."""
    from tests.gpu_util import draws_dict
    eq = orc.Cha(1000, 1.0, 5.0, 1.0)
    net = orc.make_mlp(1001, [128] * 4, seed=7, gain=1.2)
    n, M = 3, 24
    d = orc.draw(eq, n, M, M, None, torch.Generator().manual_seed(1), round_fp32=True)
    tx = orc.sample_tx(eq, d)
    y_ref = orc.generate_with_gradients(eq, net, tx, d, M, M)
    e = _eng(eq, net)
    y = e.label_gen(tx, M, M, precision="fp32", draws=draws_dict(d)).double().cpu()
    assert orc.rel_l2(y, y_ref) < 1e-4
    assert orc.rel_l2(y[:, 1:], y_ref[:, 1:]) < 1e-4
    # the tensor-core kernel at this width: first layer accumulated over 8 chunks of 128 input dims
    y16 = e.label_gen(tx, M, M, precision="f16", draws=draws_dict(d)).double().cpu()
    assert orc.rel_l2(y16, y_ref) < 1e-2
    assert orc.rel_l2(y16[:, 1:], y_ref[:, 1:]) < 1e-2
    # Philox mode at this width: shard-invariant
    txp = e.sample_tx(6, seed=3)
    yp = e.label_gen(txp, 32, 32, seed=5)
    yq = torch.cat([e.label_gen(txp[:2], 32, 32, seed=5), e.label_gen(txp[2:], 32, 32, seed=5, point_index_base=2)])
    assert torch.equal(yp, yq) and torch.isfinite(yp).all()
    e.close()


@pytest.mark.parametrize("nx", [130, 200, 520, 777, 1000, 1024])  # 2, 2, 5, 7, 8, 8 chunks; 130 and 777: ragged last octet
def test_chunked_tensor_core_path_tracks_fp32_on_the_same_philox_draws(nx):
    """128 < d <= 1024 on tcgen05: the first layer is accumulated over ceil(d/128) chunks and the increments wait for
    their Bismut coefficients in a per-CTA ring in global memory.  Same Philox counters as the fp32 kernel, so the two
    differ only by operand rounding; bitwise shard invariance (which is also what catches a race on the ring);
    ragged sample counts."""
    eq = orc.Cha(nx, 1.0, 5.0, 1.0)
    net = orc.make_mlp(1 + nx, [128] * 4, seed=11, gain=1.3)
    e = _eng(eq, net)
    tx = e.sample_tx(10, seed=4)
    for M_I in (300, 2100):  # 3 tiles (one ragged) / 17 tiles = two units, the second with a single ragged tile
        y32 = e.label_gen(tx, 64, M_I, precision="fp32", seed=8).double().cpu()
        y16 = e.label_gen(tx, 64, M_I, precision="f16", seed=8).double().cpu()
        assert torch.isfinite(y16).all()
        assert orc.rel_l2(y16, y32) < 1e-2
        assert orc.rel_l2(y16[:, 1:], y32[:, 1:]) < 1e-2
        yb = torch.cat([e.label_gen(tx[:3], 64, M_I, precision="f16", seed=8),
                        e.label_gen(tx[3:], 64, M_I, precision="f16", seed=8, point_index_base=3)]).double().cpu()
        assert torch.equal(y16, yb)
    with pytest.raises(NotImplementedError):
        e2 = _eng(orc.Cha(1100, 1.0, 5.0, 1.0), orc.make_mlp(1101, [128] * 4, seed=1))
        e2.label_gen(e2.sample_tx(2, seed=1), 8, 8, precision="f16")  # d > 1024: fp32 path only
    e.close()


def test_chunked_tensor_core_path_is_reproducible_run_to_run():
    """The increment ring of the d > 128 kernel is shared between warps (written by the generating threads, read by
    the warp that owns the chunk's Bismut sums): a race on it would show up as run-to-run differences.  Many tiles
    per CTA, several repetitions, bitwise equality."""
    eq = orc.Cha(300, 1.0, 5.0, 1.0)
    e = _eng(eq, orc.make_mlp(301, [128] * 4, seed=2, gain=1.2))
    tx = e.sample_tx(600, seed=5)  # 2 units per point at M_I = 2500 -> ~8 units = 80 tiles per CTA
    y0 = e.label_gen(tx, 16, 2500, precision="f16", seed=4)
    assert torch.isfinite(y0).all()
    for _ in range(4):
        assert torch.equal(e.label_gen(tx, 16, 2500, precision="f16", seed=4), y0)
    e.close()


def test_large_call_is_chunked_transparently():
    """More points than the unit-partial scratch holds at once (256 MB): the call is split inside the library and the
    labels equal those of explicit shards."""
    eq = orc.Cha(100, 1.0, 5.0, 1.0)
    e = _eng(eq, orc.make_mlp(101, [128] * 4, seed=3))
    n = 20000  # M_I = 512 -> 1 unit per point, 16 KB per unit: 16384 points per chunk
    tx = e.sample_tx(n, seed=9)
    y = e.label_gen(tx, 32, 512, precision="f16", seed=2)
    cut = 7777
    y_b = torch.cat([e.label_gen(tx[:cut], 32, 512, precision="f16", seed=2),
                     e.label_gen(tx[cut:], 32, 512, precision="f16", seed=2, point_index_base=cut)])
    assert torch.equal(y, y_b)
    e.close()


def test_full_size_labels_are_unbiased_for_the_exact_solution():
    """Size-independent property at the BASELINE sizes (M = 4096): if u_k is the exact solution, the Picard map
    reproduces it, so labels generated from a network that equals u* must scatter around (u*, grad u*) with zero mean.
    Used here with u_k == 0 replaced by the terminal-only identity: E[g(X_T)] for Cha is available in closed form only
    through u* itself, so we check the first Picard iterate's value labels against a high-accuracy MC of the oracle."""
    eq = orc.Cha(100, 1.0, 5.0, 1.0)
    e = _eng(eq, orc.ZeroNet())
    n, M = 256, 4096
    tx = e.sample_tx(n, seed=4)
    y_a = e.label_gen(tx, M, M, seed=1).double().cpu()
    y_b = e.label_gen(tx, M, M, seed=2).double().cpu()
    # two independent full-size runs: their difference is pure MC noise with the predicted 1/sqrt(M) scale
    dv = (y_a[:, 0] - y_b[:, 0])
    assert abs(float(dv.mean())) < 4 * float(dv.std()) / n ** 0.5
    assert float(dv.std()) < 0.5 * (2.0 / M) ** 0.5  # |g| <= 1 bounds the per-sample variance by 1/4
    # against the fp64 oracle with its own draws
    d = orc.draw(eq, n, M, 1, None, torch.Generator().manual_seed(0))
    y_o = orc.estimate_terminal_with_gradients(eq, tx.double().cpu(), d.zT, M)
    dv = (y_a[:, 0] - y_o[:, 0])
    assert abs(float(dv.mean())) < 4 * float(dv.std()) / n ** 0.5
    e.close()


def test_gbm_and_hjb_zero_network_iterations_are_shard_invariant():
    from tests.golden_util import Golden
    for name, v in (("gbm_100d_zero_sdgd", 100), ("hjb_100d_zero", 0)):
        g = Golden(name)
        e = _eng(g.eq, orc.ZeroNet(), v)
        tx = e.sample_tx(40, seed=6)
        y = e.label_gen(tx, 256, 256, seed=8)
        y_b = torch.cat([e.label_gen(tx[:13], 256, 256, seed=8), e.label_gen(tx[13:], 256, 256, seed=8, point_index_base=13)])
        assert torch.equal(y, y_b) and torch.isfinite(y).all()
        e.close()


def test_hjb_narrow_network_with_a_single_integral_sample():
    """PISGradNet path at the constructor default n_estimate_integral = 1 with a narrow nn_module: the MC reduction's
    slice partials live in their own scratch (they used to overrun the activation buffer when
    8 (1 + d) > (M_I + 1) hmax); checked against the oracle on replayed draws, fp32 and tf32."""
    from tests.golden_util import Golden
    from tests.gpu_util import draws_dict
    g = Golden("hjb_100d_pisgrad")
    net = orc.make_pisgrad(100, [64, 64], T=1.0, seed=5, phase_scale=0.3)
    e = _eng(g.eq, net)
    gen = torch.Generator().manual_seed(8)
    for M_T, M_I, n in ((7, 1, 9), (3, 2, 33)):
        d = orc.draw(g.eq, n, M_T, M_I, None, gen, round_fp32=True)
        tx = orc.sample_tx(g.eq, d)
        ref = orc.generate_with_gradients(g.eq, net, tx, d, M_T, M_I)
        for prec, tol in (("fp32", 1e-4), ("tf32", 1e-2)):
            y = e.label_gen(tx, M_T, M_I, precision=prec, draws=draws_dict(d)).double().cpu()
            assert torch.isfinite(y).all() and orc.rel_l2(y, ref) < tol, (prec, M_T, M_I, orc.rel_l2(y, ref))
    e.close()


@pytest.mark.parametrize("with_tx", [True, False])
def test_large_host_call_is_pipelined_and_bitwise_the_device_result(with_tx):
    """dpi_label_gen_host splits calls of >= 256 MB into ~64 MB chunks (copies on two handle streams overlapping the
    kernels).  Labels are keyed by the global point index, so the pipelined call must reproduce the one-shot
    device-buffer call bit for bit -- with caller-supplied points and with points sampled on the device."""
    from tests.gpu_util import engine_for
    nx, n, M = 1000, 70000, 4  # 70000 x 1001 x 4 B = 280 MB each way, 5 chunks
    eq = orc.Cha(nx, 1.0, 5.0, 1.0)
    e = engine_for(eq, orc.make_mlp(1 + nx, [128] * 4, seed=3))
    if with_tx:
        tx = e.sample_tx(n, seed=4, point_index_base=11)
        tx_host = tx.cpu().pin_memory()
        _, y_host = e.label_gen_host(n, M, M, tx_host=tx_host, precision="f16", seed=9, point_index_base=11)
    else:
        tx_host, y_host = e.label_gen_host(n, M, M, precision="f16", seed=9, point_index_base=11)
        tx = e.sample_tx(n, seed=9, point_index_base=11)
        assert torch.equal(tx.cpu(), tx_host)
    y = e.label_gen(tx, M, M, precision="f16", seed=9, point_index_base=11)
    assert torch.isfinite(y_host).all()
    assert torch.equal(y.cpu(), y_host)
    e.close()
